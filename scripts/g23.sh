cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in c1 c3 c2 c4; do
timeout 300 python bench.py --steps 1000 --warmup 20 --workload $w --no-cpu > gpurun_out/g23_$w.json 2> gpurun_out/g23_$w.err
done
