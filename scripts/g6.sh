cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t6.log 2>&1; echo "rc=$?" >> gpurun_out/t6.log
tail -5 gpurun_out/t6.log
for w in c1 c3 c4 c5 c2; do
timeout 300 python bench.py --steps 1000 --warmup 20 --workload $w --no-cpu > gpurun_out/g6_$w.json 2> gpurun_out/g6_$w.err
done
