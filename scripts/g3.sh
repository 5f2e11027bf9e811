cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t3.log 2>&1; echo "rc=$?" >> gpurun_out/t3.log
tail -15 gpurun_out/t3.log
for w in c4 c3; do
  timeout 300 python bench.py --steps 1000 --warmup 20 --workload $w --no-cpu > gpurun_out/g3_$w.json 2> gpurun_out/g3_$w.err
done
