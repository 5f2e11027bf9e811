cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t2.log 2>&1; echo "rc=$?" >> gpurun_out/t2.log
tail -15 gpurun_out/t2.log
for w in c2 c3 c4 c1; do
  timeout 300 python bench.py --steps 1000 --warmup 20 --workload $w --no-cpu > gpurun_out/g2_$w.json 2> gpurun_out/g2_$w.err
done
timeout 300 python bench.py --steps 1000 --warmup 20 --variant distinct --no-cpu > gpurun_out/g2_c2d.json 2> gpurun_out/g2_c2d.err
