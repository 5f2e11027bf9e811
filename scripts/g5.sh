cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t5.log 2>&1; echo "rc=$?" >> gpurun_out/t5.log
tail -5 gpurun_out/t5.log
for v in shared distinct; do
  timeout 300 python bench.py --steps 2000 --warmup 20 --variant $v --no-cpu > gpurun_out/g5_$v.json 2> gpurun_out/g5_$v.err
done
for w in c1 c3 c4 c5; do
timeout 300 python bench.py --steps 1000 --warmup 20 --workload $w --no-cpu > gpurun_out/g5_$w.json 2> gpurun_out/g5_$w.err
done
timeout 300 python bench.py --steps 1000 --warmup 20 --workload c5 --streams 256 --no-cpu > gpurun_out/g5_c5n.json 2> gpurun_out/g5_c5n.err
