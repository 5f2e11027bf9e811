cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t4.log 2>&1; echo "rc=$?" >> gpurun_out/t4.log
tail -15 gpurun_out/t4.log
for bg in 1 2; do
  for v in shared distinct; do
    PGX_BG_STREAMS=$bg timeout 300 python bench.py --steps 2000 --warmup 20 --variant $v --no-cpu > gpurun_out/g4_bg${bg}_$v.json 2> gpurun_out/g4_bg${bg}_$v.err
  done
  PGX_BG_STREAMS=$bg timeout 300 python bench.py --steps 1000 --warmup 20 --workload c4 --no-cpu > gpurun_out/g4_bg${bg}_c4.json 2> gpurun_out/g4_bg${bg}_c4.err
  PGX_BG_STREAMS=$bg timeout 300 python bench.py --steps 1000 --warmup 20 --workload c5 --streams 256 --no-cpu > gpurun_out/g4_bg${bg}_c5n.json 2> gpurun_out/g4_bg${bg}_c5n.err
done
