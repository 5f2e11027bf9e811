set -x
cd $GRAFT_REPO_ROOT
PGX_MAC=tma timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_tma.log 2>&1; echo "rc=$?" >> gpurun_out/t_tma.log
tail -3 gpurun_out/t_tma.log
for m in ldg tma; do
  for v in shared distinct; do
    PGX_MAC=$m timeout 300 python bench.py --steps 2000 --warmup 20 --variant $v --no-cpu > gpurun_out/ab2_${m}_${v}.json 2> gpurun_out/ab2_${m}_${v}.err
    PGX_DEBUG_SERIAL=1 PGX_MAC=$m timeout 300 python bench.py --steps 2000 --warmup 20 --variant $v --no-cpu > gpurun_out/ab2s_${m}_${v}.json 2> gpurun_out/ab2s_${m}_${v}.err
  done
  PGX_MAC=$m timeout 300 python bench.py --steps 1000 --warmup 20 --workload c4 --no-cpu > gpurun_out/ab2_${m}_c4.json 2> gpurun_out/ab2_${m}_c4.err
done
