#!/bin/bash
# ncu --set full of the fused B=4096 step in its three variants (C1) and of k_mix1 (C3); each after a plain run.
tag=${1:-r02}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu --reps 3"
for v in 0 1 3; do
  PGX_FFT16=$v $B --workload c1 > gpurun_out/${tag}_plain_c1_v$v.log 2>&1 || { echo "plain c1 v$v failed"; tail -3 gpurun_out/${tag}_plain_c1_v$v.log; }
  PGX_FFT16=$v ncu --set full --clock-control none --import-source on -k regex:k_conv1 -s 10 -c 1 -f -o gpurun_out/${tag}_conv1_c1_v$v $B --workload c1 > gpurun_out/${tag}_ncu_c1_v$v.log 2>&1
done
$B --workload c3 > gpurun_out/${tag}_plain_c3.log 2>&1 || echo "plain c3 failed"
ncu --set full --clock-control none --import-source on -k regex:k_mix1 -s 10 -c 1 -f -o gpurun_out/${tag}_mix1_c3 $B --workload c3 > gpurun_out/${tag}_ncu_c3.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
tail -n 2 gpurun_out/${tag}_ncu_*.log
