#!/bin/bash
# ncu evidence for one round (B200_PROFILING.md recipe): launch list of the bench command, then one
# `--set full` capture per kernel of interest.  Every ncu run follows a plain run of the same command.
#   gpurun --timeout 2400 -- 'bash scripts/ncu_round.sh r01'
tag=${1:-rXX}
cd "${GRAFT_REPO_ROOT:-.}"
B="python bench.py --steps 20 --warmup 3 --no-cpu"
$B > gpurun_out/${tag}_plain_shared.log 2>&1 || exit 1
# launch list: skip the delay-line fill (259 steps x 3 launches) and take 90 launches of the steady state
ncu --metrics gpu__time_duration.sum --clock-control none -s 780 -c 90 --csv --log-file gpurun_out/${tag}_launches_c2_shared.csv $B > gpurun_out/${tag}_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fdl_mac -s 262 -c 2 -f -o gpurun_out/${tag}_mac_shared $B > gpurun_out/${tag}_ncu_s.log 2>&1
$B --variant distinct > gpurun_out/${tag}_plain_distinct.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_fdl_mac -s 262 -c 2 -f -o gpurun_out/${tag}_mac_distinct $B --variant distinct > gpurun_out/${tag}_ncu_d.log 2>&1
PGX_MAC=tma $B > gpurun_out/${tag}_plain_tma.log 2>&1 || exit 1
PGX_MAC=tma ncu --set full --clock-control none --import-source on -k regex:k_fdl_mac_tma -s 262 -c 2 -f -o gpurun_out/${tag}_mac_tma_shared $B > gpurun_out/${tag}_ncu_t.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_r2c|k_c2r" -s 524 -c 2 -f -o gpurun_out/${tag}_fft_c2 $B > gpurun_out/${tag}_ncu_f.log 2>&1
$B --workload c4 > gpurun_out/${tag}_plain_c4.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_fdl_mac -s 350 -c 2 -f -o gpurun_out/${tag}_mac_c4 $B --workload c4 > gpurun_out/${tag}_ncu_c4.log 2>&1
$B --workload c1 > gpurun_out/${tag}_plain_c1.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_conv1 -s 10 -c 1 -f -o gpurun_out/${tag}_conv1_c1 $B --workload c1 > gpurun_out/${tag}_ncu_c1.log 2>&1
$B --workload c5v > gpurun_out/${tag}_plain_c5v.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_blit|k_mix_sum" -s 400 -c 2 -f -o gpurun_out/${tag}_osc_c5v $B --workload c5v > gpurun_out/${tag}_ncu_o.log 2>&1
tail -n 2 gpurun_out/${tag}_ncu_*.log
