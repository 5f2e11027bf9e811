#!/bin/bash
# Round-2 ncu evidence (B200_PROFILING.md recipe): launch lists of the bench command, one `--set full` capture per kernel of
# interest, and the LDG-vs-TMA accumulate A/B on the final kernels.  Every ncu run follows a plain run of the same command.
tag=${1:-r02}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu --reps 3"
$B > gpurun_out/${tag}_plain_shared.log 2>&1 || { echo plain failed; tail -3 gpurun_out/${tag}_plain_shared.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 780 -c 90 --csv --log-file gpurun_out/${tag}_launches_c2_shared.csv $B > gpurun_out/${tag}_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fdl_mac -s 262 -c 2 -f -o gpurun_out/${tag}_mac_shared $B > gpurun_out/${tag}_ncu_s.log 2>&1
$B --variant distinct > gpurun_out/${tag}_plain_distinct.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fdl_mac -s 262 -c 2 -f -o gpurun_out/${tag}_mac_distinct $B --variant distinct > gpurun_out/${tag}_ncu_d.log 2>&1
PGX_MAC=tma $B > gpurun_out/${tag}_plain_tma.log 2>&1
PGX_MAC=tma ncu --set full --clock-control none --import-source on -k regex:k_fdl_mac_tma -s 262 -c 2 -f -o gpurun_out/${tag}_mac_tma_shared $B > gpurun_out/${tag}_ncu_t.log 2>&1
$B --workload c4 > gpurun_out/${tag}_plain_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fdl_mac -s 350 -c 2 -f -o gpurun_out/${tag}_mac_c4 $B --workload c4 > gpurun_out/${tag}_ncu_c4.log 2>&1
$B --workload c1 > gpurun_out/${tag}_plain_c1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 30 --csv --log-file gpurun_out/${tag}_launches_c1.csv $B --workload c1 > gpurun_out/${tag}_ncu_l1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_conv1 -s 10 -c 1 -f -o gpurun_out/${tag}_conv1_c1_v3 $B --workload c1 > gpurun_out/${tag}_ncu_c1.log 2>&1
$B --workload c3 > gpurun_out/${tag}_plain_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 30 --csv --log-file gpurun_out/${tag}_launches_c3.csv $B --workload c3 > gpurun_out/${tag}_ncu_l3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mix1 -s 10 -c 1 -f -o gpurun_out/${tag}_mix1_c3_v2 $B --workload c3 > gpurun_out/${tag}_ncu_c3.log 2>&1
tail -n 1 gpurun_out/${tag}_ncu_*.log
# LDG vs TMA inside the step, final kernels
for m in ldg tma; do
  for v in shared distinct; do
    PGX_MAC=$m timeout 300 python bench.py --steps 500 --warmup 20 --variant $v --no-cpu > gpurun_out/${tag}_ab_${m}_${v}.json 2> gpurun_out/${tag}_ab_${m}_${v}.err
  done
  PGX_MAC=$m timeout 300 python bench.py --steps 500 --warmup 20 --workload c4 --no-cpu > gpurun_out/${tag}_ab_${m}_c4.json 2> gpurun_out/${tag}_ab_${m}_c4.err
done
ls gpurun_out/${tag}_*.ncu-rep
