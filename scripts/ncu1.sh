cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_r2c|k_c2r" --launch-skip 12 --launch-count 2 -f -o gpurun_out/fft_c1_r01b python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_fft_c1.log 2>&1
tail -3 gpurun_out/ncu_fft_c1.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_r2c|k_c2r|k_fdl" --launch-skip 12 --launch-count 3 -f -o gpurun_out/c3_r01b python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_c3.log 2>&1
tail -3 gpurun_out/ncu_c3.log
