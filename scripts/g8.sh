cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --steps 1000 --warmup 20 --workload c5v --no-cpu > gpurun_out/g8_c5v.json 2> gpurun_out/g8_c5v.err
tail -3 gpurun_out/g8_c5v.err
timeout 300 python bench.py --steps 2000 --warmup 20 --reverb --no-cpu > gpurun_out/g8_c2rev.json 2> gpurun_out/g8_c2rev.err
timeout 300 python bench.py --steps 2000 --warmup 20 --no-cpu > gpurun_out/g8_c2.json 2> gpurun_out/g8_c2.err
timeout 300 python bench.py --steps 1000 --warmup 20 --workload c5 --no-cpu > gpurun_out/g8_c5.json 2> gpurun_out/g8_c5.err
