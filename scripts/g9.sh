cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t9.log 2>&1; echo "rc=$?" >> gpurun_out/t9.log
tail -5 gpurun_out/t9.log
timeout 300 python bench.py --steps 1000 --warmup 20 --workload c5v --no-cpu > gpurun_out/g9_c5v.json 2> gpurun_out/g9_c5v.err
tail -3 gpurun_out/g9_c5v.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 60 --csv --log-file gpurun_out/launches_c5v.csv python bench.py --steps 200 --warmup 5 --workload c5v --no-cpu > gpurun_out/ncu_c5v.log 2>&1
