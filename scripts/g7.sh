cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t7.log 2>&1; echo "rc=$?" >> gpurun_out/t7.log
tail -30 gpurun_out/t7.log
