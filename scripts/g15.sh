cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python scripts/named_configs.py --seconds 2 > gpurun_out/named2.jsonl 2> gpurun_out/named2.err; tail -n 3 gpurun_out/named2.err
