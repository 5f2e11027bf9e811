#!/bin/bash
# One GPU box visit: parity tests, then every workload's bench line into gpurun_out/<tag>_*.json.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh <tag>'
tag=${1:-chk}
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -4 gpurun_out/${tag}_tests.log
timeout 400 python bench.py --steps 2000 --warmup 20 > gpurun_out/${tag}_c2.json 2> gpurun_out/${tag}_c2.err
timeout 300 python bench.py --steps 2000 --warmup 20 --variant distinct --no-cpu > gpurun_out/${tag}_c2d.json 2> gpurun_out/${tag}_c2d.err
for w in c1 c3 c4 c5 c5v; do
  timeout 300 python bench.py --steps 1000 --warmup 20 --workload $w --no-cpu > gpurun_out/${tag}_$w.json 2> gpurun_out/${tag}_$w.err
done
timeout 300 python bench.py --steps 1000 --warmup 20 --workload c5 --streams 256 --no-cpu > gpurun_out/${tag}_c5n.json 2> gpurun_out/${tag}_c5n.err
cat gpurun_out/${tag}_*.err | tail -5
