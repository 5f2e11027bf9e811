#!/bin/bash
# Round-2 single-GPU visit: smoke, parity tests (no -x: list every failure), driver-shaped bench (20 steps) and the long one.
tag=${1:-r2}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_smoke.log
tail -3 gpurun_out/${tag}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/${tag}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -40 gpurun_out/${tag}_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_c2_k20.json 2> gpurun_out/${tag}_c2_k20.err; echo "bench rc=$?"
timeout 400 python bench.py --steps 2000 --warmup 20 --no-cpu > gpurun_out/${tag}_c2.json 2> gpurun_out/${tag}_c2.err; echo "bench rc=$?"
for w in c1 c3 c4 c5; do
  timeout 300 python bench.py --steps 500 --warmup 20 --workload $w --no-cpu > gpurun_out/${tag}_$w.json 2> gpurun_out/${tag}_$w.err; echo "$w rc=$?"
done
tail -5 gpurun_out/${tag}_*.err
