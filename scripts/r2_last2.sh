#!/bin/bash
tag=${1:-r02e}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
PGX_TILE=4 PGX_TILE_MIN=0 timeout 70 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "time_tiled or full_plan_c2 or submit or device_queue or random_operation or reverb" > gpurun_out/${tag}_tests_forced.log 2>&1; echo "forced rc=$?"; tail -2 gpurun_out/${tag}_tests_forced.log
timeout 50 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${tag}_bench_c2_k20.json 2> gpurun_out/${tag}_bench_c2_k20.err; echo "c2 rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench_c2_k20.json").read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["frac_of_value"], d["parity"]["max_rel_err"])
PY
timeout 100 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${tag}_gpu_tests.log
