#!/bin/bash
tag=${1:-r02d}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 120 python scripts/named_configs.py --seconds 2 > gpurun_out/${tag}_named_configs.jsonl 2> gpurun_out/${tag}_named.err; echo "named rc=$?"
timeout 60 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${tag}_bench_c2_k20.json 2> gpurun_out/${tag}_bench_c2_k20.err; echo "c2 rc=$?"
timeout 150 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${tag}_gpu_tests.log
python - <<PY
import json
for l in open("gpurun_out/${tag}_named_configs.jsonl"):
    d=json.loads(l); print("  %-60s %8.1f us/pull"%(d["config"][:60], d["ms_per_pull"]*1e3))
d=json.loads(open("gpurun_out/${tag}_bench_c2_k20.json").read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["parity"]["max_rel_err"])
PY
