#!/bin/bash
tag=${1:-r2d}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/${tag}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -25 gpurun_out/${tag}_tests.log
timeout 600 python scripts/named_configs.py --seconds 2 > gpurun_out/${tag}_named.jsonl 2> gpurun_out/${tag}_named.err; echo "named rc=$?"
python - <<PY
import json
for l in open("gpurun_out/${tag}_named.jsonl"):
    d=json.loads(l); print("  %-60s %8.1f us/pull  x%.0f realtime"%(d["config"][:60], d["ms_per_pull"]*1e3, d["x_realtime"]))
PY
timeout 600 python scripts/prof_host.py > gpurun_out/${tag}_prof_host.txt 2>&1; echo "prof rc=$?"
timeout 1200 python scripts/soak.py --seeds 8 --steps 100 > gpurun_out/${tag}_soak.log 2>&1; echo "soak rc=$?"; tail -14 gpurun_out/${tag}_soak.log
