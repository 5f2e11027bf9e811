#!/bin/bash
# Round-2 latency visit: the five named configurations through the PE API (graph replay on / off), FFT variants A/B,
# and the single-stream / small-mix benches.
tag=${1:-r2l}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python scripts/named_configs.py --seconds 2 > gpurun_out/${tag}_named_graph.jsonl 2> gpurun_out/${tag}_named_graph.err; echo "named rc=$?"
PGX_GRAPH=0 timeout 600 python scripts/named_configs.py --seconds 2 > gpurun_out/${tag}_named_nograph.jsonl 2> gpurun_out/${tag}_named_nograph.err; echo "named(nograph) rc=$?"
for w in c5 c3 c1; do
  timeout 300 python bench.py --steps 500 --warmup 20 --workload $w --no-cpu > gpurun_out/${tag}_$w.json 2> gpurun_out/${tag}_$w.err; echo "$w rc=$?"
  PGX_GRAPH=0 timeout 300 python bench.py --steps 500 --warmup 20 --workload $w --no-cpu > gpurun_out/${tag}_${w}_nograph.json 2> gpurun_out/${tag}_${w}_nograph.err; echo "$w nograph rc=$?"
done
timeout 300 python bench.py --steps 500 --warmup 20 --workload c5v --no-cpu > gpurun_out/${tag}_c5v.json 2> gpurun_out/${tag}_c5v.err; echo "c5v rc=$?"
bash scripts/r2_fft.sh ${tag}f
python - <<PY
import json
for f in ("named_graph","named_nograph"):
    print("==",f)
    for l in open("gpurun_out/${tag}_"+f+".jsonl"):
        d=json.loads(l); print("  %-60s %8.1f us/pull  x%.0f realtime"%(d["config"][:60], d["ms_per_pull"]*1e3, d["x_realtime"]))
PY
tail -n 3 gpurun_out/${tag}_*.err | tail -30
