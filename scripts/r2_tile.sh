#!/bin/bash
# time-tiled accumulate pass: parity under forced tiling, then the same-box A/B on the headline workloads
tag=${1:-r2t}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "time_tiled" > gpurun_out/${tag}_tests_new.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests_new.log
tail -15 gpurun_out/${tag}_tests_new.log
PGX_TILE=4 PGX_TILE_MIN=0 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests_t4.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests_t4.log
tail -30 gpurun_out/${tag}_tests_t4.log
PGX_TILE=2 PGX_TILE_MIN=0 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests_t2.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests_t2.log
tail -8 gpurun_out/${tag}_tests_t2.log
for t in 1 2 4; do
  PGX_TILE=$t timeout 300 python bench.py --steps 500 --warmup 20 --reps 5 --no-cpu > gpurun_out/${tag}_c2_t$t.json 2> gpurun_out/${tag}_c2_t$t.err; echo "c2 t$t rc=$?"
  PGX_TILE=$t timeout 300 python bench.py --workload c2 --variant distinct --steps 300 --warmup 20 --reps 5 --no-cpu > gpurun_out/${tag}_c2d_t$t.json 2> gpurun_out/${tag}_c2d_t$t.err; echo "c2d t$t rc=$?"
done
python - <<PY
import json
for w in ("c2","c2d"):
  for t in (1,2,4):
    try:
        d=json.loads(open("gpurun_out/${tag}_%s_t%d.json"%(w,t)).read().strip().splitlines()[-1])
        print(w,t,d["ms_per_step"],d["value"],d["e2e"]["value"],d.get("parity"),d["roofline"]["frac"])
    except Exception as e: print(w,t,"ERR",e)
PY
