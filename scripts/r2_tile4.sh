#!/bin/bash
tag=${1:-r2x}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
PGX_TILE=4 PGX_TILE_MIN=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "time_tiled" > gpurun_out/${tag}_tests_new.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests_new.log
tail -5 gpurun_out/${tag}_tests_new.log
if ! grep -q "rc=0" gpurun_out/${tag}_tests_new.log; then exit 1; fi
PGX_TILE=4 PGX_TILE_MIN=0 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests_t4.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests_t4.log
tail -5 gpurun_out/${tag}_tests_t4.log
PGX_TILE=2 PGX_TILE_MIN=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/${tag}_tests_t2.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests_t2.log
tail -3 gpurun_out/${tag}_tests_t2.log
run() {  # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 400 --warmup 20 --reps 5 --no-cpu $BARGS > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; echo "$name rc=$?"
}
BARGS=""
run c2_t4_ldg PGX_TILE=4 PGX_TILE_TMA=0
run c2_t4_s2p2g3 PGX_TILE=4 PGX_TILE_ST=2 PGX_TILE_TPS=2 PGX_TILE_STAGES=3
run c2_t4_s2p2g4 PGX_TILE=4 PGX_TILE_ST=2 PGX_TILE_TPS=2 PGX_TILE_STAGES=4
run c2_t4_s2p4g3 PGX_TILE=4 PGX_TILE_ST=2 PGX_TILE_TPS=4 PGX_TILE_STAGES=3
run c2_t4_s4p2g3 PGX_TILE=4 PGX_TILE_ST=4 PGX_TILE_TPS=2 PGX_TILE_STAGES=3
run c2_t4_s4p2g4 PGX_TILE=4 PGX_TILE_ST=4 PGX_TILE_TPS=2 PGX_TILE_STAGES=4
run c2_t4_s2p2g3_sp2 PGX_TILE=4 PGX_TILE_ST=2 PGX_TILE_TPS=2 PGX_TILE_STAGES=3 PGX_TILE_SPLIT=2
run c2_t4_s4p2g3_sp2 PGX_TILE=4 PGX_TILE_ST=4 PGX_TILE_TPS=2 PGX_TILE_STAGES=3 PGX_TILE_SPLIT=2
run c2_t2_s2p2g4 PGX_TILE=2 PGX_TILE_ST=2 PGX_TILE_TPS=2 PGX_TILE_STAGES=4
run c2_t2_s4p2g4 PGX_TILE=2 PGX_TILE_ST=4 PGX_TILE_TPS=2 PGX_TILE_STAGES=4
BARGS="--variant distinct"
run c2d_t4_ldg PGX_TILE=4 PGX_TILE_TMA=0
run c2d_t4_p4g3 PGX_TILE=4 PGX_TILE_TPS=4 PGX_TILE_STAGES=3
run c2d_t4_p4g4 PGX_TILE=4 PGX_TILE_TPS=4 PGX_TILE_STAGES=4
run c2d_t4_p2g4 PGX_TILE=4 PGX_TILE_TPS=2 PGX_TILE_STAGES=4
run c2d_t2_p4g4 PGX_TILE=2 PGX_TILE_TPS=4 PGX_TILE_STAGES=4
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-28s %.4f ms/step  value %.0f  e2e %.0f  parity %.2e  rf %.2f %s"%(f.split("${tag}_")[1], d["ms_per_step"], d["value"], d["e2e"]["value"], d["parity"]["max_rel_err"], d["roofline"]["frac"], d["roofline"]["launch_plan"]))
    except Exception as e: print(f,"ERR",e)
PY
