#!/bin/bash
tag=${1:-r2z}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
PGX_TILE=4 PGX_TILE_MIN=0 PGX_TILE_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "time_tiled or full_plan or random_operation or two_level" > gpurun_out/${tag}_tests_tma.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests_tma.log
tail -4 gpurun_out/${tag}_tests_tma.log
PGX_TILE=4 PGX_TILE_MIN=0 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests_t4.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests_t4.log
tail -4 gpurun_out/${tag}_tests_t4.log
run() {  # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 400 --warmup 20 --reps 5 --no-cpu $BARGS > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; echo "$name rc=$?"
}
BARGS=""
run c2_t4_ldg PGX_TILE=4
run c2_t2_ldg PGX_TILE=2
run c2_t4_tma_s2p4g3 PGX_TILE=4 PGX_TILE_TMA=1 PGX_TILE_ST=2 PGX_TILE_TPS=4 PGX_TILE_STAGES=3 PGX_TILE_CTAS=4
run c2_t4_tma_s2p4g2 PGX_TILE=4 PGX_TILE_TMA=1 PGX_TILE_ST=2 PGX_TILE_TPS=4 PGX_TILE_STAGES=2 PGX_TILE_CTAS=4
run c2_t4_tma_s2p2g3 PGX_TILE=4 PGX_TILE_TMA=1 PGX_TILE_ST=2 PGX_TILE_TPS=2 PGX_TILE_STAGES=3 PGX_TILE_CTAS=4
run c2_t4_tma_s2p4g3_128 PGX_TILE=4 PGX_TILE_TMA=1 PGX_TILE_ST=2 PGX_TILE_TPS=4 PGX_TILE_STAGES=3 PGX_TILE_CTAS=4 PGX_TILE_SEG=128
BARGS="--variant distinct"
run c2d_t4_ldg PGX_TILE=4
run c2d_t4_tma_p4g3 PGX_TILE=4 PGX_TILE_TMA=1 PGX_TILE_TPS=4 PGX_TILE_STAGES=3 PGX_TILE_CTAS=4
run c2d_t4_tma_p8g2 PGX_TILE=4 PGX_TILE_TMA=1 PGX_TILE_TPS=8 PGX_TILE_STAGES=2 PGX_TILE_CTAS=4
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        ins=d["roofline"]["instrumented"]
        print("%-28s %.4f ms/step  value %.0f  e2e %.0f  parity %.2e  rf %.2f mac %.4f %s"%(f.split("${tag}_")[1], d["ms_per_step"], d["value"], d["e2e"]["value"], d["parity"]["max_rel_err"], d["roofline"]["frac"], ins["k_fdl_mac_busy_ms_per_launch"], d["roofline"]["launch_plan"]))
    except Exception as e: print(f,"ERR",e)
PY
