#!/bin/bash
tag=${1:-r2y}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() {  # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 400 --warmup 20 --reps 5 --no-cpu $BARGS > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; echo "$name rc=$?"
}
BARGS=""
run c2_t4_ldg_sp2 PGX_TILE=4 PGX_TILE_TMA=0 PGX_TILE_SPLIT=2
run c2_t4_ldg_sp4 PGX_TILE=4 PGX_TILE_TMA=0 PGX_TILE_SPLIT=4
run c2_t4_ldg_sp8 PGX_TILE=4 PGX_TILE_TMA=0 PGX_TILE_SPLIT=8
run c2_t4_s2p2g4_c2 PGX_TILE=4 PGX_TILE_ST=2 PGX_TILE_TPS=2 PGX_TILE_STAGES=4 PGX_TILE_CTAS=2
run c2_t4_s2p2g4_c3 PGX_TILE=4 PGX_TILE_ST=2 PGX_TILE_TPS=2 PGX_TILE_STAGES=4 PGX_TILE_CTAS=3
run c2_t4_s2p2g4_c2_sp1 PGX_TILE=4 PGX_TILE_ST=2 PGX_TILE_TPS=2 PGX_TILE_STAGES=4 PGX_TILE_CTAS=2 PGX_TILE_SPLIT=1
run c2_t4_s2p4g3_c2 PGX_TILE=4 PGX_TILE_ST=2 PGX_TILE_TPS=4 PGX_TILE_STAGES=3 PGX_TILE_CTAS=2
run c2_t4_s4p2g4_c1 PGX_TILE=4 PGX_TILE_ST=4 PGX_TILE_TPS=2 PGX_TILE_STAGES=4 PGX_TILE_CTAS=1
run c2_t4_s4p2g4_c2_sp2 PGX_TILE=4 PGX_TILE_ST=4 PGX_TILE_TPS=2 PGX_TILE_STAGES=4 PGX_TILE_CTAS=2 PGX_TILE_SPLIT=2
BARGS="--variant distinct"
run c2d_t4_ldg_sp4 PGX_TILE=4 PGX_TILE_TMA=0 PGX_TILE_SPLIT=4
run c2d_t4_ldg_sp8 PGX_TILE=4 PGX_TILE_TMA=0 PGX_TILE_SPLIT=8
run c2d_t4_p4g3_c2 PGX_TILE=4 PGX_TILE_TPS=4 PGX_TILE_STAGES=3 PGX_TILE_CTAS=2
run c2d_t4_p4g4_c2 PGX_TILE=4 PGX_TILE_TPS=4 PGX_TILE_STAGES=4 PGX_TILE_CTAS=2
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-28s %.4f ms/step  value %.0f  e2e %.0f  parity %.2e  rf %.2f %s"%(f.split("${tag}_")[1], d["ms_per_step"], d["value"], d["e2e"]["value"], d["parity"]["max_rel_err"], d["roofline"]["frac"], d["roofline"]["launch_plan"]))
    except Exception as e: print(f,"ERR",e)
PY
