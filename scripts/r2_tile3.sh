#!/bin/bash
tag=${1:-r2w}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
PGX_TILE=4 PGX_TILE_MIN=0 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests_t4.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests_t4.log
tail -8 gpurun_out/${tag}_tests_t4.log
PGX_TILE=2 PGX_TILE_MIN=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/${tag}_tests_t2.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests_t2.log
tail -4 gpurun_out/${tag}_tests_t2.log
run() {  # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 400 --warmup 20 --reps 5 --no-cpu $BARGS > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; echo "$name rc=$?"
}
BARGS=""
run c2_t1 PGX_TILE=1
run c2_t2_s2u4 PGX_TILE=2 PGX_TILE_ST=2 PGX_TILE_U=4
run c2_t2_s4u2 PGX_TILE=2 PGX_TILE_ST=4 PGX_TILE_U=2
run c2_t4_s2u4 PGX_TILE=4 PGX_TILE_ST=2 PGX_TILE_U=4
run c2_t4_s2u2 PGX_TILE=4 PGX_TILE_ST=2 PGX_TILE_U=2
BARGS="--variant distinct"
run c2d_t1 PGX_TILE=1
run c2d_t2_u4 PGX_TILE=2 PGX_TILE_U=4
run c2d_t2_u8 PGX_TILE=2 PGX_TILE_U=8
run c2d_t4_u2 PGX_TILE=4 PGX_TILE_U=2
run c2d_t4_u4 PGX_TILE=4 PGX_TILE_U=4
BARGS="--workload c4"
run c4_t1 PGX_TILE=1
run c4_t4 PGX_TILE=4
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-28s %.4f ms/step  value %.0f  e2e %.0f  parity %.2e  rf %.2f %s"%(f.split("${tag}_")[1], d["ms_per_step"], d["value"], d["e2e"]["value"], d["parity"]["max_rel_err"], d["roofline"]["frac"], d["roofline"]["launch_plan"]))
    except Exception as e: print(f,"ERR",e)
PY
PGX_TILE=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fdl_mac_tile -s 6 -c 2 -f -o gpurun_out/${tag}_tile4_c2 python bench.py --steps 40 --warmup 5 --reps 1 --no-cpu > gpurun_out/${tag}_ncu1.log 2>&1; echo "ncu1 rc=$?"
