#!/bin/bash
# Final single-GPU visit of round 2 (after time tiling became the default): smoke, all GPU tests (default, forced tiling,
# bulk-async accumulate kernel), every bench workload and variant, the same-box untiled baseline, the named configurations,
# the reference arm, the ncu launch list and one full capture of the tiled kernel per filter layout.
tag=${1:-r02c}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${tag}_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${tag}_gpu_tests.log
PGX_TILE=4 PGX_TILE_MIN=0 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests_forced_tile4.log 2>&1; echo "tests(tile4 forced) rc=$?"; tail -1 gpurun_out/${tag}_tests_forced_tile4.log
PGX_TILE=2 PGX_TILE_MIN=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/${tag}_tests_forced_tile2.log 2>&1; echo "tests(tile2 forced) rc=$?"; tail -1 gpurun_out/${tag}_tests_forced_tile2.log
PGX_MAC=tma PGX_TILE_TMA=1 PGX_TILE_MIN=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/${tag}_gpu_tests_tma.log 2>&1; echo "tests(tma) rc=$?"; tail -1 gpurun_out/${tag}_gpu_tests_tma.log
b() { o=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/${tag}_bench_$o.json 2> gpurun_out/${tag}_bench_$o.err; echo "$o rc=$?"; }
b c2_k20 --steps 20 --warmup 5
b c2 --steps 2000 --warmup 20 --no-cpu
b c2_distinct --steps 500 --warmup 20 --variant distinct --no-cpu
PGX_TILE=1 b c2_untiled --steps 2000 --warmup 20 --no-cpu
PGX_TILE=1 b c2_distinct_untiled --steps 500 --warmup 20 --variant distinct --no-cpu
PGX_TILE=2 b c2_tile2 --steps 1000 --warmup 20 --no-cpu
b c2_reverb --steps 500 --warmup 20 --reverb --no-cpu
b c2_twolevel --steps 500 --warmup 20 --tail-block 4096 --no-cpu
b c2_pcm16 --steps 500 --warmup 20 --pcm16 --no-cpu
for w in c1 c3 c4 c5 c5v; do b $w --steps 500 --warmup 20 --workload $w --no-cpu; done
b c5_n256 --steps 500 --warmup 20 --workload c5 --streams 256 --no-cpu
b c3_n4096 --steps 500 --warmup 20 --workload c3 --streams 4096 --no-cpu
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; echo "ref rc=$?"
timeout 600 python scripts/named_configs.py --seconds 2 > gpurun_out/${tag}_named_configs.jsonl 2> gpurun_out/${tag}_named.err; echo "named rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench_*.json")):
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        p = d.get("parity")
        print("%-40s value %12.0f ms/step %.4f e2e %.2f rf %.2f step-frac %s parity %s" % (f.split("/")[-1], d["value"], d["ms_per_step"], d["e2e"].get("frac_of_value", 0) if isinstance(d.get("e2e"), dict) else 0, d.get("roofline", {}).get("frac", 0), d.get("roofline", {}).get("step", {}).get("frac"), p and "%.1e" % p["max_rel_err"]))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 2 gpurun_out/${tag}_bench_*.err | grep -v "^$" | tail -20
# ncu: every run under ncu follows a plain run of the same command (above)
B="python bench.py --steps 20 --warmup 3 --no-cpu --reps 3"
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 120 --csv --log-file gpurun_out/${tag}_launches_c2_shared.csv $B > gpurun_out/${tag}_ncu_l.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_fdl_mac_tile -s 20 -c 2 -f -o gpurun_out/${tag}_mac_tile4_shared $B > gpurun_out/${tag}_ncu_s.log 2>&1; echo "ncu shared rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_fdl_mac_tile -s 20 -c 2 -f -o gpurun_out/${tag}_mac_tile4_distinct $B --variant distinct > gpurun_out/${tag}_ncu_d.log 2>&1; echo "ncu distinct rc=$?"
ls -la gpurun_out/${tag}_*.ncu-rep
