#!/bin/bash
# FFT kernel A/B on one box: C1 (4096 streams x 4096-tap FIR, B=4096) with the radix-8 kernel and the two radix-16 variants.
tag=${1:-r2f}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "c1 or 4096" > gpurun_out/${tag}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -15 gpurun_out/${tag}_tests.log
for rep in 1 2; do
for v in 0 1 2 3 4; do
  PGX_FFT16=$v timeout 300 python bench.py --steps 200 --warmup 20 --workload c1 --no-cpu > gpurun_out/${tag}_c1_v${v}_$rep.json 2> gpurun_out/${tag}_c1_v${v}_$rep.err; echo "v$v rc=$?"
done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/'+"${tag}"+'_c1_v*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'ms/step %.4f'%d['ms_per_step'], 'parity %.2e'%d['parity']['max_rel_err'], 'frac %.3f'%d['roofline']['step']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as e: print(f,'ERR',e)
PY
