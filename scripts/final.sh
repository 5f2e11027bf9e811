#!/bin/bash
# End-of-round validation: parity tests, smoke, every bench line, named configs.
cd "${GRAFT_REPO_ROOT:-.}"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -n 2 gpurun_out/final_smoke.log
bash scripts/gpu_check.sh final
python scripts/named_configs.py --seconds 2 > gpurun_out/final_named.jsonl 2> gpurun_out/final_named.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err
tail -n 2 gpurun_out/final_*.err
