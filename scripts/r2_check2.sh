#!/bin/bash
# tests (all) + ncu captures of the fused B=4096 step variants and k_mix1
tag=${1:-r2b}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${tag}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -40 gpurun_out/${tag}_tests.log
bash scripts/r2_ncu_fft.sh r02
