#!/bin/bash
# Round-2 multi-GPU visit (gpurun --gpus N): multi-GPU parity tests, then the driver-shaped torchrun bench (with the c4 sub-record)
# with the peer-memory reduce and with the NCCL baseline.
tag=${1:-r2m}; n=${2:-2}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -15 gpurun_out/${tag}_tests.log
run() { # $1 = out tag, rest = bench args
  o=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29577 \
    bench.py --gpus $n "$@" > gpurun_out/${tag}_$o.json 2> gpurun_out/${tag}_$o.err; echo "$o rc=$?"
}
run k20 --steps 20 --warmup 5
run k20_nccl --steps 20 --warmup 5 --reduce nccl
run k500 --steps 500 --warmup 20
tail -n 8 gpurun_out/${tag}_k20.err
