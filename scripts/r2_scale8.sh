#!/bin/bash
# 8-GPU visit (gpurun --gpus 8): multi-GPU tests at world 2 and 8, then the driver-shaped bench at N = 1, 2, 4, 8 (the N > 1
# lines carry the c4 sub-record with the peer-memory reduce), the NCCL-reduce baseline at 8, and a long run at 8.
tag=${1:-r2s}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -4 gpurun_out/${tag}_tests.log
run() { n=$1; o=$2; shift 2
  if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/${tag}_$o.json 2> gpurun_out/${tag}_$o.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n "$@" > gpurun_out/${tag}_$o.json 2> gpurun_out/${tag}_$o.err; fi
  echo "$o rc=$?"; }
for n in 1 2 4 8; do run $n n${n}_k20 --steps 20 --warmup 5 --no-cpu; done
run 8 n8_k20_nccl --steps 20 --warmup 5 --no-cpu --reduce nccl
run 8 n8_k500 --steps 500 --warmup 20 --no-cpu
WORLD8=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 tests/mp_sharded_mix.py > gpurun_out/${tag}_mp8.log 2>&1; echo "mp8 rc=$?"; tail -2 gpurun_out/${tag}_mp8.log
tail -n 3 gpurun_out/${tag}_n8_k20.err
