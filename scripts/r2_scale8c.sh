#!/bin/bash
# Reduced 8-GPU visit: the driver-shaped bench at N = 1, 2, 4, 8 (the N > 1 lines carry the c4 sub-record)
tag=${1:-r2s}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { n=$1; o=$2; shift 2
  if [ $n -eq 1 ]; then timeout 300 python bench.py --gpus 1 "$@" > gpurun_out/${tag}_$o.json 2> gpurun_out/${tag}_$o.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n "$@" > gpurun_out/${tag}_$o.json 2> gpurun_out/${tag}_$o.err; fi
  echo "$o rc=$?"; }
for n in 1 2 4 8; do run $n n${n}_k20 --steps 20 --warmup 5 --no-cpu; done
tail -n 3 gpurun_out/${tag}_n8_k20.err
