#!/bin/bash
# 1 -> N GPU scaling of the headline bench (and C4 with its NCCL reduce): gpurun --gpus 8 -- 'bash scripts/scale.sh 8'
n=${1:-8}
cd "${GRAFT_REPO_ROOT:-.}"
for g in 1 2 4 8; do
  [ $g -gt $n ] && break
  if [ $g -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 1000 --warmup 20 --no-cpu > gpurun_out/scale_c2_$g.json 2> gpurun_out/scale_c2_$g.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29600+g)) bench.py --gpus $g --steps 1000 --warmup 20 --no-cpu > gpurun_out/scale_c2_$g.json 2> gpurun_out/scale_c2_$g.err
  fi
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $n --steps 1000 --warmup 20 --workload c4 --no-cpu > gpurun_out/scale_c4_$n.json 2> gpurun_out/scale_c4_$n.err
tail -n 2 gpurun_out/scale_*.err || true
