#!/bin/bash
# second 8-GPU visit: e2e staging experiments (write-combined input buffers, int16 PCM) and the world-8 sharded-mix worker
tag=${1:-r2t}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { o=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29688 bench.py --gpus 8 "$@" > gpurun_out/${tag}_$o.json 2> gpurun_out/${tag}_$o.err; echo "$o rc=$?"; }
run wc --steps 20 --warmup 5 --no-cpu --no-c4 --wc
run plain --steps 20 --warmup 5 --no-cpu --no-c4
run pcm16 --steps 20 --warmup 5 --no-cpu --no-c4 --pcm16
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 tests/mp_sharded_mix.py > gpurun_out/${tag}_mp8.log 2>&1; echo "mp8 rc=$?"; grep "world 8" gpurun_out/${tag}_mp8.log
lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" ; nvidia-smi topo -m | head -12
python - <<PY
import json
for f in ("wc","plain","pcm16"):
    d=json.loads(open("gpurun_out/${tag}_"+f+".json").read().strip().splitlines()[-1])
    print(f, "value %.0f e2e %.0f frac %.2f"%(d["value"], d["e2e"]["value"], d["e2e"]["frac_of_value"]), ["%.1f"%x for x in d["e2e"]["copy_gbs_per_rank"]])
PY
