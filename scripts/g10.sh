cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t10.log 2>&1; echo "rc=$?" >> gpurun_out/t10.log
tail -4 gpurun_out/t10.log
timeout 300 python bench.py --steps 1000 --warmup 20 --workload c5v > gpurun_out/g10_c5v.json 2> gpurun_out/g10_c5v.err
tail -3 gpurun_out/g10_c5v.err
for w in c2 c4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 20 --workload $w --no-cpu > gpurun_out/g10_2gpu_$w.json 2> gpurun_out/g10_2gpu_$w.err
tail -3 gpurun_out/g10_2gpu_$w.err
done
