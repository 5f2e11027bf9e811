"""
Worker of tests/test_gpu_multi.py::test_sharded_mix_torchrun -- launched as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P tests/mp_sharded_mix.py
One process per GPU.  N mono streams with distinct IRs are sharded over the ranks; every pull is a pipelined
host-buffer submit with PGX_PULL_MIX | PGX_PULL_REDUCE: the partial mixes are summed onto rank 0 through peer
memory (pgx_mix_reduce) and only rank 0 copies a result back.  Rank 0 checks the reduced mix against the oracle
sum over ALL streams (1e-5 of full scale) and against the library collective (NCCL) on the same partials.
Exit code 0 = parity green on every rank.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist

    import pygmu2_b200 as pg
    import pygmu2_oracle as orc  # checker
    from pygmu2_b200 import dist as pd, workloads as wl
    from pygmu2_b200._lib import PinnedArray

    rank, world, local = pd.init_process_group("nccl")
    pg.set_sample_rate(44_100)
    N, L, B, pull, pulls = 8 * world + 3, 3000, 256, 512, 14          # ragged shard sizes, 12 partitions
    sm = pd.ShardedMix(N, make_bank=lambda lo, hi: pg.ConvolveBank(
        np.stack([wl.c4_ir(s, L) for s in range(lo, hi)]), hi - lo, 1, block=B, max_pull=pull, device=local))
    comm = pd.MixComm(local, rank, world, root=0, max_floats=pull)
    sm.attach_comm(comm)
    n_loc = sm.hi - sm.lo
    x_loc = np.stack([wl.c4_input(pull * pulls, s) for s in range(sm.lo, sm.hi)])[:, None, :]
    xp = [PinnedArray((n_loc, 1, pull)) for _ in range(3)]
    yp = [PinnedArray((1, pull)) for _ in range(3)]
    out, tickets = [], []
    dist.barrier()
    for i in range(pulls):                                            # three pulls in flight
        if i >= 3:
            sm.wait(tickets[i - 3])
            if rank == 0:
                out.append(yp[(i - 3) % 3].array.copy())
        xp[i % 3].array[...] = x_loc[:, :, i * pull:(i + 1) * pull]
        tickets.append(sm.submit_mix(xp[i % 3].array, yp[i % 3].array))
    for i in range(max(pulls - 3, 0), pulls):
        sm.wait(tickets[i])
        if rank == 0:
            out.append(yp[i % 3].array.copy())
    comm.check()

    # the same partials through the library collective (baseline), on device tensors
    sm2 = pd.ShardedMix(N, make_bank=lambda lo, hi: pg.ConvolveBank(
        np.stack([wl.c4_ir(s, L) for s in range(lo, hi)]), hi - lo, 1, block=B, max_pull=pull, device=local))
    xd = torch.from_numpy(x_loc).cuda(local)
    yd = torch.empty((pulls, 1, pull), dtype=torch.float32, device=f"cuda:{local}")
    for i in range(pulls):       # on torch's DEFAULT stream: ShardedMix moves the pull to a fenced side stream
        sm2.render_mix_device(xd[:, :, i * pull:(i + 1) * pull].contiguous(), yd[i], pull)
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        y = np.concatenate(out, axis=1)[0]
        ref = np.zeros(pull * pulls)
        for s in range(N):
            ref += orc.OracleConvolve(wl.c4_ir(s, L), 1).render(wl.c4_input(pull * pulls, s)).astype(np.float64)[:, 0]
        err = float(np.max(np.abs(y - ref)) / np.max(np.abs(ref)))
        y_nccl = yd.cpu().numpy().reshape(-1)
        err_nccl = float(np.max(np.abs(y_nccl - ref)) / np.max(np.abs(ref)))
        print(f"world {world}: {N} streams, peer-memory reduce err {err:.2e}, NCCL reduce err {err_nccl:.2e}", flush=True)
        ok = err <= 1e-5 and err_nccl <= 1e-5
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    sm.bank.close(); sm2.bank.close(); comm.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
