"""
Multi-GPU sharding of independent streams / sources.

One process per GPU (torchrun), ``torch.distributed`` for the plumbing.  Streams are
independent until the final MixPE sum (reference mix_pe.py:92-94), so rank g owns the
contiguous slice [g*N/G, (g+1)*N/G) of the stream index with its filter spectra and delay
lines resident, computes its partial mix locally (fused in the accumulate kernel) and the
only collective is one sum-reduce of the (C_out, n) float32 mix per pull -- NCCL over
NVLink on GPUs, gloo in the CPU tests.  Configurations without a mix (C2: independent
reverbs) shard with no collective at all.
"""
from __future__ import annotations

import os

import numpy as np


def shard_bounds(n_units: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous slice of [0, n_units) owned by ``rank``; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(int(n_units), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_rank() -> tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def bind_host_to_gpu(local_rank: int) -> bool:
    """Pin this process to the CPUs (and so, by first touch, its pinned staging buffers to the memory) of the
    NUMA node the GPU hangs off: with one process per GPU the host-buffer path (H2D / D2H of every pull)
    otherwise crosses the socket interconnect for half the ranks.  Best effort; returns whether it took."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return True
    except Exception:
        return False


def init_process_group(backend: str | None = None):
    """Initialise torch.distributed from the environment (nccl when CUDA is present, else gloo)."""
    import torch
    import torch.distributed as dist

    rank, world, local = env_rank()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if not dist.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def reduce_mix(mix, root: int | None = 0, group=None):
    """Sum the per-rank partial mixes.  ``mix`` is a torch tensor (CUDA for nccl, CPU for gloo),
    reduced in place: onto ``root`` (``dist.reduce``), or onto every rank when root is None."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return mix
    if root is None:
        dist.all_reduce(mix, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(mix, dst=root, op=dist.ReduceOp.SUM, group=group)
    return mix


class ShardedMix:
    """N streams sharded over the ranks of a process group, mixed to one (C_out, n) signal.

    ``make_bank(lo, hi)`` builds this rank's ConvolveBank / HrtfMixBank-like object for streams
    [lo, hi): it must provide ``process_device(x_ptr, y_ptr, n, mix=True, cuda_stream=...)`` (GPU) --
    or, for CPU tests, any callable ``local_mix(x_local) -> np.ndarray`` passed as ``local_mix``.
    """

    def __init__(self, n_streams: int, *, make_bank=None, local_mix=None, root: int | None = 0):
        import torch.distributed as dist

        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.n_streams = int(n_streams)
        self.lo, self.hi = shard_bounds(self.n_streams, self.world, self.rank)
        self.root = root
        self.bank = make_bank(self.lo, self.hi) if make_bank is not None else None
        self._local_mix = local_mix

    # host arrays (gloo / tests): x_local is this rank's (hi-lo, C_in, n) slice
    def render_mix_host(self, x_local: np.ndarray) -> np.ndarray:
        import torch

        if self._local_mix is not None:
            part = np.ascontiguousarray(self._local_mix(x_local), dtype=np.float32)
        else:
            part = self.bank.process_mix(x_local)
        t = torch.from_numpy(part)
        reduce_mix(t, self.root)
        return t.numpy()

    # device tensors (nccl): everything is enqueued on torch's current stream, no host sync
    def render_mix_device(self, x_local, y_mix, n: int):
        import torch

        st = torch.cuda.current_stream().cuda_stream
        self.bank.process_device(x_local.data_ptr(), y_mix.data_ptr(), n, mix=True, cuda_stream=st)
        reduce_mix(y_mix, self.root)
        return y_mix
