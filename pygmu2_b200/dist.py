"""
Multi-GPU sharding of independent streams / sources.

One process per GPU (torchrun), ``torch.distributed`` for the plumbing.  Streams are
independent until the final MixPE sum (reference mix_pe.py:92-94), so rank g owns the
contiguous slice [g*N/G, (g+1)*N/G) of the stream index with its filter spectra and delay
lines resident, computes its partial mix locally (fused in the accumulate kernel) and the
only exchange is one sum of the (C_out, n) float32 partial mixes per pull.  On GPUs that sum
is ``MixComm`` / ``pgx_mix_reduce``: the ranks store their partials into the root's memory over
NVLink and the root adds them in rank order, one small kernel per rank on the bank's stream
(csrc/pgx_comm.cu) -- ``torch.distributed`` only carries the 128-byte mailbox handles at set-up.
``reduce_mix`` (a library collective: NCCL on GPUs, gloo in the CPU tests) stays as the baseline
the kernel path is measured against.  Configurations without a mix (C2: independent reverbs)
shard with no collective at all.
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


class MixComm:
    """The cross-GPU mix reduce of ``libpgx`` (pgx_comm_*): mailboxes in peer memory, rank-ordered float32 sum.

    ``exchange(blob: bytes) -> list[bytes]`` all-gathers the per-rank handle blobs in rank order; the default uses
    ``torch.distributed.all_gather_object`` on the initialised process group.  Several devices inside ONE process
    connect with ``MixComm.connect_local([comm0, comm1, ...])`` instead.
    """

    def __init__(self, device: int, rank: int, world: int, root: int = 0, max_floats: int = 2 * 8192, exchange=None,
                 connect: bool = True):
        import ctypes as C

        from . import _lib
        self._lib, self._C = _lib, C
        self.device, self.rank, self.world, self.root = int(device), int(rank), int(world), int(root)
        self.max_floats = int(max_floats)
        _lib.require_device()
        self._h = C.c_void_p()
        blob = (C.c_char * _lib.PGX_COMM_HANDLE_BYTES)()
        _lib.check(_lib.lib().pgx_comm_create(C.byref(self._h), self.device, self.rank, self.world, self.root,
                                              self.max_floats, blob))
        self.handle = bytes(blob)
        if connect and self.world > 1:
            if exchange is None:
                import torch.distributed as dist

                def exchange(b):
                    out = [None] * self.world
                    dist.all_gather_object(out, b)
                    return out
            self.connect(exchange(self.handle))

    def connect(self, handles) -> None:
        blob = b"".join(handles)
        if len(blob) != self.world * self._lib.PGX_COMM_HANDLE_BYTES:
            raise ValueError("need one handle blob per rank")
        self._lib.check(self._lib.lib().pgx_comm_connect(self._h, blob))

    @staticmethod
    def connect_local(comms) -> None:
        """Ranks living in one process (one device each): hand every communicator all the handles."""
        handles = [c.handle for c in sorted(comms, key=lambda c: c.rank)]
        for c in comms:
            c.connect(handles)

    @property
    def is_root(self) -> bool:
        return self.rank == self.root

    def reduce_device(self, part_ptr: int, y_ptr: int, n: int, cuda_stream: int = 0) -> None:
        """Enqueue the sum of the ranks' ``n``-float partials (device pointers) onto the root's ``y_ptr``."""
        C = self._C
        self._lib.check(self._lib.lib().pgx_mix_reduce(self._h, C.c_void_p(part_ptr), C.c_void_p(y_ptr), int(n),
                                                       C.c_void_p(cuda_stream)))

    def check(self) -> None:
        self._lib.check(self._lib.lib().pgx_comm_check(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.lib().pgx_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


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

    # device tensors: everything is enqueued on torch's current stream, no host sync.  With a MixComm attached
    # to the bank (``attach_comm``) the sum is the peer-memory kernel path; otherwise the library collective.
    def attach_comm(self, comm: "MixComm") -> None:
        self.comm = comm
        self.bank.attach_comm(comm)

    def render_mix_device(self, x_local, y_mix, n: int):
        import torch

        cur = torch.cuda.current_stream()
        if cur.cuda_stream == 0:
            # the legacy default stream has handle 0, which the C ABI reads as "the bank's own stream": work enqueued
            # there would not be ordered with torch's.  Run the pull (and the collective) on a side stream that is
            # fenced against the current one on both sides.
            side = getattr(self, "_side", None)
            if side is None:
                side = self._side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self._render_on(side.cuda_stream, x_local, y_mix, n)
            cur.wait_stream(side)
            x_local.record_stream(side)
            y_mix.record_stream(side)
            return y_mix
        self._render_on(cur.cuda_stream, x_local, y_mix, n)
        return y_mix

    def _render_on(self, st: int, x_local, y_mix, n: int) -> None:
        if getattr(self, "comm", None) is not None:
            self.bank.process_device(x_local.data_ptr(), y_mix.data_ptr(), n, mix=True, cuda_stream=st, reduce=True)
            return
        self.bank.process_device(x_local.data_ptr(), y_mix.data_ptr(), n, mix=True, cuda_stream=st)
        reduce_mix(y_mix, self.root)

    # host buffers, pipelined: pinned x_local (hi-lo, C_in, n) -> the root's pinned out (C_out, n); returns a ticket
    def submit_mix(self, x_local: np.ndarray, out: np.ndarray) -> int:
        return self.bank.submit(x_local, out, mix=True, reduce=True)

    def wait(self, ticket: int) -> None:
        self.bank.wait(ticket)
