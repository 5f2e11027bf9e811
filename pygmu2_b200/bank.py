"""
ConvolveBank / HrtfMixBank: Python owners of a device-resident ``pgx_bank``.

A bank is N independent audio streams advanced in lockstep by one C-ABI call per
pull: the batched form of the reference's one-PE-one-FFT loop (SURVEY.md §1:
"N streams = N Python objects each doing its own FFTs").  Bank-level arrays are
planar float32: x is (N, C_in, n), y is (N, C_out, n), a fused mix is (C_out, n).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import BankConfig, BankInfo, Layout, check, lib


def next_pow2(n: int) -> int:
    n = int(n)
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


def choose_block(filter_len: int, pull_hint: int | None) -> int:
    """Partition size B for a filter of ``filter_len`` taps pulled ``pull_hint`` samples at a time.

    Delay-line traffic per second is ~ L*sr/B rows and every pull costs at least one block step, so
    B = next_pow2(pull) is the sweet spot for the steady state.  The hint is only the FIRST pull of a PE, though,
    and that may be a probe (``render(0, 1)`` -- ConvolvePE issues one itself on sources without a channel
    count, convolve_pe.py:203-205): B is sticky, so it gets a floor of min(256, next_pow2(L)) and is raised until
    the filter has at most 2048 partitions.  Pulls shorter than B stay exact and zero-latency (partial block
    steps); pass ``block_size=`` to pin a smaller B (the named 64-sample configuration does).
    """
    want = next_pow2(int(pull_hint)) if pull_hint else 512
    B = int(min(max(16, want), 4096))
    B = max(B, min(256, next_pow2(int(filter_len))))
    while B < 4096 and -(-int(filter_len) // B) > 2048:
        B *= 2
    return B


class ConvolveBank:
    """N lock-stepped streams, each convolved with one of F resident filters.

    filters : (F, L) or (F, L, C_f) float32 (a single (L,) / (L, C_f) filter is F = 1)
    c_in    : source channels per stream
    Channel rules are ConvolvePE's (reference convolve_pe.py:207-223): mono filter ->
    every source channel; multi-channel filter -> fan-out of a mono source or one filter
    channel per source channel.
    """

    def __init__(self, filters, n_streams: int, c_in: int = 1, *, block: int | None = None,
                 pull_hint: int | None = None, max_pull: int | None = None, filter_of_stream=None,
                 device: int = 0, mixdown_input: bool = False, single_filter_dims: bool = False,
                 tail_block: int | None = None):
        h = np.asarray(filters, dtype=np.float32)
        if single_filter_dims:  # (L,) or (L, C_f)
            h = h[None] if h.ndim >= 1 else h
        if h.ndim == 2:  # (F, L)
            h = h[:, :, None]
        if h.ndim != 3:
            raise ValueError(f"filters must be (F, L) or (F, L, C_f), got shape {h.shape}")
        F, L, c_f = h.shape
        if L < 1:
            raise ValueError("ConvolvePE filter must be non-empty")
        c_x = 1 if mixdown_input else int(c_in)
        if c_f == 1:
            c_out = c_x
        elif c_x == 1 or c_x == c_f:
            c_out = c_f
        else:
            raise ValueError(
                f"ConvolvePE filter channels ({c_f}) must match src channels ({c_x}), "
                f"or be mono, or be multi-channel with a mono source."
            )
        B = int(block) if block else choose_block(L, pull_hint)
        self.n_streams, self.c_in, self.c_out, self.c_f = int(n_streams), int(c_in), int(c_out), int(c_f)
        self.filter_len, self.n_filters, self.block = int(L), int(F), B
        self.tail_block = int(tail_block) if tail_block and L > int(tail_block) else 0
        # two-level partitioning (extension): the first tail_block taps at block B, the rest at block tail_block
        self.partitions = -(-(self.tail_block or L) // B)
        self.tail_partitions = -(-(L - self.tail_block) // self.tail_block) if self.tail_block else 0
        self.max_pull = int(max_pull) if max_pull else max(8 * B, 4096)
        self.device = int(device)
        _lib.require_device()
        cfg = BankConfig(device=self.device, n_streams=self.n_streams, c_in=self.c_in, c_out=self.c_out,
                         filter_len=L, filter_channels=c_f, n_filters=F, block=B, max_pull=self.max_pull,
                         flags=_lib.PGX_FLAG_MIXDOWN_INPUT if mixdown_input else 0, tail_block=self.tail_block)
        h_planar = np.ascontiguousarray(np.transpose(h, (0, 2, 1)))  # [F][C_f][L]
        fmap = None
        if filter_of_stream is not None:
            fmap = np.ascontiguousarray(filter_of_stream, dtype=np.int32)
            if fmap.shape != (self.n_streams,):
                raise ValueError("filter_of_stream must have one entry per stream")
        self._h = C.c_void_p()
        check(lib().pgx_bank_create(C.byref(self._h), C.byref(cfg),
                                    h_planar.ctypes.data_as(C.POINTER(C.c_float)),
                                    fmap.ctypes.data_as(C.POINTER(C.c_int32)) if fmap is not None else None))
        self.sources = None
        self._inflight = {}
        self._submit_layouts = {}
        self._submit_ticket = C.c_int64(-1)
        self._submit_fn = lib().pgx_bank_submit

    # -- lifetime ------------------------------------------------------------
    def close(self) -> None:
        res = getattr(self, "_resident", None)
        if res is not None:
            res.close()
            self._resident = None
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().pgx_bank_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> BankInfo:
        out = BankInfo()
        check(lib().pgx_bank_get_info(self._h, C.byref(out)))
        return out

    # -- state ---------------------------------------------------------------
    def reset(self, streams=None) -> None:
        """History := 0 for all streams (and re-anchor the block grid), or for the listed ones."""
        if streams is None:
            check(lib().pgx_bank_reset(self._h, None, 0))
        else:
            ids = np.ascontiguousarray(streams, dtype=np.int32)
            check(lib().pgx_bank_reset(self._h, ids.ctypes.data_as(C.POINTER(C.c_int32)), int(ids.size)))

    def load_filter(self, filter_index: int, h) -> None:
        """Replace one resident filter: h is (L,) or (L, C_f) float32."""
        h = np.asarray(h, dtype=np.float32)
        if h.ndim == 1:
            h = h[:, None]
        if h.shape != (self.filter_len, self.c_f):
            raise ValueError(f"filter must be ({self.filter_len}, {self.c_f}), got {h.shape}")
        hp = np.ascontiguousarray(h.T)
        check(lib().pgx_bank_load_filter(self._h, int(filter_index), hp.ctypes.data_as(C.POINTER(C.c_float))))

    def set_filter_map(self, filter_of_stream) -> None:
        fmap = np.ascontiguousarray(filter_of_stream, dtype=np.int32)
        if fmap.shape != (self.n_streams,):
            raise ValueError("filter_of_stream must have one entry per stream")
        check(lib().pgx_bank_set_filter_map(self._h, fmap.ctypes.data_as(C.POINTER(C.c_int32))))

    def set_output_gains(self, wet: float = 1.0, dry: float = 0.0) -> None:
        """Fused output stage of the following pulls: y = dry * x + wet * (x * h) in float32 (ReverbPE's
        GainPE/GainPE/MixPE tail, reverb_pe.py:82-95)."""
        check(lib().pgx_bank_set_output_gains(self._h, float(np.float32(wet)), float(np.float32(dry))))

    def use_filter_map_device(self, ptr: int | None) -> None:
        """Point the bank at an int32 [n_streams] map already resident on the device (None: its own)."""
        check(lib().pgx_bank_use_filter_map_device(self._h, C.c_void_p(ptr) if ptr else None))

    # -- pulls (host buffers; copies inside) ------------------------------------
    def _chunks(self, n):
        pos = 0
        while pos < n:
            d = min(self.max_pull, n - pos)
            yield pos, d
            pos += d

    def process(self, x: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """x (N, C_in, n) float32 -> y (N, C_out, n) float32."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 3 or x.shape[0] != self.n_streams or x.shape[1] != self.c_in:
            raise ValueError(f"x must be ({self.n_streams}, {self.c_in}, n), got {x.shape}")
        n = x.shape[2]
        y = out if out is not None else np.empty((self.n_streams, self.c_out, n), dtype=np.float32)
        if n <= self.max_pull:
            check(lib().pgx_bank_process(self._h, _lib.f32_ptr(x), Layout(self.c_in * n, n, 1),
                                         _lib.f32_ptr(y), Layout(self.c_out * n, n, 1), n))
            return y
        for pos, d in self._chunks(n):
            xc = np.ascontiguousarray(x[:, :, pos:pos + d])
            yc = np.empty((self.n_streams, self.c_out, d), dtype=np.float32)
            check(lib().pgx_bank_process(self._h, _lib.f32_ptr(xc), Layout(self.c_in * d, d, 1),
                                         _lib.f32_ptr(yc), Layout(self.c_out * d, d, 1), d))
            y[:, :, pos:pos + d] = yc
        return y

    def process_mix(self, x: np.ndarray) -> np.ndarray:
        """x (N, C_in, n) -> fused MixPE sum over streams (C_out, n)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 3 or x.shape[0] != self.n_streams or x.shape[1] != self.c_in:
            raise ValueError(f"x must be ({self.n_streams}, {self.c_in}, n), got {x.shape}")
        n = x.shape[2]
        y = np.empty((self.c_out, n), dtype=np.float32)
        for pos, d in self._chunks(n):
            xc = x if d == n else np.ascontiguousarray(x[:, :, pos:pos + d])
            yc = y if d == n else np.empty((self.c_out, d), dtype=np.float32)
            check(lib().pgx_bank_process_mix(self._h, _lib.f32_ptr(xc), Layout(self.c_in * d, d, 1),
                                             _lib.f32_ptr(yc), Layout(0, d, 1), d))
            if d != n:
                y[:, pos:pos + d] = yc
        return y

    def process_interleaved(self, x: np.ndarray, pcm16_out: bool = False) -> np.ndarray:
        """Single-stream Snippet layout: x (n, C_in) -> y (n, C_out). Used by the PE shims.  An int16 ``x`` is
        raw PCM_16 converted on the device (int16/32768); ``pcm16_out`` delivers y as int16 PCM converted on the
        device (clip(rint(y*32768))) -- WAV staging, SURVEY.md §8f rank 4."""
        if self.n_streams != 1:
            raise ValueError("process_interleaved is for single-stream banks")
        pcm_in = x.dtype == np.int16
        x = np.ascontiguousarray(x, dtype=np.int16 if pcm_in else np.float32)
        n = x.shape[0]
        y = np.empty((n, self.c_out), dtype=np.int16 if pcm16_out else np.float32)
        flags = (_lib.PGX_PULL_X_PCM16 if pcm_in else 0) | (_lib.PGX_PULL_Y_PCM16 if pcm16_out else 0)
        for pos, d in self._chunks(n):
            xc, yc = x[pos:pos + d], y[pos:pos + d]  # row slices of C-contiguous arrays stay dense
            if flags == 0:
                check(lib().pgx_bank_process(self._h, _lib.f32_ptr(xc), Layout(0, 1, self.c_in),
                                             _lib.f32_ptr(yc), Layout(0, 1, self.c_out), d))
            else:
                check(lib().pgx_bank_pull(self._h, _lib.f32_ptr(xc), Layout(0, 1, self.c_in), _lib.f32_ptr(yc),
                                          Layout(0, 1, self.c_out), d, flags))
        return y

    # -- pipelined host-buffer pulls (several in flight; copies overlap the kernels) ------------
    def attach_comm(self, comm) -> None:
        """Mix pulls submitted with ``reduce=True`` are summed over the ranks of ``comm`` (a ``dist.MixComm``)
        onto its root; None detaches."""
        self._comm = comm  # keep it alive as long as the bank uses it
        check(lib().pgx_bank_attach_comm(self._h, comm._h if comm is not None else None))

    def submit(self, x: np.ndarray, out: np.ndarray, *, mix: bool = False, reduce: bool = False) -> int:
        """Enqueue one pull: x (N, C_in, n) -> out (N, C_out, n), or (C_out, n) when ``mix``.  Returns a
        ticket for ``wait``.  x and out must be C-contiguous float32 (pinned for real overlap: ``PinnedArray``)
        and must not be touched until the wait returns; at most `info().submit_depth` pulls are in flight (3; 8 on a time-tiled bank)."""
        if x.dtype not in (np.float32, np.int16) or out.dtype not in (np.float32, np.int16) \
                or not x.flags.c_contiguous or not out.flags.c_contiguous:
            raise ValueError("submit needs C-contiguous float32 (or int16 PCM) arrays")
        if x.ndim != 3 or x.shape[0] != self.n_streams or x.shape[1] != self.c_in:
            raise ValueError(f"x must be ({self.n_streams}, {self.c_in}, n), got {x.shape}")
        n = x.shape[2]
        want = (self.c_out, n) if mix else (self.n_streams, self.c_out, n)
        if out.shape != want:
            raise ValueError(f"out must be {want}, got {out.shape}")
        key = (n, mix)
        lay = self._submit_layouts.get(key)
        if lay is None:   # (the two layout structs of a pull shape are built once: this call is on the per-pull path)
            lay = self._submit_layouts[key] = (Layout(self.c_in * n, n, 1), Layout(0 if mix else self.c_out * n, n, 1))
        tk = self._submit_ticket
        check(self._submit_fn(self._h, _lib.f32_ptr(x), lay[0], _lib.f32_ptr(out), lay[1], n,
                              (1 if mix else 0) | (_lib.PGX_PULL_REDUCE if reduce else 0)
                              | (_lib.PGX_PULL_X_PCM16 if x.dtype == np.int16 else 0)
                              | (_lib.PGX_PULL_Y_PCM16 if out.dtype == np.int16 else 0), C.byref(tk)))
        t = int(tk.value)
        self._inflight[t] = (x, out)  # keep the buffers alive until waited
        return t

    def wait(self, ticket: int) -> None:
        check(lib().pgx_bank_wait(self._h, int(ticket)))
        infl = self._inflight
        for t in [t for t in infl if t <= ticket]:
            del infl[t]

    # -- device-resident INPUT (a DeviceBlock from osc_pe / VoiceBank), host output ------------------
    @property
    def stream_ptr(self) -> int:
        """The bank's critical CUDA stream (cudaStream_t as int): producers of device input enqueue on it."""
        return int(lib().pgx_bank_stream(self._h) or 0)

    def process_device_block(self, blk, *, mix: bool = False, interleaved: bool = False) -> np.ndarray:
        """One pull whose input is already in HBM (produced on ``stream_ptr``): no H2D copy.  Returns host
        (N, C_out, n), the fused mix (C_out, n), or -- single stream, ``interleaved`` -- Snippet layout (n, C_out)."""
        n = int(blk.duration)
        if blk.n_streams != self.n_streams or blk.channels != self.c_in:
            raise ValueError(f"device block is {blk.n_streams} streams x {blk.channels} channels, "
                             f"bank expects {self.n_streams} x {self.c_in}")
        if n > self.max_pull:
            raise ValueError(f"device block of {n} samples exceeds max_pull={self.max_pull}")
        if interleaved:
            if self.n_streams != 1 or mix:
                raise ValueError("interleaved output is for single-stream banks")
            y, yl = np.empty((n, self.c_out), dtype=np.float32), Layout(0, 1, self.c_out)
        elif mix:
            y, yl = np.empty((self.c_out, n), dtype=np.float32), Layout(0, n, 1)
        else:
            y, yl = np.empty((self.n_streams, self.c_out, n), dtype=np.float32), Layout(self.c_out * n, n, 1)
        flags = (_lib.PGX_PULL_MIX if mix else 0) | _lib.PGX_PULL_X_DEVICE
        check(lib().pgx_bank_pull(self._h, blk.ptr, blk.layout, y.ctypes.data, yl, n, flags))
        return y

    # -- device-resident pulls (pointers from torch / cuda-python; not synchronised) ------------
    def process_device(self, x_ptr: int, y_ptr: int, n: int, *, mix: bool = False, cuda_stream: int = 0,
                       input_resident: bool = False, x_layout: Layout | None = None,
                       y_layout: Layout | None = None, reduce: bool = False) -> None:
        """Enqueue one pull on device buffers (planar by default). ``input_resident``: x is already complete
        in memory, so its ingest may overlap the output stage of pulls queued earlier."""
        xl = x_layout or Layout(self.c_in * n, n, 1)
        yl = y_layout or Layout(0 if mix else self.c_out * n, n, 1)
        flags = (1 if mix else 0) | (2 if input_resident else 0) | (_lib.PGX_PULL_REDUCE if reduce else 0)
        check(lib().pgx_bank_process_device(self._h, C.c_void_p(x_ptr), xl, C.c_void_p(y_ptr), yl, int(n),
                                            flags, C.c_void_p(cuda_stream)))

    def synchronize(self) -> None:
        check(lib().pgx_bank_synchronize(self._h))

    # -- measurement ---------------------------------------------------------------
    def profile_begin(self) -> None:
        check(lib().pgx_bank_profile_begin(self._h))

    def profile_end(self):
        """-> _lib.Profile(ms_r2c, ms_mac, ms_c2r, steps): summed per-kernel CUDA-event durations."""
        out = _lib.Profile()
        check(lib().pgx_bank_profile_end(self._h, C.byref(out)))
        return out

    # -- batched renderer support ------------------------------------------------
    def attach_sources(self, sources, delays=None, gains=None, extents=None, silent_filter=None) -> None:
        """N host PEs feeding the N streams; enables ``render`` for BankRenderer.  ``delays`` (integer samples:
        the source is pulled that much earlier, delay_pe.py:153-160) and ``gains`` (float32, applied to the
        source samples, gain_pe.py:123-125) fold per-stream DelayPE / GainPE wrappers into the pull."""
        sources = list(sources)
        if len(sources) != self.n_streams:
            raise ValueError("need one source PE per stream")
        self.sources = sources
        self._src_delays = [0] * len(sources) if delays is None else [int(d) for d in delays]
        self._src_gains = [None] * len(sources) if gains is None else [None if g is None else np.float32(g) for g in gains]
        # ``extents`` (one Extent per stream, in output time): MixPE's gating (mix_pe.py:81-85) -- a stream whose
        # extent does not meet the request is not rendered: zero input AND the all-zero filter ``silent_filter``
        # (its carried history must not ring on), and it starts a new run when it comes back
        self._gate = None
        if extents is not None:
            lo = np.array([-np.inf if e.start is None else e.start for e in extents], dtype=np.float64)
            hi = np.array([np.inf if e.end is None else e.end for e in extents], dtype=np.float64)
            empty = np.array([e.is_empty() for e in extents], dtype=bool)
            self._gate = (lo, hi, empty)
            self._was_active = np.ones(len(sources), dtype=bool)
            self._silent = None if silent_filter is None else int(silent_filter)
            self._own_map = np.arange(len(sources), dtype=np.int32)
            self._cur_map = self._own_map.copy()
        self._resident = None
        from .resident import ResidentSources  # plain in-memory sources are uploaded once and stay in HBM
        if ResidentSources.eligible(sources, self._src_delays, self.c_in, False):
            self._resident = ResidentSources(sources, self._src_delays, self._src_gains, self.c_in, False,
                                             device=self.device)
        self._pos = None
        self.mix_output = False

    def attach_device_source(self, source) -> None:
        """One device-resident producer for all N streams (``VoiceBank`` or anything with
        ``device_block(start, duration, cuda_stream=...)`` yielding N x C_in planar samples): the lockstep
        pull then has no host input at all."""
        self.sources = source
        self._pos = None
        self.mix_output = False
        self._device_source = True
        self._resident = None

    def render(self, start: int, duration: int) -> np.ndarray:
        """One lockstep pull of every attached source: (N, C_out, n), or (C_out, n) when mix_output."""
        if self.sources is None:
            raise RuntimeError("no sources attached")
        if self._pos is None or start != self._pos:
            self.reset()  # non-contiguous pull: history := 0 (convolve_pe.py:255-256)
        if getattr(self, "_resident", None) is not None and not getattr(self, "_device_source", False):
            outs, pos = [], 0
            while pos < duration:
                d = min(self.max_pull, self._resident.max_pull, duration - pos)
                outs.append(self.process_device_block(self._resident.device_block(start + pos, d), mix=self.mix_output))
                pos += d
            self._pos = start + duration
            return outs[0] if len(outs) == 1 else np.concatenate(outs, axis=-1)
        if getattr(self, "_device_source", False):
            outs, pos = [], 0
            while pos < duration:
                d = min(self.max_pull, getattr(self.sources, "max_pull", self.max_pull), duration - pos)
                blk = self.sources.device_block(start + pos, d, cuda_stream=self.stream_ptr)
                outs.append(self.process_device_block(blk, mix=self.mix_output))
                pos += d
            self._pos = start + duration
            return outs[0] if len(outs) == 1 else np.concatenate(outs, axis=-1)
        active = None
        if getattr(self, "_gate", None) is not None:
            lo, hi, empty = self._gate
            active = ~empty & (lo < start + duration) & (hi > start)        # Extent.intersects
            back = np.flatnonzero(active & ~self._was_active)
            if back.size and self._pos == start:                           # (a full reset just happened otherwise)
                self.reset(back)
            self._was_active = active
            if self._silent is not None:
                sel = np.where(active, self._own_map, np.int32(self._silent)).astype(np.int32)
                if not np.array_equal(sel, self._cur_map):
                    self.set_filter_map(sel)
                    self._cur_map = sel
        x = np.empty((self.n_streams, self.c_in, duration), dtype=np.float32)
        for s, pe in enumerate(self.sources):
            if active is not None and not active[s]:
                x[s] = 0.0
                continue
            data = pe.render(start - self._src_delays[s], duration).data
            if self._src_gains[s] is not None:
                data = data * self._src_gains[s]
            x[s] = data.T
        self._pos = start + duration
        return self.process_mix(x) if self.mix_output else self.process(x)
