"""
ResidentSources -- in-memory sources kept in HBM (the ArrayPE case).

When every source of a fused bank is a plain ``ArrayPE`` (zero outside its data, array_pe.py:94-111) the
per-pull host work of the reference -- N ``render()`` calls, N Snippets, one (N, C, n) gather and its H2D
copy -- is replaced by ONE upload: the arrays are stacked into a device buffer ``[N][C][T]`` with each
source's integer delay baked into its position, its constant gain applied and, where the bank wants a mono
input, its channels averaged (all float32, exactly what the per-pull path computes).  A pull is then just a
pointer into that buffer (``PGX_PULL_X_DEVICE``): no host samples at all.

The buffer is a SNAPSHOT of ``ArrayPE.data`` taken when the bank adopts the sources: the reference's ArrayPE
re-reads the caller's array on every render (array_pe.py:46,94-111), so a caller that mutates the array in place
after the first pull must rebuild the graph here (rendered data is meant to be immutable,
processing_element.py:109-111).  ``pgx_device_upload`` / ``pgx_device_zero`` return only when the bytes are visible to every
CUDA stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Layout, check, lib
from .core import ExtendMode
from .osc_pe import DeviceBlock
from .sources import ArrayPE

MARGIN = 1 << 16           # zero samples either side: any pull up to this long may hang over the ends
MAX_BYTES = 8 << 30        # above this the sources stay on the host


class ResidentSources:
    @staticmethod
    def eligible(sources, delays, c_in: int, host_mixdown: bool) -> bool:
        if not sources or not all(type(p) is ArrayPE and p._extend_mode is ExtendMode.ZERO for p in sources):
            return False
        if not host_mixdown and any(p.channel_count() != c_in for p in sources):
            return False
        lo = min(int(d) for d in delays)
        hi = max(p.data.shape[0] + int(d) for p, d in zip(sources, delays))
        return len(sources) * c_in * (hi - lo + 2 * MARGIN) * 4 <= MAX_BYTES

    def __init__(self, sources, delays, gains, c_in: int, host_mixdown: bool, device: int = 0):
        self.n, self.c_in, self.device = len(sources), int(c_in), int(device)
        self.t0 = min(int(d) for d in delays)
        self.t1 = max(p.data.shape[0] + int(d) for p, d in zip(sources, delays))
        self.T = self.t1 - self.t0 + 2 * MARGIN
        self.max_pull = MARGIN
        nbytes = self.n * self.c_in * self.T * 4
        _lib.require_device()
        self._ptr = C.c_void_p()
        check(lib().pgx_device_alloc(self.device, nbytes, C.byref(self._ptr)))
        check(lib().pgx_device_zero(self.device, self._ptr, nbytes))
        self._base = int(self._ptr.value)
        row = self.T * 4
        for s, (pe, d, g) in enumerate(zip(sources, delays, gains)):
            data = pe.data                                                   # (n_s, C_s) float32
            if host_mixdown:
                data = np.mean(data, axis=1, keepdims=True).astype(np.float32)   # spatial_pe.py:483
            if g is not None:
                data = data * np.float32(g)                                  # gain_pe.py:123-125
            planar = np.ascontiguousarray(data.T, dtype=np.float32)          # (C, n_s)
            off = (int(d) - self.t0 + MARGIN) * 4
            for c in range(self.c_in):
                dst = self._ptr.value + (s * self.c_in + c) * row + off
                check(lib().pgx_device_upload(self.device, C.c_void_p(dst), planar[c].ctypes.data, planar[c].nbytes))

    def device_block(self, start: int, duration: int, cuda_stream: int = 0) -> DeviceBlock | None:
        """Samples [start, start+duration) of every source, already in HBM.  Requests wholly outside the data
        are served from the zero margins."""
        if duration > MARGIN:
            return None
        pos = int(start) - self.t0 + MARGIN
        pos = min(max(pos, 0), self.T - duration)          # beyond either end everything is zero anyway
        lay = getattr(self, "_layout", None)
        if lay is None:
            lay = self._layout = Layout(self.c_in * self.T, self.T, 1)
        return DeviceBlock(self._base + pos * 4, lay, self.n, self.c_in, int(duration))

    def close(self) -> None:
        if getattr(self, "_ptr", None) is not None and self._ptr.value:
            lib().pgx_device_free(self.device, self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
