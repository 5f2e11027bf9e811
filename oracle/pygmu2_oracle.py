"""
CPU oracle for the ConvolvePE / SpatialHRTF / MixPE path of rdpoor/pygmu2.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pygmu2_b200/`` imports this file.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the timed
CPU baseline, never as the product path.

What it is: a numpy/scipy restatement of the reference's *algorithm* (same
schedule, same dtypes, same third-party calls), function by function, each
citing the reference file:line it follows (paths relative to the reference
tree, ``src/pygmu2/...``).  The arithmetic itself lives in third-party code
that is not under the reference tree: ``numpy.fft.rfft/irfft`` (numpy 2.3.5
pinned in the reference's uv.lock:805-806; C++ pocketfft, float64) and
``scipy.signal.fftconvolve`` (scipy 1.17.0 pinned, uv.lock:1312-1313; float32
preserved).  The oracle calls the same functions, so it is the same arithmetic.

Parity pinned: yes.  ``tests/test_oracle_golden.py`` checks every function here
against (a) the reference's own known-answer tests (tests/test_convolve_pe.py,
tests/test_mix_pe.py, tests/test_spatial_pe.py of the reference, restated in
``tests/``) and (b) golden vectors in ``tests/golden/*.npz`` produced by
importing the real Python reference in the build container
(``oracle/gen_golden.py``; the reference cannot travel to the GPU box).
``SpatialHRTF.render`` has no numeric test in the reference, so for it the
pin is (b) only.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "next_pow2",
    "OracleConvolve",
    "OracleHRTF",
    "oracle_mix",
    "ir_energy_norm",
    "kemar_entries",
    "hrtf_nearest_index",
    "direct_convolve_f64",
]


def next_pow2(n: int) -> int:
    """convolve_pe.py:34-38."""
    n = int(n)
    if n <= 1:
        return 1
    return 1 << (n - 1).bit_length()


class OracleConvolve:
    """Single-partition overlap-save streaming convolution, float64.

    Restates ConvolvePE._ensure_filter_prepared (convolve_pe.py:185-248) and
    ConvolvePE._render (convolve_pe.py:250-342) on plain arrays:

    * ``h``  : (L,) or (L, C_f) float32 filter, rendered once (:198)
    * ``x``  : (n, C_src) float32 chunks fed to ``render`` one pull at a time
    * output : (n, C_out) float32 (:342)

    ``render(x, contiguous=False)`` is what a non-contiguous ``start`` does in
    the reference: the carried tail is zeroed (:255-256).
    """

    def __init__(self, h, src_channels: int, fft_size: int | None = None):
        h = np.asarray(h, dtype=np.float32)
        if h.ndim == 1:
            h = h.reshape(-1, 1)
        h = h.astype(np.float64, copy=False)  # :198
        L = h.shape[0]
        if L < 1:
            raise ValueError("ConvolvePE filter must be non-empty")  # :195
        filt_ch = h.shape[1]
        src_ch = int(src_channels)
        # channel resolution, :207-223
        if filt_ch == 1:
            out_ch = src_ch
        elif src_ch == 1:
            out_ch = filt_ch
        elif filt_ch == src_ch:
            out_ch = filt_ch
        else:
            raise ValueError(
                f"ConvolvePE filter channels ({filt_ch}) must match src channels ({src_ch}), "
                f"or be mono, or be multi-channel with a mono source."
            )
        # FFT size, :226-231
        if fft_size is None:
            fft_size = next_pow2(max(2048, L))
        if fft_size < L:
            raise ValueError(f"fft_size ({fft_size}) must be >= filter length ({L})")
        self.nfft = int(fft_size)
        # :236-239
        if filt_ch == 1:
            self.H = np.fft.rfft(h[:, 0], n=self.nfft)
        else:
            self.H = np.fft.rfft(h, n=self.nfft, axis=0)
        self.L = L
        self.out_ch = out_ch
        self.tail = np.zeros((max(L - 1, 0), out_ch), dtype=np.float64)  # :245-248

    def reset(self) -> None:
        """_reset_state / non-contiguous pull, :146-154,255-256."""
        self.tail[:] = 0.0

    def render(self, x, contiguous: bool = True) -> np.ndarray:
        if not contiguous:
            self.tail[:] = 0.0  # :255-256
        nfft, L = self.nfft, self.L
        tail_len = L - 1
        hop = nfft - tail_len  # :258-263
        if hop < 1:
            raise ValueError(f"fft_size ({nfft}) too small for filter length ({L})")
        x = np.asarray(x, dtype=np.float32)
        if x.ndim == 1:
            x = x.reshape(-1, 1)
        x = x.astype(np.float64, copy=False)  # :266
        duration, src_ch = x.shape
        out_ch = src_ch if self.H.ndim == 1 else self.H.shape[1]  # :273-278
        if self.tail.shape[1] != out_ch:
            self.tail = np.zeros((tail_len, out_ch), dtype=np.float64)  # :281-283
        y = np.zeros((duration, out_ch), dtype=np.float64)
        pos = 0
        while pos < duration:  # :289-339
            n = min(hop, duration - pos)
            seg = x[pos:pos + n, :]
            if out_ch == src_ch:
                seg_o = seg
            elif src_ch == 1:
                seg_o = np.repeat(seg, out_ch, axis=1)  # fan-out, :303-305
            else:
                seg_o = seg[:, :out_ch]
            blk = np.zeros((nfft, out_ch), dtype=np.float64)  # :294
            if tail_len > 0:
                blk[:tail_len, :] = self.tail
            blk[tail_len:tail_len + n, :] = seg_o
            X = np.fft.rfft(blk, axis=0)  # :313
            Y = X * (self.H.reshape(-1, 1) if self.H.ndim == 1 else self.H)  # :314-317
            yb = np.fft.irfft(Y, n=nfft, axis=0)  # :318
            y[pos:pos + n, :] = yb[tail_len:tail_len + n, :]  # :321-322
            if tail_len > 0:  # :325-336
                if n >= tail_len:
                    self.tail = blk[n:tail_len + n, :].copy()
                else:
                    self.tail = np.vstack([self.tail, seg_o])[-tail_len:, :].copy()
            pos += n
        return y.astype(np.float32)  # :342


def direct_convolve_f64(x, h) -> np.ndarray:
    """Independent check of the result contract: the exact
    linear convolution of float32-quantised x and h evaluated in float64 by the
    direct sum, rounded to float32.  x: (n,), h: (L,) -> (n+L-1,)."""
    x = np.asarray(x, dtype=np.float32).astype(np.float64)
    h = np.asarray(h, dtype=np.float32).astype(np.float64)
    return np.convolve(x, h, mode="full").astype(np.float32)


def ir_energy_norm(h) -> float:
    """ConvolvePE.ir_energy_norm on a rendered finite filter, convolve_pe.py:86-108."""
    data = np.asarray(h, dtype=np.float32)
    energy_norm = np.sqrt(np.sum(data.astype(np.float64) ** 2))
    return float(energy_norm) if energy_norm > 1e-10 else 1.0


# ---------------------------------------------------------------------------
# KEMAR compact set geometry.  The reference lists the 368 (elev, az, file)
# triples literally (spatial_pe.py:318-393); they follow the published MIT
# KEMAR measurement grid (Gardner & Martin 1994): elevations -40..90 in steps
# of 10 degrees, n_az equally spaced azimuths around the full circle per
# elevation, azimuth rounded to the nearest degree, right hemisphere kept.
_KEMAR_AZ_COUNTS = {-40: 56, -30: 60, -20: 72, -10: 72, 0: 72, 10: 72, 20: 72,
                    30: 60, 40: 56, 50: 45, 60: 36, 70: 24, 80: 12, 90: 1}


def kemar_entries() -> list[tuple[int, int, str]]:
    """(elev, az, filename) in the reference's table order (spatial_pe.py:318-393)."""
    out = []
    for elev in sorted(_KEMAR_AZ_COUNTS):
        n_az = _KEMAR_AZ_COUNTS[elev]
        for i in range(n_az):
            az = int(np.floor(i * 360.0 / n_az + 0.5))
            if az > 180:
                break
            out.append((elev, az, f"H{elev}e{az:03d}a.wav"))
    return out


def hrtf_nearest_index(azimuth: float, elevation: float, entries=None) -> int:
    """Index into ``kemar_entries()`` picked by SpatialHRTF.hrtf_filename_for
    (spatial_pe.py:395-426): az clamped to min(180,|az|), squared Euclidean
    distance in (elev, az), first minimum in table order (python ``min``)."""
    if entries is None:
        entries = kemar_entries()
    az = min(180.0, abs(float(azimuth)))
    elev = float(elevation)
    best_i, best_d = 0, None
    for i, e in enumerate(entries):
        d = (e[0] - elev) ** 2 + (e[1] - az) ** 2
        if best_d is None or d < best_d:
            best_i, best_d = i, d
    return best_i


class OracleHRTF:
    """SpatialHRTF.render restated (spatial_pe.py:465-518), float32 throughout.

    ``table``: (368, Lh, 2) float32 IR pairs in ``kemar_entries()`` order
    (what ``sf.read(..., dtype='float32')`` returns per file, :452).
    ``azimuth``/``elevation`` are public and may be mutated between pulls;
    the IR is re-resolved on every render (:446-449,472).
    """

    def __init__(self, table, azimuth: float, elevation: float = 0.0, entries=None):
        self.table = np.asarray(table, dtype=np.float32)
        self.entries = entries if entries is not None else kemar_entries()
        self.azimuth = float(azimuth)
        self.elevation = float(elevation)
        self.tail = None
        self.last_end = None

    def render(self, src, start: int) -> np.ndarray:
        from scipy.signal import fftconvolve  # same call as :503-504

        src = np.asarray(src, dtype=np.float32)
        if src.ndim == 1:
            src = src.reshape(-1, 1)
        duration = src.shape[0]
        ir = self.table[hrtf_nearest_index(self.azimuth, self.elevation, self.entries)]
        mono = np.mean(src, axis=1).astype(np.float32, copy=False)  # :483
        left_ir, right_ir = ir[:, 0], ir[:, 1]
        if self.azimuth < 0:  # :486-489
            left_ir, right_ir = right_ir, left_ir
        tail_len = max(left_ir.shape[0] - 1, 0)
        if self.last_end is None or start != self.last_end:  # :461-463
            self.tail = None
        if tail_len == 0:
            x = mono
        else:
            if self.tail is None or self.tail.shape[0] != tail_len:
                self.tail = np.zeros(tail_len, dtype=np.float32)
            x = np.concatenate([self.tail, mono])  # :499-501
        left = fftconvolve(x, left_ir, mode="full")  # :503
        right = fftconvolve(x, right_ir, mode="full")  # :504
        out_l = left[tail_len:tail_len + duration]  # :506-511
        out_r = right[tail_len:tail_len + duration]
        if tail_len > 0:
            self.tail = x[-tail_len:]  # :514
        self.last_end = start + duration
        return np.column_stack([out_l, out_r]).astype(np.float32, copy=False)  # :517


def oracle_mix(inputs) -> np.ndarray:
    """MixPE._render's arithmetic (mix_pe.py:92-94): copy of the first rendered
    input, then sequential float32 ``+=`` in constructor order."""
    inputs = [np.asarray(a, dtype=np.float32) for a in inputs]
    result = inputs[0].copy()
    for a in inputs[1:]:
        result += a
    return result
