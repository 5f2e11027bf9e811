"""
Host-side input vehicles around the hot path: the few trivial pygmu2 PEs that
the reference's own hot-path tests, examples and the BASELINE configs use to
feed ConvolvePE / SpatialPE / MixPE.  They are *not* accelerated (SURVEY.md §2b:
out of scope) and exist so graphs can be built on a machine without pygmu2.

ArrayPE   array_pe.py:17-133     ConstantPE constant_pe.py:15-72
GainPE    gain_pe.py:60-155      DelayPE    delay_pe.py:19-231 (integer: folded into fused mixes; float / PE: interpolated)
CropPE    crop_pe.py:18-96       CachePE    cache_pe.py:17-84
"""
from __future__ import annotations

import enum

import numpy as np

from .core import Extent, ExtendMode, ProcessingElement, Snippet, SourcePE


class ArrayPE(SourcePE):
    def __init__(self, data, extend_mode: ExtendMode = ExtendMode.ZERO):
        arr = np.asarray(data, dtype=np.float32)
        if arr.ndim == 1:
            arr = arr.reshape(-1, 1)
        elif arr.ndim > 2:
            raise ValueError(f"ArrayPE data must be 1D or 2D, got {arr.ndim}D")
        if arr.shape[0] == 0:
            raise ValueError("ArrayPE data cannot be empty")
        self._data = arr
        self._extend_mode = extend_mode

    data = property(lambda self: self._data)

    def channel_count(self) -> int:
        return self._data.shape[1]

    def _compute_extent(self) -> Extent:
        return Extent(0, self._data.shape[0])

    def _render(self, start: int, duration: int) -> Snippet:
        n, ch = self._data.shape
        out = np.zeros((duration, ch), dtype=np.float32)
        lo, hi = max(0, start), min(n, start + duration)
        if lo < hi:
            out[lo - start:hi - start] = self._data[lo:hi]
        mode = self._extend_mode
        if mode in (ExtendMode.HOLD_FIRST, ExtendMode.HOLD_BOTH) and start < 0:
            out[:min(duration, -start)] = self._data[0]
        if mode in (ExtendMode.HOLD_LAST, ExtendMode.HOLD_BOTH) and start + duration > n:
            first = max(0, n - start)
            if first < duration:
                out[first:] = self._data[-1]
        return Snippet(start, out)

    def __repr__(self):
        return f"ArrayPE(shape={self._data.shape})"


class ConstantPE(SourcePE):
    def __init__(self, value: float, channels: int = 1):
        self._value, self._channels = value, channels

    value = property(lambda self: self._value)

    def channel_count(self) -> int:
        return self._channels

    def _render(self, start: int, duration: int) -> Snippet:
        return Snippet(start, np.full((duration, self._channels), self._value, dtype=np.float32))

    def __repr__(self):
        return f"ConstantPE(value={self._value}, channels={self._channels})"


class _Unary(ProcessingElement):
    def __init__(self, source: ProcessingElement):
        self._source = source

    source = property(lambda self: self._source)

    def inputs(self) -> list:
        return [self._source]

    def is_pure(self) -> bool:
        return True

    def channel_count(self):
        return self._source.channel_count()

    def _compute_extent(self) -> Extent:
        return self._source.extent()


class GainPE(_Unary):
    """gain_pe.py:60-155: constant gain (float32 multiply) or a PE of per-sample gains."""

    def __init__(self, source: ProcessingElement, gain=1.0):
        super().__init__(source)
        self._gain = gain
        self._gain_is_pe = isinstance(gain, ProcessingElement)

    gain = property(lambda self: self._gain)

    def inputs(self) -> list:
        return [self._source, self._gain] if self._gain_is_pe else [self._source]

    def _compute_extent(self) -> Extent:
        e = self._source.extent()
        return e.intersection(self._gain.extent()) if self._gain_is_pe else e

    def _render(self, start: int, duration: int) -> Snippet:
        x = self._source.render(start, duration).data
        if not self._gain_is_pe:
            return Snippet(start, x * np.float32(self._gain))
        g = self._gain.render(start, duration).data.astype(np.float32, copy=False)
        if g.shape[1] == 1 and x.shape[1] > 1:
            g = np.tile(g, (1, x.shape[1]))
        return Snippet(start, x * g)


class InterpolationMode(enum.Enum):
    """wavetable_pe.py:19-22 (the reference keeps it there; DelayPE is its user on this path)."""
    LINEAR = "linear"
    CUBIC = "cubic"


def _interpolate(source: ProcessingElement, out_start: int, where: np.ndarray, mode, silent: np.ndarray | None) -> Snippet:
    """Sample ``source`` at the fractional positions ``where`` (interpolated_lookup.py:89-145): ONE render of the
    window that covers every tap, taps clamped to that window, float64 weights on the float32 samples, rounded to
    float32 at the end; positions flagged in ``silent`` (outside a finite source extent) give 0."""
    where = np.asarray(where, dtype=np.float64).reshape(-1)
    if where.size == 0:
        return Snippet.from_zeros(out_start, 0, source.channel_count() or 1)
    cubic = str(getattr(mode, "value", mode)).lower() == "cubic"
    reach = 2 if cubic else 1                                   # taps below / above the bracketing pair
    lo = int(np.floor(where.min())) - (reach - 1)
    hi = int(np.ceil(where.max())) + reach
    window = source.render(lo, hi - lo).data
    base = np.floor(where).astype(np.int64)
    f = (where - base).reshape(-1, 1)
    last = window.shape[0] - 1

    def tap(offset):
        return window[np.clip(base - lo + offset, 0, last)]

    if cubic:   # Catmull-Rom through the four neighbours, in the reference's grouping (its float32 sub-sums included)
        a, b_, c, d = tap(-1), tap(0), tap(1), tap(2)
        f2 = f * f
        f3 = f2 * f
        y = 0.5 * ((2.0 * b_) + (-a + c) * f + (2.0 * a - 5.0 * b_ + 4.0 * c - d) * f2 + (-a + 3.0 * b_ - 3.0 * c + d) * f3)
    else:
        y = (1.0 - f) * tap(0) + f * tap(1)
    if silent is not None and silent.any():
        y = np.array(y, copy=True)
        y[silent] = 0.0
    return Snippet(out_start, y.astype(np.float32, copy=False))


class DelayPE(_Unary):
    """delay_pe.py:19-231.  An integer delay pulls the source earlier (and is what a fused MixPE folds into its bank);
    a fractional or PE-valued delay reads the source at ``t - delay[t]`` with linear or cubic interpolation.  The
    interpolated modes are host-side input vehicles (numpy), like PE-valued GainPE."""

    def __init__(self, source: ProcessingElement, delay, interpolation=InterpolationMode.LINEAR):
        super().__init__(source)
        self._interpolation = interpolation
        if isinstance(delay, ProcessingElement):
            self._mode, self._delay = "pe", delay
        elif isinstance(delay, float) and not delay.is_integer():
            self._mode, self._delay = "float", delay
        else:
            self._mode, self._delay = "int", int(delay)

    delay = property(lambda self: self._delay)
    mode = property(lambda self: self._mode)
    interpolation = property(lambda self: self._interpolation)

    def inputs(self) -> list:
        return [self._source, self._delay] if self._mode == "pe" else [self._source]

    def _compute_extent(self) -> Extent:
        e = self._source.extent()
        if self._mode == "pe":                       # defined where both the source and the control are
            return e.intersection(self._delay.extent())
        lo = None if e.start is None else e.start + self._delay
        hi = None if e.end is None else e.end + self._delay
        if self._mode == "float":                    # extents are integer: round outwards
            lo = None if lo is None else int(np.floor(lo))
            hi = None if hi is None else int(np.ceil(hi))
        return Extent(lo, hi)

    def _render(self, start: int, duration: int) -> Snippet:
        if self._mode == "int":
            return Snippet(start, self._source.render(start - self._delay, duration).data)
        t = np.arange(start, start + duration, dtype=np.float64)
        if self._mode == "float":
            where = t - self._delay
        else:                                         # the control's first channel, per sample
            where = t - self._delay.render(start, duration).data[:, 0].astype(np.float64)
        e = self._source.extent()
        silent = None
        if e.start is not None and e.end is not None:
            silent = (where < e.start) | (where >= e.end)
        return _interpolate(self._source, start, where, self._interpolation, silent)


class CropPE(_Unary):
    """Zero outside [start, start+duration) of the source's timeline."""

    def __init__(self, source: ProcessingElement, start: int, duration: int):
        super().__init__(source)
        self._crop = Extent(int(start), int(start) + int(duration))

    def _compute_extent(self) -> Extent:
        return self._source.extent().intersection(self._crop)

    def _render(self, start: int, duration: int) -> Snippet:
        ch = self.channel_count() or 1
        out = np.zeros((duration, ch), dtype=np.float32)
        ov = self.extent().intersection(Extent(start, start + duration))
        if not ov.is_empty() and ov.start is not None and ov.end is not None and ov.end > ov.start:
            seg = self._source.render(ov.start, ov.end - ov.start).data
            out[ov.start - start:ov.end - start] = seg
        return Snippet(start, out)


class CachePE(_Unary):
    """Memoises the last (start, duration) pull so two sinks cost one source render."""

    def __init__(self, source: ProcessingElement):
        super().__init__(source)
        self._key = None
        self._snip = None

    def _reset_state(self) -> None:
        self._key = self._snip = None

    _on_start = _on_stop = _reset_state

    def _render(self, start: int, duration: int) -> Snippet:
        if self._snip is not None and self._key == (start, duration):
            return self._snip
        self._snip = self._source.render(start, duration)
        self._key = (start, duration)
        return self._snip
