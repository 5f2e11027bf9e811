"""
Host-side mirror of pygmu2's plugin contract: the value types and the
ProcessingElement (PE) protocol the hot path sits behind.

pygmu2 has no registry or FFI: a PE is "plugged in" by subclassing
``ProcessingElement`` (reference src/pygmu2/processing_element.py:28-294) and
returning ``Snippet`` objects (snippet.py:14-108) described by ``Extent``
intervals (extent.py:21-205), under a global sample rate and error policy
(config.py:18-109).  The reference is pure Python and is not installed on the
GPU box, so this module restates that contract -- same names, argument meaning
and error behaviour -- for the device-backed PEs in this package.  Nothing here
touches the GPU.  ``INTEGRATION.md`` shows how the same PEs bind to a real
pygmu2 install instead of this mirror.
"""
from __future__ import annotations

import logging
import threading
import time
from abc import ABC, abstractmethod
from enum import Enum

import numpy as np

log = logging.getLogger("pygmu2_b200")

# ---------------------------------------------------------------------------
# config.py:18-109 -- global sample rate and error policy
_state = {"sample_rate": None}


def set_sample_rate(rate: int) -> None:
    _state["sample_rate"] = int(rate)


def get_sample_rate():
    return _state["sample_rate"]


class ErrorMode(Enum):
    STRICT = "strict"
    LENIENT = "lenient"


_state["error_mode"] = ErrorMode.STRICT


def set_error_mode(mode: ErrorMode) -> None:
    _state["error_mode"] = mode


def get_error_mode() -> ErrorMode:
    return _state["error_mode"]


def handle_error(message, fatal=False, error_mode=None, exception_class=RuntimeError) -> bool:
    """Raise in STRICT mode or when ``fatal``; otherwise warn and return True (config.py:68-109)."""
    mode = _state["error_mode"] if error_mode is None else error_mode
    if fatal or mode is ErrorMode.STRICT:
        raise exception_class(message)
    log.warning(message)
    return True


# ---------------------------------------------------------------------------
# extent.py:21-205
class ExtendMode(Enum):
    ZERO = "zero"
    HOLD_FIRST = "hold_first"
    HOLD_LAST = "hold_last"
    HOLD_BOTH = "hold_both"


class Extent:
    """Half-open sample interval [start, end); ``None`` is infinite on that side."""

    __slots__ = ("_start", "_end")

    def __init__(self, start=None, end=None):
        if start is not None and end is not None and start > end:
            raise ValueError(f"start ({start}) must be less than or equal to end ({end})")
        self._start, self._end = start, end

    start = property(lambda self: self._start)
    end = property(lambda self: self._end)

    @property
    def duration(self):
        if self._start is None or self._end is None:
            return None
        return self._end - self._start

    def is_empty(self) -> bool:
        return self._start is not None and self._start == self._end

    def contains(self, i: int) -> bool:
        return (self._start is None or i >= self._start) and (self._end is None or i < self._end)

    def spans(self, start: int, duration: int) -> bool:
        if duration <= 0:
            return True
        return (self._start is None or start >= self._start) and \
               (self._end is None or start + duration <= self._end)

    def intersects(self, other: "Extent") -> bool:
        if self.is_empty() or other.is_empty():
            return False  # extent.py:106-108: empty never overlaps
        if self._end is not None and other._start is not None and self._end <= other._start:
            return False
        if other._end is not None and self._start is not None and other._end <= self._start:
            return False
        return True

    def intersection(self, other: "Extent") -> "Extent":
        if self.is_empty():
            return Extent(self._start, self._start)
        if other.is_empty():
            return Extent(other._start, other._start)
        lo = [v for v in (self._start, other._start) if v is not None]
        hi = [v for v in (self._end, other._end) if v is not None]
        s = max(lo) if lo else None
        e = min(hi) if hi else None
        if s is not None and e is not None and s > e:
            return Extent(s, s)
        return Extent(s, e)

    def union(self, other: "Extent") -> "Extent":
        if self.is_empty():
            return other
        if other.is_empty():
            return self
        s = None if (self._start is None or other._start is None) else min(self._start, other._start)
        e = None if (self._end is None or other._end is None) else max(self._end, other._end)
        return Extent(s, e)

    def __eq__(self, other):
        if not isinstance(other, Extent):
            return NotImplemented
        return self._start == other._start and self._end == other._end

    def __bool__(self):
        return not self.is_empty()

    def __repr__(self):
        s = "-∞" if self._start is None else str(self._start)
        e = "+∞" if self._end is None else str(self._end)
        return f"Extent({s}, {e})"


# ---------------------------------------------------------------------------
# snippet.py:14-108
class Snippet:
    """(samples, channels) float32 buffer with a start index. Treat ``data`` as immutable."""

    __slots__ = ("_start", "_data")

    def __init__(self, start: int, data):
        if data.ndim == 1:
            data = data.reshape(-1, 1)
        elif data.ndim != 2:
            raise ValueError(f"data must be 1D or 2D, got {data.ndim}D")
        if data.dtype != np.float32:
            data = data.astype(np.float32, copy=False)
        self._start, self._data = start, data

    start = property(lambda self: self._start)
    data = property(lambda self: self._data)
    duration = property(lambda self: self._data.shape[0])
    channels = property(lambda self: self._data.shape[1])
    end = property(lambda self: self._start + self._data.shape[0])

    @classmethod
    def from_zeros(cls, start: int, duration: int, channels: int = 1) -> "Snippet":
        return cls(start, np.zeros((duration, channels), dtype=np.float32))

    def __eq__(self, other):
        if not isinstance(other, Snippet):
            return NotImplemented
        return (self._start == other._start and self._data.shape == other._data.shape
                and np.allclose(self._data, other._data))

    def __repr__(self):
        return f"Snippet(start={self._start}, duration={self.duration}, channels={self.channels})"


# ---------------------------------------------------------------------------
# diagnostics.py:23-128 -- thread-local pull counts / per-class wall timings
class _Diag(threading.local):
    enabled = False
    pulls = None
    nanos = None


_diag = _Diag()


def enable_diagnostics(on: bool = True) -> None:
    _diag.enabled = bool(on)
    _diag.pulls, _diag.nanos = {}, {}


def diagnostics_report() -> dict:
    """{class name: (pulls, total_ms)} since enable_diagnostics()."""
    if not _diag.pulls:
        return {}
    return {k: (v, _diag.nanos.get(k, 0) / 1e6) for k, v in _diag.pulls.items()}


# ---------------------------------------------------------------------------
# processing_element.py:28-294
class ProcessingElement(ABC):
    _sample_rate = None
    _cached_extent = None

    def __new__(cls, *args, **kwargs):
        # processing_element.py:51-65: the global sample rate must exist before any PE
        rate = get_sample_rate()
        if rate is None:
            raise RuntimeError(
                "Global sample_rate is required but not set. "
                "Call pygmu2.set_sample_rate(rate) before constructing PEs."
            )
        obj = super().__new__(cls)
        obj._sample_rate = rate
        return obj

    @property
    def sample_rate(self):
        if self._sample_rate is not None:
            return self._sample_rate
        for p in self.inputs():
            if p.sample_rate is not None:
                return p.sample_rate
        return None

    def render(self, start: int, duration: int) -> Snippet:
        """Always returns exactly ``duration`` samples from ``start`` (processing_element.py:95-135)."""
        if duration < 0:
            raise ValueError(f"duration must be >= 0, got {duration}")
        if _diag.enabled:
            name = type(self).__name__
            _diag.pulls[name] = _diag.pulls.get(name, 0) + 1
        if duration == 0:
            ch = self.channel_count()
            return Snippet.from_zeros(start, 0, int(ch) if ch is not None else 1)
        if _diag.enabled:
            t0 = time.perf_counter_ns()
            out = self._render(start, duration)
            name = type(self).__name__
            _diag.nanos[name] = _diag.nanos.get(name, 0) + time.perf_counter_ns() - t0
            return out
        return self._render(start, duration)

    @abstractmethod
    def _render(self, start: int, duration: int) -> Snippet: ...

    def extent(self) -> Extent:
        if self._cached_extent is None:
            self._cached_extent = self._compute_extent()
        return self._cached_extent

    def _compute_extent(self) -> Extent:
        return Extent(None, None)

    @abstractmethod
    def inputs(self) -> list: ...

    def is_pure(self) -> bool:
        return False

    def channel_count(self):
        return None

    def required_input_channels(self):
        return None

    def resolve_channel_count(self, input_channel_counts):
        if input_channel_counts:
            return input_channel_counts[0]
        raise ValueError(f"{type(self).__name__} has no inputs but channel_count() is None")

    def on_start(self) -> None:
        hook = getattr(self, "_on_start", None)
        if hook is not None:
            hook()

    def on_stop(self) -> None:
        hook = getattr(self, "_on_stop", None)
        if hook is not None:
            hook()

    def reset_state(self) -> None:
        hook = getattr(self, "_reset_state", None)
        if hook is not None:
            hook()


class SourcePE(ProcessingElement):
    """A PE with no inputs; pure by default (source_pe.py:14-52)."""

    def inputs(self) -> list:
        return []

    def is_pure(self) -> bool:
        return True

    def channel_count(self):
        return 1
