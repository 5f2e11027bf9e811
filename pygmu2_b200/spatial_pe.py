"""
SpatialPE and its SpatialMethods -- drop-in for src/pygmu2/spatial_pe.py.

Only ``SpatialHRTF`` is convolution and runs on the GPU (a one-stream ConvolveBank with
input mix-down and mono -> stereo fan-out; the batched, mix-fused form is
``HrtfMixBank``).  ``SpatialAdapter`` / ``SpatialLinear`` / ``SpatialConstantPower`` are
stateless per-sample gain maps (not convolution, "don't break"): they are
host numpy, same arithmetic as the reference (spatial_pe.py:68-290).
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from . import kemar
from .bank import ConvolveBank, choose_block
from .core import Extent, ProcessingElement, Snippet, handle_error


class SpatialMethod(ABC):
    """spatial_pe.py:34-65."""

    @property
    @abstractmethod
    def output_channels(self) -> int: ...

    @abstractmethod
    def render(self, source_snippet: Snippet, start: int, duration: int, sample_rate: int) -> np.ndarray: ...

    def inputs(self) -> list:
        return []


class SpatialAdapter(SpatialMethod):
    """M -> N channel conversion without spatialisation (spatial_pe.py:68-147)."""

    def __init__(self, channels: int):
        if channels < 1:
            raise ValueError(f"SpatialAdapter: channels must be >= 1 (got {channels})")
        self._channels = int(channels)

    @property
    def output_channels(self) -> int:
        return self._channels

    def render(self, source_snippet, start, duration, sample_rate):
        x = source_snippet.data
        m, n = source_snippet.channels, self._channels
        if m == n:
            return x
        out = np.zeros((duration, n), dtype=np.float32)
        if m == 1:
            out[:, :] = x[:, 0:1]
        elif n == 1:
            out[:, 0] = np.mean(x, axis=1)
        elif m == 2 and n == 4:
            out[:, 0:2] = x
            out[:, 2] = out[:, 3] = np.mean(x, axis=1)
        elif m == 4 and n == 2:
            out[:, :] = x[:, 0:2]
        else:
            k = min(m, n)
            out[:, :k] = x[:, :k]
            if n > m:
                out[:, m:] = x[:, m - 1:m]
            else:
                out[:, n - 1] += np.mean(x[:, n:], axis=1)
        return out

    def __repr__(self):
        return f"SpatialAdapter(channels={self._channels})"


class _Pan(SpatialMethod):
    def __init__(self, azimuth):
        self.azimuth = azimuth

    @property
    def output_channels(self) -> int:
        return 2

    def inputs(self) -> list:
        return [self.azimuth] if isinstance(self.azimuth, ProcessingElement) else []

    def _azimuths(self, start, duration):
        if isinstance(self.azimuth, ProcessingElement):
            az = self.azimuth.render(start, duration).data[:, 0]
        else:
            az = np.full(duration, float(self.azimuth), dtype=np.float32)
        return np.clip(az, -90.0, 90.0)

    def _gains(self, az):
        raise NotImplementedError

    def render(self, source_snippet, start, duration, sample_rate):
        mono = np.mean(source_snippet.data, axis=1, keepdims=True)
        gl, gr = self._gains(self._azimuths(start, duration))
        out = np.zeros((duration, 2), dtype=np.float32)
        out[:, 0] = mono[:, 0] * gl
        out[:, 1] = mono[:, 0] * gr
        return out

    def __repr__(self):
        a = f"{self.azimuth:.1f}" if isinstance(self.azimuth, (int, float)) else type(self.azimuth).__name__
        return f"{type(self).__name__}(azimuth={a})"


class SpatialLinear(_Pan):
    """L = 1 - pan, R = pan, pan = (clip(az) + 90)/180 (spatial_pe.py:150-218)."""

    def _gains(self, az):
        pan = (az + 90.0) / 180.0
        return 1.0 - pan, pan


class SpatialConstantPower(_Pan):
    """L = cos, R = sin of (clip(az) + 90)/2 degrees (spatial_pe.py:221-290)."""

    def _gains(self, az):
        ang = np.deg2rad((az + 90.0) / 2.0)
        return np.cos(ang), np.sin(ang)


class SpatialHRTF(SpatialMethod):
    """KEMAR binaural spatialisation on the GPU (reference spatial_pe.py:293-521).

    ``azimuth`` / ``elevation`` are public floats and may be mutated between pulls; as in the
    reference the IR pair is re-resolved on every render (:446-449,472) and the IR of the
    *current* pull is applied to the whole carried history (hard switch at pull boundaries).
    """

    KEMAR_HRTF_ENTRIES = kemar.KEMAR_HRTF_ENTRIES

    @staticmethod
    def hrtf_filename_for(azimuth: float, elevation: float) -> str:
        return kemar.KEMAR_HRTF_ENTRIES[kemar.nearest_index(azimuth, elevation)][2]

    def __init__(self, azimuth, elevation=0.0, *, block_size: int | None = None, device: int = 0):
        if isinstance(azimuth, ProcessingElement) or isinstance(elevation, ProcessingElement):
            raise ValueError(
                "SpatialHRTF: azimuth and elevation must be static (float or int). "
                "Dynamic values would switch impulse responses during rendering and cause discontinuities."
            )
        self._watchers = []            # (az_array, el_array, index, dirty_cell) of the fused banks this method feeds: an
                                       # assignment lands in the bank's direction arrays and raises its dirty flag
        self.azimuth = float(azimuth)
        self.elevation = float(elevation)
        self._block_size, self._device = block_size, int(device)
        self._bank = None
        self._bank_src_ch = None
        self._loaded = None  # (table index, swapped) resident in the bank
        self._last_render_end = None
        self._warned_sr_mismatch = False

    # azimuth / elevation stay plain public attributes to the user (spatial_pe.py:434-435: mutate them between
    # pulls to move the source); as properties they can tell a fused bank that its cached selection is stale
    azimuth = property(lambda self: self._azimuth)
    elevation = property(lambda self: self._elevation)

    @azimuth.setter
    def azimuth(self, value):
        self._azimuth = value
        for az, _, i, cell in self._watchers:
            az[i] = value
            cell[0] = True

    @elevation.setter
    def elevation(self, value):
        self._elevation = value
        for _, el, i, cell in self._watchers:
            el[i] = value
            cell[0] = True

    @property
    def output_channels(self) -> int:
        return 2

    def _select(self):
        idx = kemar.nearest_index(self.azimuth, self.elevation)
        return idx, self.azimuth < 0  # left side: same file, ears swapped (spatial_pe.py:486-489)

    @staticmethod
    def _ir_pair(idx: int, swapped: bool) -> np.ndarray:
        table, _ = kemar.load_table()
        ir = table[idx]
        if ir.ndim != 2 or ir.shape[1] != 2:
            raise ValueError(f"SpatialHRTF: expected stereo IR, got shape {ir.shape}")
        return ir[:, ::-1] if swapped else ir

    def render(self, source_snippet: Snippet, start: int, duration: int, sample_rate: int) -> np.ndarray:
        _, ir_sr = kemar.load_table()
        if sample_rate != ir_sr and not self._warned_sr_mismatch:
            handle_error(
                f"SpatialHRTF: IR sample rate is {ir_sr} Hz but source is {sample_rate} Hz. "
                "Proceeding without resampling.",
                fatal=False,
            )
            self._warned_sr_mismatch = True
        x = source_snippet.data
        sel = self._select()
        if self._bank is None or self._bank_src_ch != x.shape[1]:
            ir = self._ir_pair(*sel)
            block = self._block_size or choose_block(ir.shape[0], duration)
            self._bank = ConvolveBank(ir, 1, x.shape[1], block=block, device=self._device,
                                      mixdown_input=True, single_filter_dims=True)
            self._bank_src_ch = x.shape[1]
            self._loaded = sel
            self._last_render_end = None
        elif sel != self._loaded:
            self._bank.load_filter(0, self._ir_pair(*sel))
            self._loaded = sel
        if self._last_render_end is None or start != self._last_render_end:
            self._bank.reset()  # spatial_pe.py:461-463
        y = self._bank.process_interleaved(x)
        self._last_render_end = start + duration
        return y

    def __repr__(self):
        return f"SpatialHRTF(azimuth={self.azimuth:.1f}, elevation={self.elevation:.1f})"


class SpatialPE(ProcessingElement):
    """M-channel source -> N-channel output through a SpatialMethod (spatial_pe.py:524-671)."""

    def __init__(self, source: ProcessingElement, *, method: SpatialMethod):
        if method is None:
            raise ValueError("SpatialPE: method is required")
        self._source, self._method = source, method

    source = property(lambda self: self._source)
    method = property(lambda self: self._method)

    def inputs(self) -> list:
        return [self._source, *self._method.inputs()]

    def is_pure(self) -> bool:
        return True  # as the reference declares (spatial_pe.py:629-632), HRTF history notwithstanding

    def channel_count(self):
        return self._method.output_channels

    def _compute_extent(self) -> Extent:
        return self._source.extent()

    def _render(self, start: int, duration: int) -> Snippet:
        snip = self._source.render(start, duration)
        sr = self._source.sample_rate
        if sr is None:
            handle_error("SpatialPE: sample_rate is unknown; proceeding without a configured rate.", fatal=False)
            sr = 0
        return Snippet(start, self._method.render(snip, start, duration, sr))

    def __repr__(self):
        return f"SpatialPE(source={self._source.__class__.__name__}, method={self._method})"
