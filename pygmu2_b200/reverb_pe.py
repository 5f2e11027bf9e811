"""
ReverbPE -- convolution reverb with a wet/dry mix: the principal in-tree caller of
ConvolvePE + MixPE (reference src/pygmu2/reverb_pe.py:27-138).

    out = MixPE(GainPE(CachePE(src), 1 - mix), GainPE(ConvolvePE(CachePE(src), ir), mix / ir_energy))

The wet path is the device ConvolvePE, the final sum the device MixPE; the source is pulled
once per render thanks to CachePE (pinned by reference tests/test_convolve_pe.py:214-221).
Only a constant ``mix`` is supported here (a PE-valued mix needs the PE-gain form of GainPE,
which is outside the hot path: gain_pe.py:104-121).
"""
from __future__ import annotations

from .convolve_pe import ConvolvePE
from .core import Extent, ProcessingElement, Snippet
from .mix_pe import MixPE
from .sources import CachePE, GainPE


class ReverbPE(ProcessingElement):
    def __init__(self, source: ProcessingElement, ir: ProcessingElement, mix: float = 0.5, *,
                 normalize_ir: bool = True, fft_size: int | None = None):
        if isinstance(mix, ProcessingElement):
            raise NotImplementedError("pygmu2_b200.ReverbPE supports a constant mix only")
        mix = float(mix)
        if not (0.0 <= mix <= 1.0):
            raise ValueError(f"mix must be in [0.0, 1.0], got {mix}")
        self._source = CachePE(source)  # one pull feeds both the dry and the wet path
        self._ir, self._mix = ir, mix
        self._normalize_ir, self._fft_size = bool(normalize_ir), fft_size
        self._ir_energy = ConvolvePE.ir_energy_norm(ir) if self._normalize_ir else 1.0
        self._wet_stream = ConvolvePE(self._source, ir, fft_size=fft_size)
        wet_gain = mix / self._ir_energy if self._normalize_ir else mix
        self._out = MixPE(GainPE(self._source, gain=1.0 - mix), GainPE(self._wet_stream, gain=wet_gain), fuse=False)

    source = property(lambda self: self._source)
    ir = property(lambda self: self._ir)
    mix = property(lambda self: self._mix)
    ir_energy = property(lambda self: self._ir_energy)

    def inputs(self) -> list:
        return [self._out]

    def is_pure(self) -> bool:
        return False

    def channel_count(self):
        return self._out.channel_count()

    def _compute_extent(self) -> Extent:
        return self._out.extent()

    def _render(self, start: int, duration: int) -> Snippet:
        return self._out.render(start, duration)

    def __repr__(self):
        return (f"ReverbPE(source={self._source.__class__.__name__}, ir={self._ir.__class__.__name__}, "
                f"mix={self._mix}, normalize_ir={self._normalize_ir}, fft_size={self._fft_size})")
