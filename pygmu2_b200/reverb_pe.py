"""
ReverbPE -- convolution reverb with a wet/dry mix: the principal in-tree caller of
ConvolvePE + MixPE (reference src/pygmu2/reverb_pe.py:27-138).

    out = MixPE(GainPE(CachePE(src), 1 - mix), GainPE(ConvolvePE(CachePE(src), ir), mix / ir_energy))

The source is pulled once per render thanks to CachePE (pinned by reference
tests/test_convolve_pe.py:214-221).  With a constant ``mix`` and a source whose channel count
equals the wet path's, the whole composite is ONE device call per pull: the two gains and the
add are the epilogue of the inverse-FFT kernel (``pgx_bank_set_output_gains``), rounded to
float32 exactly where GainPE / MixPE round.  A PE-valued ``mix`` keeps the composite graph
(per-sample gains on the host, the device ConvolvePE for the wet path, the device MixPE sum).
"""
from __future__ import annotations

from .convolve_pe import ConvolvePE
from .core import Extent, ProcessingElement, Snippet
from .mix_pe import MixPE
from .sources import CachePE, ConstantPE, GainPE


def f32_to_pcm16_host(x):
    from .wav_pe import f32_to_pcm16
    return f32_to_pcm16(x)


class ReverbPE(ProcessingElement):
    def __init__(self, source: ProcessingElement, ir: ProcessingElement, mix: float = 0.5, *,
                 normalize_ir: bool = True, fft_size: int | None = None):
        self._source = CachePE(source)  # one pull feeds both the dry and the wet path
        self._ir, self._mix = ir, mix
        self._normalize_ir, self._fft_size = bool(normalize_ir), fft_size
        if isinstance(mix, ProcessingElement):
            mix_ch = mix.channel_count()
            if mix_ch is not None and int(mix_ch) != 1:
                raise ValueError(f"mix PE must be mono, got {mix_ch} channels")
        else:
            mix = float(mix)
            if not (0.0 <= mix <= 1.0):
                raise ValueError(f"mix must be in [0.0, 1.0], got {mix}")
            self._mix = mix
        self._ir_energy = ConvolvePE.ir_energy_norm(ir) if self._normalize_ir else 1.0
        self._wet_stream = ConvolvePE(self._source, ir, fft_size=fft_size)
        if isinstance(mix, ProcessingElement):  # reverb_pe.py:83-87
            dry_gain = MixPE(ConstantPE(1.0), GainPE(mix, gain=-1.0), fuse=False)
            wet_gain = GainPE(mix, gain=1.0 / self._ir_energy) if self._normalize_ir else mix
        else:                                   # reverb_pe.py:88-92
            dry_gain = 1.0 - mix
            wet_gain = mix / self._ir_energy if self._normalize_ir else mix
        self._out = MixPE(GainPE(self._source, gain=dry_gain), GainPE(self._wet_stream, gain=wet_gain), fuse=False)
        # fused form: the same arithmetic as the epilogue of the wet path's inverse FFT
        self._fused = False
        if not isinstance(mix, ProcessingElement):
            src_ch, out_ch = source.channel_count(), self._wet_stream.channel_count()
            if src_ch is not None and src_ch == out_ch:
                self._wet_stream._out_gains = (wet_gain, dry_gain)
                self._fused = True

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
        if self._fused:
            return self._wet_stream.render(start, duration)
        return self._out.render(start, duration)

    def render_pcm16_out(self, start: int, duration: int):
        """int16 PCM straight off the device (fused form only); None tells the writer to use ``render``."""
        if not self._fused:
            return f32_to_pcm16_host(self.render(start, duration).data)
        return self._wet_stream.render_pcm16_out(start, duration)

    def __repr__(self):
        return (f"ReverbPE(source={self._source.__class__.__name__}, ir={self._ir.__class__.__name__}, "
                f"mix={self._mix}, normalize_ir={self._normalize_ir}, fft_size={self._fft_size})")
