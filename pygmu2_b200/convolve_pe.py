"""
ConvolvePE -- drop-in for pygmu2's streaming FIR convolution PE, on the GPU.

Same constructor, properties, channel rules, extent, exceptions and statefulness as
the reference (src/pygmu2/convolve_pe.py:41-348); ``render(start, duration)`` returns a
host Snippet of (duration, channels) float32.  What changes is underneath: instead of
one float64 numpy rfft/irfft pair of size ``fft_size`` per hop (:289-339), the PE keeps
a device-resident uniformly partitioned overlap-save state (``ConvolveBank`` with one
stream) and advances it with three sm_100a kernels per block step.  ``fft_size`` keeps
its meaning as a schedule hint only -- it never changed the result in the reference
 -- but is validated and reported exactly as before.
"""
from __future__ import annotations

import numpy as np

from .bank import ConvolveBank, choose_block, next_pow2
from .core import Extent, ProcessingElement, Snippet, log


class ConvolvePE(ProcessingElement):
    """Streaming convolution y = x * h (filter finite, starting at 0).

    Args:
        src: source PE
        fir: filter PE, extent Extent(0, L)
        fft_size: reference-compatible hint; must be >= L when given
        block_size: (extension) partition size B of the device path; default picks
            next_pow2(first pull) clamped to [16, 4096]
        tail_block: (extension) two-level partitioning for a long filter pulled in small blocks: the first
            tail_block taps at block_size, the rest at tail_block (same result, far less delay-line traffic)
        device: (extension) CUDA device ordinal, default 0
        speculate: (extension) render the next block of a constant-parameter device source (oscillator / voice mix)
            right behind every pull, so that it is ready when a PACED caller (a real-time callback) asks for it; the
            samples are the same (state snapshot + rollback on a miss).  Default: the PGX_SPECULATE environment
            variable, else off -- in a back-to-back pull loop it only adds work to a full time line
    """

    def __init__(self, src: ProcessingElement, fir: ProcessingElement, *, fft_size: int | None = None,
                 block_size: int | None = None, tail_block: int | None = None, device: int = 0,
                 speculate: bool | None = None):
        self._src = src
        self._fir = fir
        self._fft_size = int(fft_size) if fft_size is not None else None
        self._block_size = int(block_size) if block_size is not None else None
        self._tail_block = int(tail_block) if tail_block else None
        self._device = int(device)
        self._fir_len = None
        self._bank = None
        self._last_render_end = None
        self._out_gains = None  # (wet, dry) set by ReverbPE: fused GainPE/GainPE/MixPE tail
        self._spec = None       # ((start, duration), DeviceBlock) of the source block rendered ahead of its pull
        import os
        # off by default: it pays when the caller is paced (a real-time callback: the source block is ready when the pull
        # arrives; C5 paced at real time 45.9 -> 38.7 us per pull) and costs a few us per pull in a back-to-back loop,
        # where the GPU's time line is full either way
        self._speculate = (os.environ.get("PGX_SPECULATE", "0") == "1") if speculate is None else bool(speculate)

    src = property(lambda self: self._src)
    fir = property(lambda self: self._fir)
    fft_size = property(lambda self: self._fft_size)

    def inputs(self) -> list:
        return [self._src, self._fir]

    @staticmethod
    def ir_energy_norm(filter_pe: ProcessingElement) -> float:
        """sqrt(sum h^2) over all channels in float64; 1.0 if unbounded or ~0 (convolve_pe.py:86-108)."""
        ext = filter_pe.extent()
        if ext.start is None or ext.end is None:
            return 1.0
        data = filter_pe.render(ext.start, ext.end - ext.start).data
        norm = float(np.sqrt(np.sum(data.astype(np.float64) ** 2)))
        return norm if norm > 1e-10 else 1.0

    def is_pure(self) -> bool:
        return False  # carries overlap state

    def channel_count(self):
        # convolve_pe.py:114-144
        src_ch, filt_ch = self._src.channel_count(), self._fir.channel_count()
        if src_ch is None and filt_ch is None:
            return None
        if src_ch is None:
            return filt_ch
        if filt_ch is None or int(filt_ch) == 1:
            return src_ch
        if int(src_ch) == 1:
            return int(filt_ch)
        return src_ch

    def _on_start(self) -> None:
        self._reset_state()

    def _on_stop(self) -> None:
        self._reset_state()

    def _drop_speculation(self) -> None:
        if self._spec is not None:
            self._spec = None
            rb = getattr(self._src, "rollback_speculation", None)
            if rb is not None and self._bank is not None:
                rb(self._bank.stream_ptr)

    def _reset_state(self) -> None:
        self._drop_speculation()
        # The reference drops _tail here but keeps _H, which makes a restart assert
        # (convolve_pe.py:152-154,186-187,252; SURVEY.md §7 "bugs not to copy").  Here the
        # prepared filter stays resident and the history is cleared on the next pull.
        self._last_render_end = None

    def _compute_extent(self) -> Extent:
        # convolve_pe.py:156-183
        src_ext, filt_ext = self._src.extent(), self._fir.extent()
        if filt_ext.start is not None and filt_ext.start != 0:
            raise ValueError(f"ConvolvePE filter extent must start at 0, got {filt_ext}")
        if filt_ext.start is None:
            raise ValueError(f"ConvolvePE filter extent must be finite and start at 0, got {filt_ext}")
        if filt_ext.end is None:
            raise ValueError(f"ConvolvePE filter extent must be finite, got {filt_ext}")
        filt_len = int(filt_ext.end - filt_ext.start)
        if filt_len < 1:
            return Extent(0, 0)
        if src_ext.end is None:
            return Extent(src_ext.start, None)
        return Extent(src_ext.start, int(src_ext.end + (filt_len - 1)))

    def _ensure_filter_prepared(self, pull_hint: int) -> None:
        if self._bank is not None:
            return
        filt_ext = self._fir.extent()
        if filt_ext.start != 0 or filt_ext.end is None:
            raise ValueError(f"ConvolvePE filter must have extent Extent(0, N), got {filt_ext}")
        filt_len = int(filt_ext.end)
        if filt_len < 1:
            raise ValueError("ConvolvePE filter must be non-empty")
        h = self._fir.render(0, filt_len).data  # rendered once, float32 (convolve_pe.py:198)
        if h.ndim != 2 or h.shape[0] != filt_len:
            raise ValueError(f"ConvolvePE filter returned invalid shape {getattr(h, 'shape', None)}")
        src_ch = self._src.channel_count()
        if src_ch is None:
            src_ch = self._src.render(0, 1).channels  # convolve_pe.py:203-205
        if self._fft_size is None:
            self._fft_size = next_pow2(max(2048, filt_len))  # convolve_pe.py:226-229
        if self._fft_size < filt_len:
            raise ValueError(f"fft_size ({self._fft_size}) must be >= filter length ({filt_len})")
        block = self._block_size or choose_block(filt_len, pull_hint)
        # channel-rule violations raise ValueError inside ConvolveBank (convolve_pe.py:219-223)
        self._bank = ConvolveBank(h, 1, int(src_ch), block=block, device=self._device, single_filter_dims=True,
                                  tail_block=self._tail_block)
        if self._out_gains is not None:
            self._bank.set_output_gains(*self._out_gains)
        self._fir_len = filt_len
        # a plain in-memory source (ArrayPE, zero outside its data) is uploaded once and stays in HBM: a pull is then a
        # pointer into that buffer instead of a host render + H2D copy per pull (same samples; see resident.py)
        from .resident import ResidentSources
        from .sources import CachePE
        self._resident = None
        inner = self._src
        while type(inner) is CachePE:          # a pure memo of the source (ReverbPE wraps its source in one): same samples
            inner = inner.source
        if ResidentSources.eligible([inner], [0], int(src_ch), False):
            self._resident = ResidentSources([inner], [0], [None], int(src_ch), False, device=self._device)

    def render_pcm16_out(self, start: int, duration: int) -> np.ndarray:
        """One pull delivered as (duration, channels) int16 PCM, converted on the device (WAV staging; what
        WavWriterPE would write for this pull).  Same state and contiguity rules as ``render``."""
        if duration < 0:
            raise ValueError(f"duration must be non-negative, got {duration}")
        if duration == 0:
            return np.zeros((0, self.channel_count() or 1), dtype=np.int16)
        return self._render(start, duration, pcm16_out=True)

    def _render(self, start: int, duration: int, pcm16_out: bool = False):
        self._ensure_filter_prepared(duration)
        if self._spec is not None and (pcm16_out or self._spec[0] != (start, duration)):
            self._drop_speculation()     # the block rendered ahead is not the one being asked for: undo its state advance
        if duration >= 32 * self._bank.block and not getattr(self, "_warned_block", False):
            self._warned_block = True   # B is fixed at the first pull; correct, but every pull is many block steps
            log.warning("ConvolvePE: pull of %d samples on a bank partitioned at B=%d (chosen at the first pull); "
                        "pass block_size= to pin a larger partition", duration, self._bank.block)
        if self._last_render_end is None or start != self._last_render_end:
            self._bank.reset()  # non-contiguous pull: prior samples are zeros (convolve_pe.py:255-256)
        res = None if pcm16_out else getattr(self, "_resident", None)
        if res is not None and duration <= min(self._bank.max_pull, res.max_pull):
            y = self._bank.process_device_block(res.device_block(start, duration), interleaved=True)
            self._last_render_end = start + duration
            return Snippet(start, y)
        dev = None if pcm16_out else getattr(self._src, "device_block", None)
        if dev is not None:  # device-resident source: its samples never visit the host
            y = self._render_from_device(dev, start, duration)
            if y is not None:
                self._last_render_end = start + duration
                return Snippet(start, y)
        raw = getattr(self._src, "render_pcm16", None)   # WavReaderPE: raw int16 frames, converted in HBM
        x = raw(start, duration) if raw is not None else self._src.render(start, duration).data
        if x.ndim != 2:
            raise ValueError(f"ConvolvePE src returned invalid shape {getattr(x, 'shape', None)}")
        if x.shape[1] != self._bank.c_in:
            raise ValueError(f"ConvolvePE src returned {x.shape[1]} channels, prepared for {self._bank.c_in}")
        y = self._bank.process_interleaved(x, pcm16_out=pcm16_out)
        self._last_render_end = start + duration
        return y if pcm16_out else Snippet(start, y)

    def _render_from_device(self, dev, start: int, duration: int):
        bank = self._bank
        if duration <= bank.max_pull:
            # One block.  If the source rendered it ahead of time (behind the previous pull, on the bank's stream) it is
            # already in HBM; then the NEXT contiguous block is requested right after this pull's work, so that a source
            # whose block costs as much as the convolution itself (1024 SuperSaw voices: ~20 us) is off the latency
            # chain of the pull that asks for it.  Same samples: the source's state is snapshotted and rolled back
            # (_drop_speculation) whenever the next pull turns out to be a different one.
            hit = self._spec is not None and self._spec[0] == (start, duration)
            blk = self._spec[1] if hit else dev(start, duration, cuda_stream=bank.stream_ptr)
            self._spec = None
            if blk is None:
                return None  # the source declined (e.g. pull larger than its buffer): host path
            if blk.n_streams != 1 or blk.channels != bank.c_in:
                raise ValueError(f"ConvolvePE src returned {blk.channels} channels, prepared for {bank.c_in}")
            y = bank.process_device_block(blk, interleaved=True)
            if self._speculate and getattr(self._src, "can_speculate", False):
                nxt = dev(start + duration, duration, cuda_stream=bank.stream_ptr, speculative=True)
                if nxt is not None:
                    self._spec = ((start + duration, duration), nxt)
            return y
        outs, pos = [], 0
        while pos < duration:
            d = min(bank.max_pull, duration - pos)
            blk = dev(start + pos, d, cuda_stream=bank.stream_ptr)
            if blk is None:
                if pos:
                    raise RuntimeError("device source stopped producing blocks in the middle of a pull")
                return None  # the source declined (e.g. pull larger than its buffer): host path
            if blk.n_streams != 1 or blk.channels != bank.c_in:
                raise ValueError(f"ConvolvePE src returned {blk.channels} channels, prepared for {bank.c_in}")
            outs.append(bank.process_device_block(blk, interleaved=True))
            pos += d
        return outs[0] if len(outs) == 1 else np.concatenate(outs, axis=0)

    @property
    def bank(self):
        """The device state (None before the first render)."""
        return self._bank

    def __repr__(self) -> str:
        return (f"ConvolvePE(src={self._src.__class__.__name__}, "
                f"fir={self._fir.__class__.__name__}, fft_size={self._fft_size})")
