"""
MixPE -- drop-in for pygmu2's summing PE (src/pygmu2/mix_pe.py:16-153), on the GPU.

Two device paths, chosen by what the inputs are:

* general inputs: every input whose extent intersects the request is rendered (host
  Snippets, as the PE protocol demands) and the float32 left-to-right sum of
  mix_pe.py:92-94 is done by kernel K5 (``pgx_mix_sum``) -- bit-exact with numpy's ``+=``;
* all inputs are fresh ``ConvolvePE``s of one shape, or all are ``SpatialPE(SpatialHRTF)``:
  the inputs are adopted into ONE device bank on the first pull and every later pull is
  a single fused call (FFT -> multiply-accumulate over partitions *and* streams -> one
  inverse FFT per output channel): the sum happens in the frequency domain inside the
  accumulation kernel.  Summation order then differs from left-to-right float32 by
  <= ~N*2^-24 relative, inside the 1e-5 tolerance.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .bank import ConvolveBank, choose_block
from .convolve_pe import ConvolvePE
from .core import Extent, ProcessingElement, Snippet
from .hrtf_bank import HrtfMixBank
from .osc_pe import BlitSawPE, SinePE, SuperSawPE, VoiceBank
from .sources import DelayPE, GainPE
from .spatial_pe import SpatialConstantPower, SpatialHRTF, SpatialLinear, SpatialPE


def device_mix_sum(arrays, device: int = 0) -> np.ndarray:
    """((a0 + a1) + a2) + ... in float32 on the GPU; arrays share one shape."""
    stack = np.ascontiguousarray(np.stack([np.asarray(a, dtype=np.float32) for a in arrays]))
    out = np.empty(stack.shape[1:], dtype=np.float32)
    _lib.require_device()
    _lib.check(_lib.lib().pgx_mix_sum(int(device), _lib.f32_ptr(stack), int(stack.shape[0]),
                                      int(out.size), _lib.f32_ptr(out)))
    return out


class MixPE(ProcessingElement):
    """Sum of two or more PEs (``MixPE(a, b, ...)`` or ``MixPE([a, b, ...])``).

    ``fuse`` (extension, default True) lets the PE adopt bank-able inputs as described above;
    ``fuse=False`` always renders the inputs one by one and sums with K5.
    """

    def __init__(self, *inputs: ProcessingElement, fuse: bool = True, device: int = 0):
        if len(inputs) == 1 and isinstance(inputs[0], (list, tuple)):
            inputs = tuple(inputs[0])
        if len(inputs) < 2:
            raise ValueError("MixPE requires at least 2 inputs")
        self._inputs = list(inputs)
        self._fuse = bool(fuse)
        self._device = int(device)
        self._fused = None       # None = undecided, False = general path, else the adopted bank
        self._fused_pos = None
        self._rest = []          # inputs that stay outside an adopted bank

    def inputs(self) -> list:
        return self._inputs

    def is_pure(self) -> bool:
        return True

    # -- fusion ----------------------------------------------------------------
    def _try_adopt(self, duration: int):
        """-> the device bank the bank-able inputs were adopted into, or False.  ``self._rest`` then holds the inputs
        that stay on the general path (e.g. a source wrapped in a PE-valued GainPE, whose per-sample gain applies AFTER
        its convolution and so cannot ride in a frequency-domain sum): one un-bankable input no longer drops the
        whole mix to N separate pulls."""
        self._rest = []
        if not self._fuse:
            return False
        ins = self._inputs
        if len({type(p) for p in ins}) == 1 and type(ins[0]) in (SuperSawPE, BlitSawPE, SinePE) \
                and all(p.is_pure() or type(p) is not SinePE for p in ins):
            try:
                return _VoiceMix(VoiceBank(ins, device=self._device))
            except ValueError:
                return False  # voices of different shapes: general path
        un = [_unwrap(p) for p in ins]
        conv = [i for i, (c, _, _) in enumerate(un) if type(c) is ConvolvePE and c.bank is None]
        spat = [i for i, (c, _, _) in enumerate(un) if type(c) is SpatialPE and _foldable_method(c.method)
                and c.source.channel_count() is not None]
        for group in sorted((conv, spat), key=len, reverse=True):
            if len(group) < 2:
                continue
            cores, delays, gains = zip(*[un[i] for i in group])
            try:
                if group is conv:
                    bank = self._adopt_convolves(duration, cores, delays, gains)
                else:
                    bank = HrtfMixBank([p.source for p in cores], [p.method for p in cores], pull_hint=duration,
                                       device=self._device, delays=delays, gains=gains)
            except _NotFusable:
                continue
            taken = set(group)
            self._rest = [p for i, p in enumerate(ins) if i not in taken]
            return bank
        return False

    def _adopt_convolves(self, duration: int, ins, delays, gains):
        taps, src_ch, exts = [], set(), []
        for p, d in zip(ins, delays):
            # settings of the PE that a shared bank would silently drop: keep such inputs on their own banks
            if p._device != self._device or p._block_size is not None or p._tail_block or p._out_gains is not None:
                raise _NotFusable
            ext = p.extent()  # validates the filter contract (raises ValueError like the reference)
            exts.append(Extent(None if ext.start is None else ext.start + d, None if ext.end is None else ext.end + d))
            fe = p.fir.extent()
            if fe.start != 0 or fe.end is None or fe.end < 1:
                raise _NotFusable
            h = p.fir.render(0, int(fe.end)).data
            taps.append(h)
            src_ch.add(p.src.channel_count())
            if p.fft_size is not None and p.fft_size < h.shape[0]:
                raise _NotFusable  # let the per-PE path raise the reference's ValueError
            del ext
        if len({h.shape for h in taps}) != 1 or len(src_ch) != 1 or None in src_ch:
            raise _NotFusable
        # one resident filter per input + an all-zero one for inputs MixPE does not render (mix_pe.py:81-85)
        bank = ConvolveBank(np.stack(taps + [np.zeros_like(taps[0])]), len(ins), int(src_ch.pop()),
                            block=choose_block(taps[0].shape[0], duration), device=self._device,
                            filter_of_stream=np.arange(len(ins), dtype=np.int32))
        bank.attach_sources([p.src for p in ins], delays=delays, gains=gains, extents=exts, silent_filter=len(ins))
        bank.mix_output = True
        return bank

    def _render(self, start: int, duration: int) -> Snippet:
        if self._fused is None:
            self._fused = self._try_adopt(duration)
        if self._fused is not False:
            y = self._fused.render(start, duration)  # (C_out, n); resets itself on a non-contiguous pull
            out = np.ascontiguousarray(y.T)
            if self._rest:                           # inputs outside the bank: rendered as they are, added in order
                req = Extent(start, start + duration)
                for p in self._rest:
                    if p.extent().intersects(req):
                        d = p.render(start, duration).data
                        if d.shape[1] != out.shape[1]:
                            raise ValueError(f"MixPE input channel mismatch: fused inputs have {out.shape[1]} channels, "
                                             f"{p.__class__.__name__} has {d.shape[1]} channels")
                        out += d
            return Snippet(start, out)
        # general path: mix_pe.py:80-96
        req = Extent(start, start + duration)
        rendered = [p.render(start, duration).data for p in self._inputs if p.extent().intersects(req)]
        if not rendered:
            return Snippet.from_zeros(start, duration, self.channel_count() or 1)
        if len(rendered) == 1:
            return Snippet(start, rendered[0].copy())
        return Snippet(start, device_mix_sum(rendered, self._device))

    @property
    def fused_bank(self):
        """The device bank the inputs were adopted into (None before the first pull / when the inputs are not
        bank-able): ``HrtfMixBank``, ``ConvolveBank`` or the voice mix."""
        return self._fused if self._fused not in (None, False) else None

    def set_trajectory(self, azimuth, elevation=None, *, hop: int, start: int = 0) -> None:
        """Moving HRTF sources without per-pull host work (extension): see ``HrtfMixBank.set_trajectory``.  Adopts
        the inputs first if that has not happened yet (``hop`` doubles as the pull-size hint)."""
        if self._fused is None:
            self._fused = self._try_adopt(int(hop))
        if not isinstance(self._fused, HrtfMixBank):
            raise ValueError("set_trajectory needs a MixPE whose inputs are all SpatialPE(..., SpatialHRTF) sources")
        self._fused.set_trajectory(azimuth, elevation, hop=hop, start=start)

    def device_block(self, start: int, duration: int, cuda_stream: int = 0, speculative: bool = False):
        """The mix left in HBM (only when the inputs are oscillator voices fused into one VoiceBank)."""
        if self._fused is None:
            self._fused = self._try_adopt(duration)
        if isinstance(self._fused, _VoiceMix):
            return self._fused.device_block(start, duration, cuda_stream, speculative)
        return None

    @property
    def can_speculate(self) -> bool:
        """A voice mix depends on nothing but (start, duration) and the oscillators' carried state: its consumer may
        render the next block ahead of the pull that asks for it."""
        return isinstance(self._fused, _VoiceMix)

    def rollback_speculation(self, cuda_stream: int = 0) -> None:
        if isinstance(self._fused, _VoiceMix):
            self._fused.vb.rollback(cuda_stream)

    def _reset_state(self) -> None:
        self._fused_pos = None
        if self._fused not in (None, False):
            self._fused.reset()

    _on_start = _on_stop = _reset_state

    def _compute_extent(self) -> Extent:
        result = self._inputs[0].extent()
        for p in self._inputs[1:]:
            result = result.union(p.extent())
        return result

    def channel_count(self):
        return self._inputs[0].channel_count() if self._inputs else None

    def resolve_channel_count(self, input_channel_counts):
        if not input_channel_counts:
            raise ValueError("MixPE has no inputs")
        first = input_channel_counts[0]
        for i, count in enumerate(input_channel_counts[1:], start=2):
            if count != first:
                raise ValueError(f"MixPE input channel mismatch: input 1 has {first} channels, "
                                 f"input {i} has {count} channels")
        return first

    def __repr__(self):
        return f"MixPE({', '.join(p.__class__.__name__ for p in self._inputs)})"


class _NotFusable(Exception):
    pass


def _unwrap(pe):
    """Peel integer DelayPE and constant GainPE wrappers off a MixPE input:
    -> (core PE, total delay in samples, combined float32 gain or None)."""
    delay, gain = 0, None
    while True:
        if type(pe) is DelayPE and pe.mode == "int":
            delay += pe.delay
            pe = pe.source
        elif type(pe) is GainPE and not isinstance(pe.gain, ProcessingElement):
            g = np.float32(pe.gain)
            gain = g if gain is None else np.float32(gain * g)
            pe = pe.source
        else:
            return pe, delay, gain


def _foldable_method(m) -> bool:
    if type(m) is SpatialHRTF:
        return m._bank is None
    if type(m) in (SpatialLinear, SpatialConstantPower):
        return not isinstance(m.azimuth, ProcessingElement)
    return False


class _VoiceMix:
    """MixPE over oscillator voices: one VoiceBank launch plus the bit-exact float32 voice sum (K5)."""

    def __init__(self, vb: VoiceBank):
        self.vb = vb

    def render(self, start: int, duration: int) -> np.ndarray:
        return self.vb.render(start, duration, mix=True)

    def device_block(self, start: int, duration: int, cuda_stream: int = 0, speculative: bool = False):
        return self.vb.device_block(start, duration, mix=True, cuda_stream=cuda_stream, speculative=speculative)

    def reset(self) -> None:
        self.vb.reset()
