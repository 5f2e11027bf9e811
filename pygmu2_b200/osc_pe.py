"""
Device-resident sources: SinePE, BlitSawPE, SuperSawPE.

Drop-ins for the constant-parameter forms of src/pygmu2/sine_pe.py, blit_saw_pe.py and
super_saw_pe.py.  Host logic (detune ratios, mix gains, the seeded initial phases, harmonic
count rules) is the reference's, parameter for parameter; the samples are produced by the
sm_100a kernels of ``csrc/k_osc.cu`` through ``pgx_osc_*`` -- float64 arithmetic, float32
rounding exactly where the reference rounds.  PE-valued (modulated) parameters -- frequency, amplitude,
phase, harmonic count -- are rendered for the pull and handed to the device as control vectors
(``pgx_osc_render_modulated``); the stateful recurrences are walked in the reference's own order.

Besides ``render()`` (host Snippet, as the PE protocol demands) every class offers
``device_block(start, duration, cuda_stream)``: the same samples left in HBM, which is how
``ConvolvePE`` / ``MixPE`` / ``ConvolveBank`` consume them without a host round trip.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple

import numpy as np

from . import _lib
from ._lib import Layout, OscConfig, check, lib
from .core import Extent, ProcessingElement, Snippet


class DeviceBlock(NamedTuple):
    """Samples resident on the device: element (stream s, channel c, sample i) at ptr[s*stream + c*chan + i*samp]."""
    ptr: int
    layout: Layout
    n_streams: int
    channels: int
    duration: int


def _f64(a, n=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} values, got {a.size}")
    return a


class OscBank:
    """Owner of one ``pgx_osc`` handle: V sine streams, or V voices x U BLIT sawtooth oscillators."""

    def __init__(self, kind: int, freq, gain, phase, *, unison: int = 1, amp=None, m_fixed=None, channels: int = 1,
                 sample_rate: int, leak: float = 0.999, max_pull: int = 8192, device: int = 0):
        self.kind, self.unison, self.channels = int(kind), int(unison), int(channels)
        freq, gain, phase = _f64(freq), _f64(gain), _f64(phase)
        n_osc = freq.size
        if n_osc < 1 or n_osc % self.unison or gain.size != n_osc or phase.size != n_osc:
            raise ValueError("freq / gain / phase must have n_voices * unison entries each")
        self.n_voices = n_osc // self.unison
        self.max_pull, self.device = int(max_pull), int(device)
        amp_a = _f64(amp if amp is not None else np.ones(self.n_voices), self.n_voices)
        m_a = None
        if m_fixed is not None:
            m_a = np.ascontiguousarray(np.asarray(m_fixed, dtype=np.int32).reshape(-1))
            if m_a.size != n_osc:
                raise ValueError("m_fixed must have one entry per oscillator")
        _lib.require_device()
        cfg = OscConfig(device=self.device, kind=self.kind, n_voices=self.n_voices, unison=self.unison,
                        channels=self.channels, sample_rate=int(sample_rate), max_pull=self.max_pull, reserved=0,
                        leak=float(leak))
        self._h = C.c_void_p()
        check(lib().pgx_osc_create(C.byref(self._h), C.byref(cfg), freq.ctypes.data, gain.ctypes.data,
                                   phase.ctypes.data, m_a.ctypes.data if m_a is not None else None,
                                   amp_a.ctypes.data))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().pgx_osc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self) -> None:
        check(lib().pgx_osc_reset(self._h))

    @property
    def launches(self) -> int:
        n = C.c_int64(0)
        check(lib().pgx_osc_launches(self._h, C.byref(n)))
        return int(n.value)

    def render(self, start: int, duration: int, mix: bool = False) -> np.ndarray:
        """(V, C, n) float32, or the MixPE sum over voices (C, n); host memory."""
        rows = () if mix else (self.n_voices,)
        y = np.empty(rows + (self.channels, duration), dtype=np.float32)
        pos = 0
        while pos < duration:
            d = min(self.max_pull, duration - pos)
            yc = y if d == duration else np.empty(rows + (self.channels, d), dtype=np.float32)
            check(lib().pgx_osc_render(self._h, int(start + pos), d, 1 if mix else 0, _lib.f32_ptr(yc)))
            if d != duration:
                y[..., pos:pos + d] = yc
            pos += d
        return y

    def _modulated(self, duration: int, freq, amp, phase, mix: bool, cuda_stream: int, y_host, amp_osc: bool = False):
        n = int(duration)
        ptrs, keep, on_host, on_dev = [], [], False, False
        for ctl in (freq, amp, phase):
            if ctl is None:
                ptrs.append(None)
            elif isinstance(ctl, DeviceBlock):
                if ctl.n_streams * ctl.channels != self.n_voices or ctl.duration != n or ctl.layout.samp != 1:
                    raise ValueError("control block must be (n_voices, n) planar")
                ptrs.append(C.c_void_p(ctl.ptr))
                on_dev = True
            else:
                a = np.ascontiguousarray(ctl, dtype=np.float32).reshape(self.n_voices, n)
                keep.append(a)
                ptrs.append(a.ctypes.data_as(C.c_void_p))
                on_host = True
        if on_host and on_dev:
            raise ValueError("control vectors must be all host arrays or all device blocks")
        out = C.c_void_p()
        check(lib().pgx_osc_render_modulated(self._h, n, 1 if mix else 0, ptrs[0], ptrs[1], ptrs[2],
                                             (_lib.PGX_CTL_HOST if on_host else 0) | (_lib.PGX_CTL_AMP_OSC if amp_osc else 0),
                                             C.c_void_p(cuda_stream) if cuda_stream else None, C.byref(out),
                                             y_host.ctypes.data_as(C.c_void_p) if y_host is not None else None))
        return DeviceBlock(int(out.value), Layout(self.channels * n, n, 1), 1 if mix else self.n_voices,
                           self.channels, n)

    def render_modulated_device(self, duration: int, freq=None, amp=None, phase=None, *, mix: bool = False,
                                cuda_stream: int = 0, amp_osc: bool = False) -> DeviceBlock:
        """Modulated sine pull (stateful branch of sine_pe.py): ``freq`` / ``amp`` / ``phase`` are per-sample control
        vectors -- host float32 arrays (n_voices, n) or ``DeviceBlock``s produced on ``cuda_stream`` -- or None for a
        parameter that is constant."""
        return self._modulated(duration, freq, amp, phase, mix, cuda_stream, None, amp_osc)

    def render_modulated(self, duration: int, freq=None, amp=None, phase=None, *, amp_osc: bool = False) -> np.ndarray:
        """Same pull delivered to the host: (n_voices, channels, n) float32."""
        y = np.empty((self.n_voices, self.channels, int(duration)), dtype=np.float32)
        self._modulated(duration, freq, amp, phase, False, 0, y, amp_osc)
        return y

    def rollback(self, cuda_stream: int = 0) -> None:
        """Undo a speculative pull (``render_device(..., speculative=True)``) that turned out not to be the next one."""
        check(lib().pgx_osc_rollback(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def render_device(self, start: int, duration: int, mix: bool = False, cuda_stream: int = 0,
                      speculative: bool = False) -> DeviceBlock:
        """Enqueue one pull (duration <= max_pull) on ``cuda_stream``; the block lives in the handle's buffer
        until the next render.  ``speculative``: keep the state the pull starts from (see ``rollback``)."""
        out = C.c_void_p()
        check(lib().pgx_osc_render_device(self._h, int(start), int(duration),
                                          (1 if mix else 0) | (_lib.PGX_OSC_SNAPSHOT if speculative else 0),
                                          C.c_void_p(cuda_stream) if cuda_stream else None, C.byref(out)))
        n = int(duration)
        return DeviceBlock(int(out.value), Layout(self.channels * n, n, 1), 1 if mix else self.n_voices,
                           self.channels, n)


class _OscPE(ProcessingElement):
    """Shared plumbing: lazy device handle, host render, device block, start/stop reset."""

    _osc: OscBank | None = None
    _channels = 1
    _device = 0
    _max_pull = 8192

    def inputs(self) -> list:
        return []

    def channel_count(self) -> int:
        return self._channels

    def _compute_extent(self) -> Extent:
        return Extent(None, None)

    def _make_bank(self) -> OscBank:
        raise NotImplementedError

    def _bank(self) -> OscBank:
        if self._osc is None:
            if self.sample_rate is None:
                raise RuntimeError("Sample rate not set. Call pg.set_sample_rate() first.")
            self._osc = self._make_bank()
        return self._osc

    def _reset_state(self) -> None:
        if self._osc is not None:
            self._osc.reset()

    _on_start = _on_stop = _reset_state

    def _render(self, start: int, duration: int) -> Snippet:
        y = self._bank().render(start, duration)           # (1, C, n)
        return Snippet(start, np.ascontiguousarray(y[0].T))

    def device_block(self, start: int, duration: int, cuda_stream: int = 0, speculative: bool = False) -> DeviceBlock | None:
        if duration > self._max_pull:
            return None
        return self._bank().render_device(start, duration, cuda_stream=cuda_stream, speculative=speculative)

    def rollback_speculation(self, cuda_stream: int = 0) -> None:
        if self._osc is not None:
            self._osc.rollback(cuda_stream)

    can_speculate = True      # constant parameters: a block depends on nothing but (start, duration) and the carried state


class SinePE(_OscPE):
    """sine_pe.py:18-270.  Constant frequency / amplitude / phase: pure, phase from the sample index.  Any of them a
    PE (FM / AM / PM): the reference's stateful branch (:188-232) -- the parameter PEs are rendered for the pull
    (``_scalar_or_pe_values``: channel 0, widened to float64) and the phase is the running float64 sum of
    2 pi f / sr, carried from pull to pull; on the device that sum is walked left to right like np.cumsum."""

    def __init__(self, frequency=440.0, amplitude=1.0, phase=0.0, channels: int = 1, *, device: int = 0):
        self._params = {"frequency": frequency, "amplitude": amplitude, "phase": phase}
        self._pe_params = {k: v for k, v in self._params.items() if isinstance(v, ProcessingElement)}
        self._frequency = frequency if "frequency" in self._pe_params else float(frequency)
        self._amplitude = amplitude if "amplitude" in self._pe_params else float(amplitude)
        self._phase = phase if "phase" in self._pe_params else float(phase)
        self._channels, self._device = int(channels), int(device)

    frequency = property(lambda self: self._frequency)
    amplitude = property(lambda self: self._amplitude)
    initial_phase = property(lambda self: self._phase)

    def inputs(self) -> list:
        return [self._params[k] for k in ("frequency", "amplitude", "phase") if k in self._pe_params]   # sine_pe.py:91-100

    def is_pure(self) -> bool:
        return not self._pe_params                                   # sine_pe.py:102-107

    def _compute_extent(self) -> Extent:
        result = Extent(None, None)                                  # sine_pe.py:234-245
        for pe in self.inputs():
            result = result.intersection(pe.extent())
        return result

    def _const_or_zero(self, name):
        return 0.0 if name in self._pe_params else float(self._params[name])

    def _make_bank(self) -> OscBank:
        return OscBank(_lib.PGX_OSC_SINE, [self._const_or_zero("frequency")], [self._const_or_zero("amplitude")],
                       [self._const_or_zero("phase")], channels=self._channels, sample_rate=self.sample_rate,
                       max_pull=self._max_pull, device=self._device)

    def _controls(self, start: int, duration: int):
        """The PE-valued parameters rendered for this pull: channel 0 of each (processing_element.py:296-340)."""
        return {k: np.ascontiguousarray(pe.render(start, duration).data[:, 0], dtype=np.float32)
                for k, pe in self._pe_params.items()}

    def _render(self, start: int, duration: int) -> Snippet:
        if not self._pe_params:
            return super()._render(start, duration)
        bank, outs, pos = self._bank(), [], 0
        while pos < duration:
            d = min(self._max_pull, duration - pos)
            ctl = self._controls(start + pos, d)
            outs.append(bank.render_modulated(d, ctl.get("frequency"), ctl.get("amplitude"), ctl.get("phase"))[0].T)
            pos += d
        return Snippet(start, np.ascontiguousarray(outs[0] if len(outs) == 1 else np.concatenate(outs, axis=0)))

    can_speculate = property(lambda self: not self._pe_params)   # control PEs cannot be rendered ahead of their time

    def device_block(self, start: int, duration: int, cuda_stream: int = 0, speculative: bool = False) -> DeviceBlock | None:
        if duration > self._max_pull:
            return None
        if not self._pe_params:
            return self._bank().render_device(start, duration, cuda_stream=cuda_stream, speculative=speculative)
        if speculative:
            return None
        ctl = self._controls(start, duration)
        return self._bank().render_modulated_device(duration, ctl.get("frequency"), ctl.get("amplitude"),
                                                    ctl.get("phase"), cuda_stream=cuda_stream)

    def _reset_state(self) -> None:
        if self._pe_params and self._osc is not None:                # sine_pe.py:110-118: the accumulated phase starts over
            self._osc.reset()

    _on_start = _on_stop = _reset_state

    def __repr__(self):
        f = self._frequency.__class__.__name__ if "frequency" in self._pe_params else self._frequency
        a = self._amplitude.__class__.__name__ if "amplitude" in self._pe_params else self._amplitude
        return f"SinePE(frequency={f}, amplitude={a})"


class _ModulatedBlit(_OscPE):
    """BLIT oscillators whose frequency and / or amplitude may be PEs (blit_saw_pe.py:161-163, super_saw_pe.py:223-246,
    287): the parameter PEs are rendered for the pull (channel 0, float32) and handed to the device as control
    vectors; the phase and the leaky integrator are walked in the reference's order and carried between pulls."""

    _amp_per_osc = False     # BlitSawPE: the amplitude scales the oscillator; SuperSawPE: the float64 voice sum
    _pe_params: dict = {}
    _last_end = None

    def _split_params(self, frequency, amplitude, m=None):
        self._pe_params = {k: v for k, v in (("frequency", frequency), ("amplitude", amplitude), ("m", m))
                           if isinstance(v, ProcessingElement)}
        f = frequency if "frequency" in self._pe_params else float(frequency)
        a = amplitude if "amplitude" in self._pe_params else float(amplitude)
        return f, a

    def inputs(self) -> list:
        return [self._pe_params[k] for k in ("frequency", "amplitude", "m") if k in self._pe_params]   # blit_saw_pe.py:110-119

    def _compute_extent(self) -> Extent:
        result = Extent(None, None)
        for pe in self.inputs():
            result = result.intersection(pe.extent())
        return result

    def _controls(self, start: int, duration: int):
        return {k: np.ascontiguousarray(pe.render(start, duration).data[:, 0], dtype=np.float32)
                for k, pe in self._pe_params.items()}

    def _render(self, start: int, duration: int) -> Snippet:
        if not self._pe_params:
            return super()._render(start, duration)
        bank = self._bank()
        if self._last_end is None or start != self._last_end:      # blit_saw_pe.py:183-186: a new run
            bank.reset()
        outs, pos = [], 0
        while pos < duration:
            d = min(self._max_pull, duration - pos)
            ctl = self._controls(start + pos, d)
            outs.append(bank.render_modulated(d, ctl.get("frequency"), ctl.get("amplitude"), ctl.get("m"),
                                              amp_osc=self._amp_per_osc)[0].T)
            pos += d
        self._last_end = start + duration
        return Snippet(start, np.ascontiguousarray(outs[0] if len(outs) == 1 else np.concatenate(outs, axis=0)))

    can_speculate = property(lambda self: not self._pe_params)

    def device_block(self, start: int, duration: int, cuda_stream: int = 0, speculative: bool = False) -> DeviceBlock | None:
        if duration > self._max_pull:
            return None
        if not self._pe_params:
            return self._bank().render_device(start, duration, cuda_stream=cuda_stream, speculative=speculative)
        if speculative:
            return None
        bank = self._bank()
        if self._last_end is None or start != self._last_end:
            bank.reset()
        ctl = self._controls(start, duration)
        self._last_end = start + duration
        return bank.render_modulated_device(duration, ctl.get("frequency"), ctl.get("amplitude"), ctl.get("m"),
                                            cuda_stream=cuda_stream, amp_osc=self._amp_per_osc)

    def _reset_state(self) -> None:
        self._last_end = None
        if self._osc is not None:
            self._osc.reset()

    _on_start = _on_stop = _reset_state


class BlitSawPE(_ModulatedBlit):
    """blit_saw_pe.py:26-299; frequency / amplitude constant or PE-valued, m constant (never pure: integrator state)."""

    _amp_per_osc = True

    def __init__(self, frequency, amplitude=1.0, initial_phase: float = 0.0, m=None, leak: float = 0.999,
                 channels: int = 1, *, device: int = 0):
        self._frequency, self._amplitude = self._split_params(frequency, amplitude, m)
        self._initial_phase = float(np.asarray(initial_phase, dtype=np.float64).reshape(-1)[0]) % 1.0
        self._m = None if m is None else (m if "m" in self._pe_params else int(m))
        self._leak, self._channels, self._device = float(leak), int(channels), int(device)

    frequency = property(lambda self: self._frequency)
    amplitude = property(lambda self: self._amplitude)
    m = property(lambda self: self._m)
    leak = property(lambda self: self._leak)
    initial_phase = property(lambda self: self._initial_phase)

    def is_pure(self) -> bool:
        return False

    def _make_bank(self) -> OscBank:
        m_fixed = None if (self._m is None or "m" in self._pe_params) else [max(self._m, 1)]     # blit_saw_pe.py:176-177
        # a PE-valued frequency is used as it is (ratio 1), a PE-valued amplitude replaces the oscillator amplitude
        f = 1.0 if "frequency" in self._pe_params else self._frequency
        a = 1.0 if "amplitude" in self._pe_params else self._amplitude
        return OscBank(_lib.PGX_OSC_BLIT, [f], [a], [self._initial_phase], unison=1,
                       m_fixed=m_fixed, channels=self._channels, sample_rate=self.sample_rate, leak=self._leak,
                       max_pull=self._max_pull, device=self._device)

    def __repr__(self):
        m = "auto" if self._m is None else (self._m.__class__.__name__ if "m" in self._pe_params else str(self._m))
        return (f"BlitSawPE(frequency={self._frequency}, amplitude={self._amplitude}, m={m}, "
                f"leak={self._leak}, channels={self._channels})")


class SuperSawPE(_ModulatedBlit):
    """super_saw_pe.py:26-342: ``voices`` detuned BlitSaw oscillators; frequency / amplitude constant or PE-valued
    (each oscillator then runs at float32(frequency * ratio), GainPE's product, super_saw_pe.py:236-240)."""

    MIX_EQUAL = "equal"
    MIX_CENTER_HEAVY = "center_heavy"
    MIX_LINEAR = "linear"

    def __init__(self, frequency, amplitude=1.0, voices: int = 7, detune_cents: float = 20.0,
                 mix_mode: str = "center_heavy", channels: int = 1, randomize_phase: bool = True,
                 seed: int | None = None, *, device: int = 0):
        if voices < 1:
            voices = 1
        self._frequency, self._amplitude = self._split_params(frequency, amplitude)
        self._voices, self._detune_cents, self._mix_mode = int(voices), detune_cents, mix_mode
        self._channels, self._device = int(channels), int(device)
        self._randomize_phase = bool(randomize_phase)
        self._rng = np.random.default_rng(seed)
        self._detune_ratios = self._compute_detune_ratios()
        self._mix_gains = self._compute_mix_gains()
        # one oscillator per detune ratio; its initial phase is drawn at construction, in order
        # (super_saw_pe.py:219-231)
        # a PE-valued frequency: the oscillators carry their detune RATIOS, the product is taken per sample on the device
        self._osc_freq = [float(r) if "frequency" in self._pe_params else float(self._frequency * r)
                          for r in self._detune_ratios]
        self._osc_gain = [float(self._mix_gains[i]) for i in range(len(self._detune_ratios))]
        self._osc_phase = [float(self._rng.random(1)[0]) % 1.0 if self._randomize_phase else 0.0
                           for _ in self._detune_ratios]

    frequency = property(lambda self: self._frequency)
    amplitude = property(lambda self: self._amplitude)
    voices = property(lambda self: self._voices)
    detune_cents = property(lambda self: self._detune_cents)
    mix_mode = property(lambda self: self._mix_mode)

    def _compute_detune_ratios(self) -> np.ndarray:      # super_saw_pe.py:128-145
        if self._voices == 1 or self._detune_cents == 0:
            return np.array([1.0])
        cents = np.linspace(-self._detune_cents, self._detune_cents, self._voices)
        return 2 ** (cents / 1200.0)

    def _compute_mix_gains(self) -> np.ndarray:          # super_saw_pe.py:148-207
        n, mode = self._voices, self._mix_mode
        if n == 1:
            return np.array([1.0])
        gains = np.ones(n, dtype=np.float32)
        if mode == self.MIX_EQUAL:
            pass
        elif mode == self.MIX_LINEAR:
            d = np.abs(np.arange(n, dtype=np.float32) - (n - 1) / 2.0)
            gains = 0.5 + 0.5 * (1.0 - d / np.max(d))
        elif mode == self.MIX_CENTER_HEAVY:
            gains[:] = 0.5
            if n % 2 == 1:
                gains[n // 2] = 1.0
            else:
                gains[n // 2 - 1] = 1.0
                gains[n // 2] = 1.0
        else:
            raise ValueError(f"Unknown mix mode: {mode}")
        return gains / np.sqrt(np.sum(gains ** 2))

    def is_pure(self) -> bool:
        return False

    @property
    def n_oscillators(self) -> int:
        return len(self._osc_freq)

    def _make_bank(self) -> OscBank:
        return OscBank(_lib.PGX_OSC_BLIT, self._osc_freq, self._osc_gain, self._osc_phase,
                       unison=self.n_oscillators, amp=[1.0 if "amplitude" in self._pe_params else self._amplitude],
                       channels=self._channels,
                       sample_rate=self.sample_rate, max_pull=self._max_pull, device=self._device)

    def __repr__(self):
        f = self._frequency.__class__.__name__ if "frequency" in self._pe_params else self._frequency
        return (f"SuperSawPE(frequency={f}, voices={self._voices}, "
                f"detune_cents={self._detune_cents}, mix_mode={self._mix_mode!r})")


class VoiceBank:
    """Many SuperSawPE / BlitSawPE / SinePE of one shape as ONE device handle, optionally mixed (MixPE) --
    the C5 front end: 1024 voices are one kernel launch plus the bit-exact float32 voice sum."""

    def __init__(self, pes, *, max_pull: int = 8192, device: int = 0):
        pes = list(pes)
        if not pes:
            raise ValueError("VoiceBank needs at least one PE")
        kinds = {type(p) for p in pes}
        if len(kinds) != 1 or not issubclass(next(iter(kinds)), (SinePE, BlitSawPE, SuperSawPE)):
            raise ValueError("VoiceBank needs PEs of one oscillator class")
        chans = {p.channel_count() for p in pes}
        if len(chans) != 1:
            raise ValueError("VoiceBank voices must share a channel count")
        self.pes, self.channels = pes, int(chans.pop())
        sr = pes[0].sample_rate
        if sr is None:
            raise RuntimeError("Sample rate not set. Call pg.set_sample_rate() first.")
        p0 = pes[0]
        if any(getattr(p, "_pe_params", None) for p in pes):
            raise ValueError("VoiceBank voices must have constant parameters (a modulated oscillator renders on its own)")
        if isinstance(p0, SinePE):
            self.bank = OscBank(_lib.PGX_OSC_SINE, [p._frequency for p in pes], [p._amplitude for p in pes],
                                [p._phase for p in pes], channels=self.channels, sample_rate=sr,
                                max_pull=max_pull, device=device)
        elif isinstance(p0, BlitSawPE):
            if len({p._leak for p in pes}) != 1:
                raise ValueError("VoiceBank BlitSaw voices must share the leak coefficient")
            self.bank = OscBank(_lib.PGX_OSC_BLIT, [p._frequency for p in pes], [p._amplitude for p in pes],
                                [p._initial_phase for p in pes], unison=1,
                                m_fixed=[0 if p._m is None else max(p._m, 1) for p in pes],
                                channels=self.channels, sample_rate=sr, leak=p0._leak, max_pull=max_pull, device=device)
        else:
            if len({p.n_oscillators for p in pes}) != 1:
                raise ValueError("VoiceBank SuperSaw voices must share the oscillator count")
            self.bank = OscBank(_lib.PGX_OSC_BLIT, np.concatenate([p._osc_freq for p in pes]),
                                np.concatenate([p._osc_gain for p in pes]), np.concatenate([p._osc_phase for p in pes]),
                                unison=p0.n_oscillators, amp=[p._amplitude for p in pes], channels=self.channels,
                                sample_rate=sr, max_pull=max_pull, device=device)
        self.n_voices, self.max_pull = len(pes), int(max_pull)

    def reset(self) -> None:
        self.bank.reset()

    def render(self, start: int, duration: int, mix: bool = False) -> np.ndarray:
        return self.bank.render(start, duration, mix=mix)

    def device_block(self, start: int, duration: int, mix: bool = False, cuda_stream: int = 0,
                     speculative: bool = False) -> DeviceBlock | None:
        if duration > self.max_pull:
            return None
        return self.bank.render_device(start, duration, mix=mix, cuda_stream=cuda_stream, speculative=speculative)

    def rollback(self, cuda_stream: int = 0) -> None:
        self.bank.rollback(cuda_stream)

    def close(self) -> None:
        self.bank.close()
