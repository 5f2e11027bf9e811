"""
Device-resident sources: SinePE, BlitSawPE, SuperSawPE.

Drop-ins for the constant-parameter forms of src/pygmu2/sine_pe.py, blit_saw_pe.py and
super_saw_pe.py.  Host logic (detune ratios, mix gains, the seeded initial phases, harmonic
count rules) is the reference's, parameter for parameter; the samples are produced by the
sm_100a kernels of ``csrc/k_osc.cu`` through ``pgx_osc_*`` -- float64 arithmetic, float32
rounding exactly where the reference rounds.  A PE-valued (modulated) parameter is outside the
path and raises ``NotImplementedError``.

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

    def render_device(self, start: int, duration: int, mix: bool = False, cuda_stream: int = 0) -> DeviceBlock:
        """Enqueue one pull (duration <= max_pull) on ``cuda_stream``; the block lives in the handle's buffer
        until the next render."""
        out = C.c_void_p()
        check(lib().pgx_osc_render_device(self._h, int(start), int(duration), 1 if mix else 0,
                                          C.c_void_p(cuda_stream) if cuda_stream else None, C.byref(out)))
        n = int(duration)
        return DeviceBlock(int(out.value), Layout(self.channels * n, n, 1), 1 if mix else self.n_voices,
                           self.channels, n)


def _const(name: str, v):
    if isinstance(v, ProcessingElement):
        raise NotImplementedError(f"pygmu2_b200: a PE-valued {name} (modulation) is outside the device path")
    return v


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

    def device_block(self, start: int, duration: int, cuda_stream: int = 0) -> DeviceBlock | None:
        if duration > self._max_pull:
            return None
        return self._bank().render_device(start, duration, cuda_stream=cuda_stream)


class SinePE(_OscPE):
    """sine_pe.py:18-270 with constant frequency / amplitude / phase (pure)."""

    def __init__(self, frequency=440.0, amplitude=1.0, phase=0.0, channels: int = 1, *, device: int = 0):
        self._frequency = float(_const("frequency", frequency))
        self._amplitude = float(_const("amplitude", amplitude))
        self._phase = float(_const("phase", phase))
        self._channels, self._device = int(channels), int(device)

    frequency = property(lambda self: self._frequency)
    amplitude = property(lambda self: self._amplitude)
    initial_phase = property(lambda self: self._phase)

    def is_pure(self) -> bool:
        return True

    def _make_bank(self) -> OscBank:
        return OscBank(_lib.PGX_OSC_SINE, [self._frequency], [self._amplitude], [self._phase],
                       channels=self._channels, sample_rate=self.sample_rate, max_pull=self._max_pull,
                       device=self._device)

    def _reset_state(self) -> None:  # stateless
        pass

    _on_start = _on_stop = _reset_state

    def __repr__(self):
        return f"SinePE(frequency={self._frequency}, amplitude={self._amplitude})"


class BlitSawPE(_OscPE):
    """blit_saw_pe.py:26-299 with constant frequency / amplitude / m (never pure: integrator state)."""

    def __init__(self, frequency, amplitude=1.0, initial_phase: float = 0.0, m=None, leak: float = 0.999,
                 channels: int = 1, *, device: int = 0):
        self._frequency = float(_const("frequency", frequency))
        self._amplitude = float(_const("amplitude", amplitude))
        self._initial_phase = float(np.asarray(initial_phase, dtype=np.float64).reshape(-1)[0]) % 1.0
        self._m = None if m is None else int(_const("m", m))
        self._leak, self._channels, self._device = float(leak), int(channels), int(device)

    frequency = property(lambda self: self._frequency)
    amplitude = property(lambda self: self._amplitude)
    m = property(lambda self: self._m)
    leak = property(lambda self: self._leak)
    initial_phase = property(lambda self: self._initial_phase)

    def is_pure(self) -> bool:
        return False

    def _make_bank(self) -> OscBank:
        m_fixed = None if self._m is None else [max(self._m, 1)]     # blit_saw_pe.py:176-177
        return OscBank(_lib.PGX_OSC_BLIT, [self._frequency], [self._amplitude], [self._initial_phase], unison=1,
                       m_fixed=m_fixed, channels=self._channels, sample_rate=self.sample_rate, leak=self._leak,
                       max_pull=self._max_pull, device=self._device)

    def __repr__(self):
        m = "auto" if self._m is None else str(self._m)
        return (f"BlitSawPE(frequency={self._frequency}, amplitude={self._amplitude}, m={m}, "
                f"leak={self._leak}, channels={self._channels})")


class SuperSawPE(_OscPE):
    """super_saw_pe.py:26-342 with constant frequency / amplitude: ``voices`` detuned BlitSaw oscillators."""

    MIX_EQUAL = "equal"
    MIX_CENTER_HEAVY = "center_heavy"
    MIX_LINEAR = "linear"

    def __init__(self, frequency, amplitude=1.0, voices: int = 7, detune_cents: float = 20.0,
                 mix_mode: str = "center_heavy", channels: int = 1, randomize_phase: bool = True,
                 seed: int | None = None, *, device: int = 0):
        if voices < 1:
            voices = 1
        self._frequency = float(_const("frequency", frequency))
        self._amplitude = float(_const("amplitude", amplitude))
        self._voices, self._detune_cents, self._mix_mode = int(voices), detune_cents, mix_mode
        self._channels, self._device = int(channels), int(device)
        self._randomize_phase = bool(randomize_phase)
        self._rng = np.random.default_rng(seed)
        self._detune_ratios = self._compute_detune_ratios()
        self._mix_gains = self._compute_mix_gains()
        # one oscillator per detune ratio; its initial phase is drawn at construction, in order
        # (super_saw_pe.py:219-231)
        self._osc_freq = [float(self._frequency * r) for r in self._detune_ratios]
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
                       unison=self.n_oscillators, amp=[self._amplitude], channels=self._channels,
                       sample_rate=self.sample_rate, max_pull=self._max_pull, device=self._device)

    def __repr__(self):
        return (f"SuperSawPE(frequency={self._frequency}, voices={self._voices}, "
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

    def device_block(self, start: int, duration: int, mix: bool = False, cuda_stream: int = 0) -> DeviceBlock | None:
        if duration > self.max_pull:
            return None
        return self.bank.render_device(start, duration, mix=mix, cuda_stream=cuda_stream)

    def close(self) -> None:
        self.bank.close()
