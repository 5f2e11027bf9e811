"""
CPU oracle for the device-resident sources of SURVEY.md §8(f) rank 1: constant-parameter SinePE,
BlitSawPE and SuperSawPE of rdpoor/pygmu2.

TEST INFRASTRUCTURE ONLY (same rule as pygmu2_oracle.py: nothing under ``pygmu2_b200/`` imports it).

A numpy/scipy restatement, function by function, of
  sine_pe.py:119-175       (pure sine: float64 phase from the sample index)
  sine_pe.py:135-232       (modulated sine, PE-valued parameters: running float64 phase, carried between pulls)
  blit_saw_pe.py:152-264   (BLIT: Dirichlet kernel / period, DC removed, leaky integrator via lfilter)
  super_saw_pe.py:128-318  (detune ratios, mix gains, rng phases, float64 sum of float32 voices)
Parity pinned: ``tests/test_oracle_golden.py`` checks it against ``tests/golden/src_*.npz`` produced by the
real reference (``oracle/gen_golden_sources.py``).
"""
from __future__ import annotations

import numpy as np
from scipy.signal import lfilter


def sine(frequency: float, amplitude: float, phase: float, sample_rate: int, start: int, duration: int) -> np.ndarray:
    """sine_pe.py:135-175 with constant parameters -> (duration,) float32."""
    idx = np.arange(start, start + duration, dtype=np.float64)          # :173
    time = idx / sample_rate                                            # :174
    ph = phase + 2.0 * np.pi * frequency * time                         # :175
    return (np.float64(amplitude) * np.sin(ph)).astype(np.float32)      # :146,157


class OracleSineModulated:
    """sine_pe.py:135-232, the stateful branch taken when any of frequency / amplitude / phase is a PE.  ``freq`` /
    ``amp`` / ``phase`` given to render() are the control vectors the parameter PEs rendered for the pull (float32,
    channel 0), or None where the parameter is the constant passed to the constructor."""

    def __init__(self, frequency=440.0, amplitude=1.0, phase=0.0, sample_rate=44100, channels=1):
        self.f0, self.a0, self.p0, self.sr, self.channels = float(frequency), float(amplitude), float(phase), int(sample_rate), channels
        self.reset()

    def reset(self):                                                    # :110-118
        self.acc, self.init = 0.0, False

    def render(self, duration, freq=None, amp=None, phase=None):
        f = np.full(duration, self.f0) if freq is None else np.asarray(freq, np.float32).astype(np.float64)   # :135
        a = (np.full(duration, self.a0) if amp is None else np.asarray(amp, np.float32).astype(np.float64)).reshape(-1, 1)
        inc = (2.0 * np.pi * f.reshape(-1, 1) / self.sr)                # :198
        if not self.init:                                               # :203-210
            initial = self.p0 if phase is None else 0.0
            self.init = True
        else:
            initial = self.acc
        cum = np.cumsum(inc, axis=0) + initial                          # :216
        # :137,219-222: _scalar_or_pe_values turns a CONSTANT phase into a constant array too (processing_element.py:
        # 360-365), so "isinstance(phase_mod, np.ndarray)" always holds and the constant is added to every sample -- on
        # top of its use as the initial phase -- and, through :229, again on every later pull.  Reference behaviour,
        # restated as it is.
        pm = np.full(duration, self.p0) if phase is None else np.asarray(phase, np.float32).astype(np.float64)
        cum = cum + pm.reshape(-1, 1)
        self.acc = float(cum[-1, 0])                                    # :229
        y = a * np.sin(cum)                                             # :146
        if self.channels > 1:
            y = np.tile(y, (1, self.channels))                          # :153-155
        return y.astype(np.float32)                                     # :157


class OracleBlitSaw:
    """blit_saw_pe.py:67-264, constant frequency / amplitude, m=None (auto) or a fixed int."""

    def __init__(self, frequency, amplitude=1.0, initial_phase=0.0, m=None, leak=0.999, sample_rate=44100):
        self.f, self.amp = float(frequency), float(amplitude)
        self.initial_phase = float(np.asarray(initial_phase).reshape(-1)[0]) % 1.0   # :77
        self.m, self.leak, self.sr = m, float(leak), int(sample_rate)
        self.reset()

    def reset(self):                                                    # :137-142
        self.phase, self.integ, self.last_end = self.initial_phase, 0.0, None

    def render(self, start: int, duration: int, freq_ctl=None, amp_ctl=None, m_ctl=None) -> np.ndarray:
        """freq_ctl / amp_ctl / m_ctl: what a PE-valued frequency / amplitude / m rendered for this pull (float32,
        widened by _scalar_or_pe_values, :162-163,176), or None for the constants."""
        freq = np.full(duration, self.f, dtype=np.float64) if freq_ctl is None else \
            np.asarray(freq_ctl, np.float32).astype(np.float64)         # :162
        amp = np.full(duration, self.amp, dtype=np.float64) if amp_ctl is None else \
            np.asarray(amp_ctl, np.float32).astype(np.float64)          # :163
        if m_ctl is not None:                                           # :175-177
            m = np.maximum(np.asarray(m_ctl, np.float32).astype(np.float64).astype(np.int32), 1).astype(np.float64)
        elif self.m is None:                                            # :167-174
            m = np.floor(self.sr / (2.0 * np.maximum(freq, 1.0))).astype(np.int32)
            m = m - (1 - m % 2)
            m = np.maximum(m, 1)
        else:
            m = np.maximum(np.full(duration, float(self.m)).astype(np.int32), 1).astype(np.float64)
        m = np.atleast_1d(m).astype(np.float64, copy=False)
        if self.last_end is None or start != self.last_end:             # :183-186
            self.phase, self.integ = self.initial_phase, 0.0
        phase = np.mod(self.phase + np.cumsum(freq / self.sr), 1.0)     # :189-195
        P = self.sr / np.maximum(freq, 1.0)                             # :198
        theta = np.pi * phase
        sin_den = np.sin(theta)
        blit = np.where(np.abs(sin_den) < 1e-9, m / P, np.sin(m * theta) / (P * sin_den))  # :203-214
        blit_ac = blit - 1.0 / P                                        # :218
        saw, _ = lfilter(np.array([1.0]), np.array([1.0, -self.leak]), blit_ac,
                         zi=np.array([self.leak * self.integ]))         # :225-236
        self.phase, self.integ, self.last_end = phase[-1], saw[-1], start + duration   # :249-251
        return (saw * 2.0 * amp).astype(np.float32)                     # :256-262


def supersaw_params(frequency, voices=7, detune_cents=20.0, mix_mode="center_heavy", randomize_phase=True, seed=None):
    """super_saw_pe.py:128-232 -> (freqs f64[U], gains f64[U], phases f64[U])."""
    voices = max(int(voices), 1)
    if voices == 1 or detune_cents == 0:                                # :135-136
        ratios = np.array([1.0])
    else:
        ratios = 2 ** (np.linspace(-detune_cents, detune_cents, voices) / 1200.0)   # :138-139
    n = voices
    if n == 1:                                                          # :176-177
        gains = np.array([1.0])
    else:
        g = np.ones(n, dtype=np.float32)
        if mix_mode == "equal":
            pass
        elif mix_mode == "linear":                                      # :184-190
            d = np.abs(np.arange(n, dtype=np.float32) - (n - 1) / 2.0)
            g = 0.5 + 0.5 * (1.0 - d / np.max(d))
        elif mix_mode == "center_heavy":                                # :192-199
            g[:] = 0.5
            if n % 2 == 1:
                g[n // 2] = 1.0
            else:
                g[n // 2 - 1] = g[n // 2] = 1.0
        else:
            raise ValueError(f"Unknown mix mode: {mix_mode}")
        gains = g / np.sqrt(np.sum(g ** 2))                             # :206
    rng = np.random.default_rng(seed)                                   # :95
    freqs, phases, gs = [], [], []
    for i, ratio in enumerate(ratios):                                  # :219-231 (one rng draw per oscillator)
        freqs.append(float(frequency * ratio))
        gs.append(float(gains[i]))
        phases.append(float(rng.random(1)[0]) if randomize_phase else 0.0)
    return np.array(freqs), np.array(gs), np.array(phases)


class OracleSuperSaw:
    """super_saw_pe.py:69-318 with constant frequency / amplitude, mono."""

    def __init__(self, frequency, amplitude=1.0, voices=7, detune_cents=20.0, mix_mode="center_heavy",
                 randomize_phase=True, seed=None, sample_rate=44100):
        f, g, p = supersaw_params(frequency, voices, detune_cents, mix_mode, randomize_phase, seed)
        self.amp = float(amplitude)
        self.osc = [OracleBlitSaw(fi, gi, pi, sample_rate=sample_rate) for fi, gi, pi in zip(f, g, p)]

    def reset(self):
        for o in self.osc:
            o.reset()

    def render(self, start: int, duration: int, freq_ctl=None, amp_ctl=None) -> np.ndarray:
        """freq_ctl: the frequency PE's float32 output for the pull -- every oscillator then runs at
        GainPE(frequency_pe, ratio) = freq_ctl * float32(ratio) (super_saw_pe.py:236-240, gain_pe.py:123-125), construct
        the oracle with frequency=1.0 so that its oscillator frequencies ARE the ratios; amp_ctl: the amplitude PE's
        output, applied to the float64 sum (:287,300)."""
        result = np.zeros(duration, dtype=np.float64)                   # :292
        for o in self.osc:                                              # :295-297 (float32 snippets summed in f64)
            fc = None if freq_ctl is None else np.asarray(freq_ctl, np.float32) * np.float32(o.f)
            result += o.render(start, duration, freq_ctl=fc)
        amp = np.float64(self.amp) if amp_ctl is None else np.asarray(amp_ctl, np.float32).astype(np.float64)
        return (result * amp).astype(np.float32)                        # :300-303
