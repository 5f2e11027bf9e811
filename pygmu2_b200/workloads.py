"""
Synthetic inputs for the five BASELINE.json configurations.

Pure numpy, seeded, no device code.  Shared by ``bench.py``, the parity tests
and ``oracle/gen_golden.py`` so that the CUDA path, the oracle and the real
reference all see bit-identical float32 inputs.

Every generator takes the *named* sizes as defaults and accepts smaller ones so
that tests can run the same shape family in seconds.
"""
from __future__ import annotations

import numpy as np

SR_441 = 44_100
SR_48 = 48_000


def _decaying_ir(rng, length: int, channels: int, tau: float) -> np.ndarray:
    """Gaussian noise under an exponential envelope, unit energy per channel, float32."""
    n = np.arange(length, dtype=np.float64)[:, None]
    ir = rng.standard_normal((length, channels)) * np.exp(-n / tau)
    ir /= np.sqrt(np.sum(ir * ir, axis=0, keepdims=True))
    return ir.astype(np.float32)


# -- C1: SinePE 440 Hz -> 4096-tap FIR, mono, 44.1 kHz -----------------------
def c1_fir(stream: int = 0, taps: int = 4096) -> np.ndarray:
    """(taps,) float32; stream 0 is the named config, stream i>0 the throughput copies."""
    rng = np.random.default_rng(1234 + stream)
    return (rng.standard_normal(taps) / 64.0).astype(np.float32)


def c1_sine(n_samples: int, stream: int = 0, n_streams: int = 4096, sr: int = SR_441) -> np.ndarray:
    """What the reference's SinePE(frequency=f) renders from sample 0 (sine_pe.py:119-157):
    float64 phase, amplitude 1, rounded to float32. f = 440*2^(stream/n_streams)."""
    f = 440.0 * 2.0 ** (stream / float(n_streams))
    time = np.arange(n_samples, dtype=np.float64) / sr          # sine_pe.py:173-174
    phase = 0.0 + 2.0 * np.pi * f * time                        # sine_pe.py:175
    return (1.0 * np.sin(phase)).astype(np.float32)             # sine_pe.py:146,157


# -- C2: stereo convolution reverb, 3 s IR (132 300 taps) @ 48 kHz, 512 pulls --
C2_L = 132_300
C2_PULL = 512


def c2_ir(length: int = C2_L, stream: int | None = None) -> np.ndarray:
    """(length, 2) float32.  stream=None is the shared IR (seed 3); stream=i the distinct-IR variant."""
    rng = np.random.default_rng(3 if stream is None else 30_000 + stream)
    return _decaying_ir(rng, length, 2, length / 6.0)


def c2_input(n_samples: int, stream: int = 0) -> np.ndarray:
    """(n_samples, 2) float32 uniform(-1, 1); stream 0 uses seed 2 as named."""
    rng = np.random.default_rng(2 if stream == 0 else 20_000 + stream)
    return rng.uniform(-1.0, 1.0, (n_samples, 2)).astype(np.float32)


# -- C3: 256 moving sources x HRTF -> stereo mix, 44.1 kHz, 512 pulls ---------
C3_SOURCES = 256
C3_PULL = 512


def c3_source(n_samples: int, source: int, n_sources: int = C3_SOURCES) -> np.ndarray:
    """(n_samples,) float32: uniform(-1,1)/n_sources, one rng stream per source."""
    rng = np.random.default_rng(40_000 + source)
    return (rng.uniform(-1.0, 1.0, n_samples) / float(n_sources)).astype(np.float32)


def c3_elevations(n_sources: int = C3_SOURCES) -> np.ndarray:
    rng = np.random.default_rng(6)
    return rng.choice(np.arange(-40, 91, 10), size=n_sources).astype(np.float64)


def c3_azimuth(source: int, pull_index: int, n_pulls_total: int, n_sources: int = C3_SOURCES) -> float:
    """Linear sweep -170 -> +170 degrees over the run, per-source phase offset, wrapped to [-180, 180)."""
    frac = pull_index / max(1, n_pulls_total - 1)
    az = -170.0 + 340.0 * frac + source * 360.0 / n_sources
    return float((az + 180.0) % 360.0 - 180.0)


def c3_synthetic_hrtf_table(taps: int = 512, n_entries: int = 368) -> np.ndarray:
    """(n_entries, taps, 2) float32 stand-in for the named 512-tap-per-ear shape
    (the shipped KEMAR set is 128-tap; SURVEY.md headline fact 4)."""
    rng = np.random.default_rng(5)
    n = np.arange(taps, dtype=np.float64)[None, :, None]
    return (rng.standard_normal((n_entries, taps, 2)) * np.exp(-n / 64.0) * 0.1).astype(np.float32)


# -- C4: 4096 streams x 2 s random IRs (88 200 taps), mixed ------------------
C4_L = 88_200
C4_PULL = 512
C4_STREAMS = 4096


def c4_input(n_samples: int, stream: int) -> np.ndarray:
    rng = np.random.default_rng(100 + stream)
    return rng.uniform(-1.0, 1.0, n_samples).astype(np.float32)


def c4_ir(stream: int, length: int = C4_L) -> np.ndarray:
    rng = np.random.default_rng(5000 + stream)
    return _decaying_ir(rng, length, 1, length / 6.0)[:, 0]


# -- C5: 1024-voice mix -> 10 s IR (441 000 taps) at 64-sample pulls ----------
C5_L = 441_000
C5_PULL = 64
C5_VOICES = 1024


def c5_voices(n_samples: int, n_voices: int = C5_VOICES) -> np.ndarray:
    """(n_voices, n_samples) float32 = uniform(-1,1)/32 (the §8d stand-in for SuperSawPE voices)."""
    rng = np.random.default_rng(7)
    return (rng.uniform(-1.0, 1.0, (n_voices, n_samples)) / 32.0).astype(np.float32)


def c5_ir(length: int = C5_L) -> np.ndarray:
    rng = np.random.default_rng(8)
    return _decaying_ir(rng, length, 1, length / 6.0)[:, 0]


# -- roofline bookkeeping (SURVEY.md §8d "Algorithmic bytes") -----------------
def bytes_per_block_step(n_streams: int, c_in: int, c_out: int, L: int, B: int,
                         distinct_filters: bool) -> int:
    """bytes_step = N*C_in*P*K*8 + [distinct ? N : 0]*C_out*P*K*8 + N*C_in*K*8 + N*(C_in+C_out)*B*4
    with K = B+1 bins and P = ceil(L/B) partitions."""
    K = B + 1
    P = -(-L // B)
    b = n_streams * c_in * P * K * 8
    if distinct_filters:
        b += n_streams * c_out * P * K * 8
    b += n_streams * c_in * K * 8
    b += n_streams * (c_in + c_out) * B * 4
    return int(b)
