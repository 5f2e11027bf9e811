#!/usr/bin/env python
"""
Golden vectors for the MODULATED sources (PE-valued parameters), produced by the REAL reference (rdpoor/pygmu2 at
/root/reference, imported with oracle/stubs/).  Test infrastructure only.

    python oracle/gen_golden_modulated.py     # rewrites tests/golden/src_modulated.npz

Cases (sine_pe.py:134-232, the stateful branch): FM (frequency = 440 + 50 Hz LFO), AM (amplitude = tremolo PE),
PM (phase = PE), all three at once, and a restart through on_stop/on_start.  The control signals are stored with the
outputs so that the tests feed the device PE the very same float32 control vectors.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("PYGMU2_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "stubs"))
sys.path.insert(0, os.path.join(REF, "src"))

import pygmu2 as pg  # noqa: E402  (the real reference)

GOLD = os.path.join(ROOT, "tests", "golden")
PULLS = [64] * 4 + [1, 63, 17, 500, 128, 1000, 7, 512, 512]
SR = 44_100


def pull(pe, pulls, start=0):
    out, pos = [], start
    for d in pulls:
        out.append(pe.render(pos, d).data.copy())
        pos += d
    return np.concatenate(out, axis=0)


def main():
    pg.set_sample_rate(SR)
    n = sum(PULLS)
    rng = np.random.default_rng(99)
    t = np.arange(n) / SR
    ctl = {
        "fm_freq": (440.0 + 50.0 * np.sin(2 * np.pi * 5.0 * t)).astype(np.float32),
        "am_amp": (0.6 + 0.4 * np.sin(2 * np.pi * 3.0 * t)).astype(np.float32),
        "pm_phase": (2.0 * np.sin(2 * np.pi * 110.0 * t) + 0.01 * rng.standard_normal(n)).astype(np.float32),
        "sweep_freq": np.geomspace(20.0, 12000.0, n).astype(np.float32),
    }
    out = {"pulls": np.array(PULLS, np.int64), **{f"ctl_{k}": v for k, v in ctl.items()}}
    A = lambda k: pg.ArrayPE(ctl[k])  # noqa: E731
    out["fm"] = pull(pg.SinePE(frequency=A("fm_freq")), PULLS)
    out["fm_phase0p7_amp0p5"] = pull(pg.SinePE(frequency=A("fm_freq"), amplitude=0.5, phase=0.7), PULLS)
    out["am"] = pull(pg.SinePE(frequency=1000.0, amplitude=A("am_amp"), phase=0.25), PULLS)
    out["pm"] = pull(pg.SinePE(frequency=220.0, phase=A("pm_phase")), PULLS)
    out["all3_stereo"] = pull(pg.SinePE(frequency=A("sweep_freq"), amplitude=A("am_amp"), phase=A("pm_phase"), channels=2), PULLS)
    # restart: stop/start resets the accumulated phase (sine_pe.py:110-118)
    pe = pg.SinePE(frequency=A("fm_freq"), phase=A("pm_phase"))
    r = pg.NullRenderer(sample_rate=SR)
    r.set_source(pe)
    r.start()
    a = pull(pe, [256, 256])
    r.stop()
    r.start()
    b = pull(pe, [256, 256])
    r.stop()
    out["restart"] = np.concatenate([a, b])
    # ---- BLIT oscillators with PE-valued parameters (blit_saw_pe.py:161-262, super_saw_pe.py:223-246,287-303)
    ctl2 = {
        "vib_freq": (220.0 * 2.0 ** (0.5 * np.sin(2 * np.pi * 6.0 * t) / 12.0)).astype(np.float32),   # +-50 cent vibrato
        "glide_freq": np.geomspace(55.0, 3520.0, n).astype(np.float32),                                 # harmonic count changes
        "env_amp": np.minimum(1.0, np.arange(n) / 800.0).astype(np.float32) * 0.8,
    }
    out.update({f"ctl_{k}": v for k, v in ctl2.items()})
    B = lambda k: pg.ArrayPE(ctl2[k])  # noqa: E731
    out["blit_vib"] = pull(pg.BlitSawPE(frequency=B("vib_freq"), amplitude=0.7, initial_phase=0.3), PULLS)
    out["blit_glide_env"] = pull(pg.BlitSawPE(frequency=B("glide_freq"), amplitude=B("env_amp")), PULLS)
    out["blit_amp_only_m12"] = pull(pg.BlitSawPE(frequency=330.0, amplitude=B("env_amp"), m=12, leak=0.995), PULLS)
    out["ssaw_vib"] = pull(pg.SuperSawPE(frequency=B("vib_freq"), amplitude=0.5, seed=3), PULLS)
    out["ssaw_glide_env_stereo"] = pull(pg.SuperSawPE(frequency=B("glide_freq"), amplitude=B("env_amp"), voices=5,
                                                     detune_cents=35.0, mix_mode="linear", channels=2, seed=4), PULLS)
    out["ssaw_amp_only"] = pull(pg.SuperSawPE(frequency=110.0, amplitude=B("env_amp"), seed=5), PULLS)
    ctl2["m_steps"] = np.repeat(np.array([1, 3, 7.9, 0.2, 15, 40], np.float32), n // 6 + 1)[:n]   # truncation, clamp to >= 1
    out["ctl_m_steps"] = ctl2["m_steps"]
    out["blit_m_pe"] = pull(pg.BlitSawPE(frequency=B("vib_freq"), amplitude=0.6, m=B("m_steps")), PULLS)
    pe = pg.BlitSawPE(frequency=B("vib_freq"), amplitude=B("env_amp"))
    a = pull(pe, [256, 256])
    b = pull(pe, [256, 256], start=1024)            # non-contiguous: state resets (blit_saw_pe.py:183-186)
    out["blit_gap"] = np.concatenate([a, b])
    np.savez_compressed(os.path.join(GOLD, "src_modulated.npz"), **out)
    print("wrote src_modulated.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
