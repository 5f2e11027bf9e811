#!/usr/bin/env python
"""
Golden vectors for the device-resident sources, produced by the REAL reference
(rdpoor/pygmu2 at /root/reference, imported with oracle/stubs/).  Test infrastructure only.

    python oracle/gen_golden_sources.py     # rewrites tests/golden/src_*.npz
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
PULLS = [64] * 6 + [1, 63, 17, 500, 128, 1000, 7, 64, 64]          # ragged, contiguous
SR = 44_100


def pull(pe, pulls, start=0):
    out, pos = [], start
    for d in pulls:
        out.append(pe.render(pos, d).data.copy())
        pos += d
    return np.concatenate(out, axis=0)


def main():
    pg.set_sample_rate(SR)
    out = {"pulls": np.array(PULLS, np.int64)}
    # SinePE: constant parameters, incl. a late start (large sample index) and stereo replication
    out["sine_440"] = pull(pg.SinePE(frequency=440.0), PULLS)
    out["sine_params"] = np.array([[440.0, 1.0, 0.0], [1234.5, 0.25, 0.7], [20.0, 2.0, -1.0]])
    out["sine_b"] = pull(pg.SinePE(frequency=1234.5, amplitude=0.25, phase=0.7, channels=2), PULLS, start=4_000_000)
    out["sine_c"] = pull(pg.SinePE(frequency=20.0, amplitude=2.0, phase=-1.0), PULLS, start=-300)
    # BlitSawPE: auto M, fixed M, low frequency (< 1 Hz clamp), initial phase, leak; restart after a gap
    cases = [(440.0, 1.0, 0.0, None, 0.999), (55.0, 0.5, 0.3, None, 0.999), (3000.0, 1.0, 0.9, None, 0.995),
             (220.0, 0.8, 0.25, 20, 0.999), (0.5, 1.0, 0.0, None, 0.999), (10000.0, 1.0, 0.5, None, 0.999)]
    out["blit_cases"] = np.array([[f, a, p, -1 if m is None else m, lk] for f, a, p, m, lk in cases])
    for i, (f, a, p, m, lk) in enumerate(cases):
        pe = pg.BlitSawPE(frequency=f, amplitude=a, initial_phase=p, m=m, leak=lk)
        y = pull(pe, PULLS)
        y2 = pull(pe, [64, 64], start=10_000)          # non-contiguous pull: state resets (blit_saw_pe.py:183-186)
        out[f"blit_{i}"] = np.concatenate([y, y2])
    # SuperSawPE: seeds, voice counts, mix modes
    scases = [(440.0, 1.0, 7, 20.0, "center_heavy", True, 0), (110.0, 0.5, 5, 35.0, "linear", True, 1),
              (880.0, 1.0, 4, 10.0, "equal", True, 2), (330.0, 0.7, 1, 20.0, "center_heavy", True, 3),
              (250.0, 1.0, 7, 0.0, "center_heavy", False, 4), (1000.0, 0.3, 9, 50.0, "center_heavy", True, 5)]
    out["ssaw_cases"] = np.array([[f, a, v, d, {"center_heavy": 0, "linear": 1, "equal": 2}[mm], int(rp), sd]
                                  for f, a, v, d, mm, rp, sd in scases])
    for i, (f, a, v, d, mm, rp, sd) in enumerate(scases):
        pe = pg.SuperSawPE(frequency=f, amplitude=a, voices=v, detune_cents=d, mix_mode=mm,
                           randomize_phase=rp, seed=sd)
        out[f"ssaw_{i}"] = pull(pe, PULLS)
    # C5 front end in miniature: MixPE of 16 SuperSaw voices, 64-sample pulls
    voices = [pg.SuperSawPE(frequency=110.0 * 2 ** (i / 12.0), amplitude=1.0 / 16, seed=i) for i in range(16)]
    out["c5_voicemix16"] = pull(pg.MixPE(*voices), [64] * 12)
    np.savez_compressed(os.path.join(GOLD, "src_oscillators.npz"), **out)
    print("wrote", os.path.join(GOLD, "src_oscillators.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
