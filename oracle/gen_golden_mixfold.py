#!/usr/bin/env python
"""
Golden vectors for SURVEY.md §8f rank 3 (per-source DelayPE / GainPE / panning folded into the fused mix) and
for MixPE's extent gating of SpatialPE inputs, produced by the REAL reference (rdpoor/pygmu2 at /root/reference,
imported with oracle/stubs/).  Test infrastructure only.

    python oracle/gen_golden_mixfold.py     # rewrites tests/golden/mix_fold.npz

The graph is the shape of examples/27_spatial.py:223-234: sources of different lengths, spatialised by HRTF or
by a pan law, delayed and scaled, summed by one MixPE and pulled in 512-sample blocks.
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
SR = 44_100
LENGTHS = [3000, 4000, 2000, 5000, 700, 1500]
CHANNELS = [1, 2, 1, 1, 1, 1]


def sources(mod):
    out = []
    for i, (n, c) in enumerate(zip(LENGTHS, CHANNELS)):
        rng = np.random.default_rng(700 + i)
        x = rng.uniform(-1, 1, (n, c)).astype(np.float32) / 4
        out.append(mod.ArrayPE(x if c > 1 else x[:, 0]))
    return out


def graph(mod):
    s = sources(mod)
    return mod.MixPE(
        mod.DelayPE(mod.SpatialPE(s[0], method=mod.SpatialHRTF(30.0, 0.0)), 300),
        mod.GainPE(mod.SpatialPE(s[1], method=mod.SpatialHRTF(-100.0, 20.0)), 0.5),
        mod.DelayPE(mod.GainPE(mod.SpatialPE(s[2], method=mod.SpatialLinear(-45.0)), 0.8), 1000),
        mod.SpatialPE(s[3], method=mod.SpatialConstantPower(60.0)),
        mod.SpatialPE(s[4], method=mod.SpatialHRTF(170.0, -10.0)),      # ends at 700: its tail is cut by MixPE
        mod.DelayPE(mod.SpatialPE(s[5], method=mod.SpatialHRTF(0.0, 90.0)), 2500),
    )


def controls():
    """Per-sample control signals of the second graph: a tremolo gain and a pan sweep."""
    n = 6144
    t = np.arange(n) / SR
    gain = (0.6 + 0.4 * np.sin(2 * np.pi * 9.0 * t)).astype(np.float32)
    az = (80.0 * np.sin(2 * np.pi * 1.5 * t)).astype(np.float32)
    return gain, az


def graph_pe_controls(mod, **kw):
    """HRTF sources the bank can hold + one whose GainPE is PE-valued (gain_pe.py:105-121: the per-sample gain applies
    AFTER its HRTF convolution) + one panned by a PE-valued azimuth (spatial_pe.py:179-214): the un-bankable two."""
    s = sources(mod)
    gain, az = controls()
    return mod.MixPE(
        mod.DelayPE(mod.SpatialPE(s[0], method=mod.SpatialHRTF(30.0, 0.0)), 300),
        mod.GainPE(mod.SpatialPE(s[1], method=mod.SpatialHRTF(-100.0, 20.0)), 0.5),
        mod.GainPE(mod.SpatialPE(s[2], method=mod.SpatialHRTF(75.0, 10.0)), mod.ArrayPE(gain)),
        mod.SpatialPE(s[3], method=mod.SpatialLinear(mod.ArrayPE(az))),
        mod.SpatialPE(s[4], method=mod.SpatialHRTF(170.0, -10.0)),
        mod.DelayPE(mod.SpatialPE(s[5], method=mod.SpatialConstantPower(-20.0)), 2500),
        **kw)


def main():
    pg.set_sample_rate(SR)
    mix2 = graph_pe_controls(pg)
    out, pos = [], 0
    for d in [512] * 12:
        out.append(mix2.render(pos, d).data.copy())
        pos += d
    ext2 = mix2.extent()
    np.savez_compressed(os.path.join(GOLD, "mix_fold_pe_controls.npz"), y=np.concatenate(out), pulls=np.array([512] * 12),
                        extent=np.array([ext2.start, ext2.end]))
    print("wrote mix_fold_pe_controls.npz", np.concatenate(out).shape, "extent", ext2)
    mix = graph(pg)
    pulls = [512] * 12
    out, pos = [], 0
    for d in pulls:
        out.append(mix.render(pos, d).data.copy())
        pos += d
    y = np.concatenate(out)
    ext = mix.extent()
    np.savez_compressed(os.path.join(GOLD, "mix_fold.npz"), y=y, pulls=np.array(pulls), extent=np.array([ext.start, ext.end]))
    print("wrote mix_fold.npz", y.shape, "extent", ext, "max", float(np.abs(y).max()))


if __name__ == "__main__":
    main()
