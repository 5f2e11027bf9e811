#!/usr/bin/env python
"""
Golden vectors for the interpolated modes of DelayPE (fractional and PE-valued delay, linear and cubic; delay_pe.py:163-228,
interpolated_lookup.py:89-145), produced by the REAL reference (rdpoor/pygmu2 at /root/reference, imported with
oracle/stubs/).  Test infrastructure only.

    python oracle/gen_golden_delay.py      # rewrites tests/golden/delay_interp.npz

Cases: a stereo array delayed by 10.5 and by -3.25 samples (the second reads ahead of the source), a vibrato (a sine-shaped
delay control around 100 samples), each linear and cubic, pulled in ragged chunks from before the source starts to
after it ends, plus the extents the reference reports.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("PYGMU2_REFERENCE", "/root/reference")

PULLS = (100, 37, 512, 1, 255, 1024, 700, 300, 171)
START = -64


def inputs():
    rng = np.random.default_rng(4242)
    x = rng.uniform(-1, 1, (2000, 2)).astype(np.float32)
    n = sum(PULLS)
    t = np.arange(n + 256)
    ctl = (100.0 + 50.0 * np.sin(2 * np.pi * t / 333.0)).astype(np.float32)
    return x, ctl


def cases(mod, interp_of):
    """name -> PE, for a module that looks like pygmu2 (the reference, or pygmu2_b200)."""
    x, ctl = inputs()
    out = {}
    for iname in ("linear", "cubic"):
        im = interp_of(iname)
        out[f"float_10p5_{iname}"] = mod.DelayPE(mod.ArrayPE(x), 10.5, interpolation=im)
        out[f"float_m3p25_{iname}"] = mod.DelayPE(mod.ArrayPE(x), -3.25, interpolation=im)
        out[f"vibrato_{iname}"] = mod.DelayPE(mod.ArrayPE(x), mod.ArrayPE(ctl), interpolation=im)
    return out


def pull(pe):
    pos, ys = START, []
    for d in PULLS:
        ys.append(pe.render(pos, d).data.copy())
        pos += d
    return np.concatenate(ys, axis=0)


def main():
    sys.path.insert(0, os.path.join(HERE, "stubs"))      # (only here: the tests import this module for cases() / pull())
    sys.path.insert(0, os.path.join(REF, "src"))
    import pygmu2 as ref  # noqa: E402  (the real reference)
    from pygmu2.wavetable_pe import InterpolationMode
    ref.set_sample_rate(44_100)
    blob = {}
    for name, pe in cases(ref, lambda n: InterpolationMode(n)).items():
        blob[name] = pull(pe)
        e = pe.extent()
        blob[name + "_extent"] = np.array([e.start, e.end], dtype=np.int64)
    path = os.path.join(ROOT, "tests", "golden", "delay_interp.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, {k: v.shape for k, v in blob.items() if not k.endswith("_extent")})


if __name__ == "__main__":
    main()
