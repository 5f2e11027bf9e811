"""
The oracle port (oracle/pygmu2_oracle.py) against the REAL reference package wherever oracle/_ref is present
(built by oracle/build_ref.py in the build container; travels to the GPU box, git-ignored).  Complements
tests/test_oracle_golden.py, which pins the port to committed outputs of the same reference.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build_ref  # noqa: E402
import pygmu2_oracle as orc  # noqa: E402

ref = build_ref.import_ref()
pytestmark = pytest.mark.skipif(ref is None, reason="oracle/_ref not built (python oracle/build_ref.py)")


def _ref_pulls(make_pe, pulls, sr=48_000):
    ref.set_sample_rate(sr)
    pe = make_pe()
    out, pos = [], 0
    with ref.NullRenderer(sample_rate=sr) as r:
        r.set_source(pe)
        r.start()
        for d in pulls:
            out.append(pe.render(pos, d).data.copy())
            pos += d
    return np.concatenate(out)


@pytest.mark.parametrize("L,c_src,c_f", [(300, 1, 1), (2048, 2, 1), (1000, 1, 2), (4097, 2, 2)])
def test_oracle_convolve_is_the_reference_bit_for_bit(L, c_src, c_f):
    rng = np.random.default_rng(L)
    x = rng.uniform(-1, 1, (3000, c_src)).astype(np.float32)
    h = (rng.standard_normal((L, c_f)) / np.sqrt(L)).astype(np.float32)
    pulls = (512, 17, 700, 1, 1770)
    y_ref = _ref_pulls(lambda: ref.ConvolvePE(ref.ArrayPE(x), ref.ArrayPE(h)), pulls)
    conv = orc.OracleConvolve(h, c_src)
    pos, ys = 0, []
    for d in pulls:
        ys.append(conv.render(x[pos:pos + d]))
        pos += d
    assert np.array_equal(np.concatenate(ys), y_ref)


def test_oracle_mix_is_the_reference_bit_for_bit():
    rng = np.random.default_rng(1)
    arrs = [rng.standard_normal((400, 2)).astype(np.float32) for _ in range(9)]
    y_ref = _ref_pulls(lambda: ref.MixPE(*[ref.ArrayPE(a) for a in arrs]), (400,))
    assert np.array_equal(orc.oracle_mix(arrs), y_ref)
