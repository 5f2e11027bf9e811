"""
Pins the oracle (oracle/pygmu2_oracle.py) before anything trusts it:
  (a) the reference's own known-answer tests for this path, restated on the oracle
      (reference tests/test_convolve_pe.py:49-162,165-185; tests/test_mix_pe.py:73-152;
       tests/test_spatial_pe.py:417-451);
  (b) golden outputs of the REAL reference (tests/golden/*.npz, made by oracle/gen_golden.py).
CPU only.
"""
import numpy as np
import pytest

import pygmu2_oracle as orc
from pygmu2_b200 import workloads as wl
from conftest import golden


def _pulls(conv, x, pulls):
    out, pos = [], 0
    for d in pulls:
        seg = x[pos:pos + d]
        if seg.shape[0] < d:  # past the source extent: zeros (array_pe.py:94-111)
            seg = np.concatenate([seg, np.zeros((d - seg.shape[0],) + seg.shape[1:], np.float32)])
        out.append(conv.render(seg))
        pos += d
    return np.concatenate(out, axis=0)


# ---- (a) reference known-answer tests on the oracle -------------------------
def test_matches_numpy_convolve_mono():
    x = np.array([1, 2, 3, 4], dtype=np.float32)
    h = np.array([1, 0.5, -1], dtype=np.float32)
    y = orc.OracleConvolve(h, 1, fft_size=16).render(np.concatenate([x, np.zeros(2, np.float32)]))[:, 0]
    np.testing.assert_allclose(y, np.convolve(x, h, mode="full"), atol=1e-5, rtol=0)
    np.testing.assert_allclose(orc.direct_convolve_f64(x, h), np.convolve(x, h), atol=1e-6)


def test_dirac_identity():
    x = np.random.default_rng(0).normal(size=64).astype(np.float32)
    y = orc.OracleConvolve([1.0], 1, fft_size=64).render(x)[:, 0]
    np.testing.assert_allclose(y, x, atol=1e-6, rtol=0)


def test_fft_size_too_small():
    with pytest.raises(ValueError):
        orc.OracleConvolve(np.ones(20, np.float32), 1, fft_size=16)


def test_channel_mismatch():
    with pytest.raises(ValueError):
        orc.OracleConvolve(np.ones((4, 3), np.float32), 2)


def test_ir_energy_norm():
    assert orc.ir_energy_norm([3.0, 4.0]) == pytest.approx(5.0, rel=1e-5)
    assert orc.ir_energy_norm([1.0]) == pytest.approx(1.0)
    assert orc.ir_energy_norm([0.0, 0.0]) == 1.0
    assert orc.ir_energy_norm(np.ones((2, 2), np.float32)) == pytest.approx(2.0)


def test_mix_constants():
    a = np.full((10, 1), 0.3, np.float32)
    b = np.full((10, 1), 0.4, np.float32)
    np.testing.assert_allclose(orc.oracle_mix([a, b]), 0.7, atol=1e-6)
    np.testing.assert_allclose(orc.oracle_mix([a * 0 + 0.8, a * 0 + 0.8]), 1.6, atol=1e-6)  # no clipping


def test_hrtf_filename_cases():
    ent = orc.kemar_entries()
    assert len(ent) == 368
    assert ent[orc.hrtf_nearest_index(0, 0)][2] == "H0e000a.wav"
    assert ent[orc.hrtf_nearest_index(45, 0)][2] == "H0e045a.wav"
    assert orc.hrtf_nearest_index(-45, 0) == orc.hrtf_nearest_index(45, 0)
    assert ent[orc.hrtf_nearest_index(90, 0)][2] == "H0e090a.wav"
    assert ent[orc.hrtf_nearest_index(0, 30)][2].startswith("H30")


# ---- (b) golden vectors from the real reference -----------------------------
def test_golden_unit():
    g = golden("convolve_unit.npz")
    x = np.array([1, 2, 3, 4], dtype=np.float32)
    pad = lambda a, n: np.concatenate([a, np.zeros((n - a.shape[0],) + a.shape[1:], np.float32)])
    y = orc.OracleConvolve(np.array([1, 0.5, -1], np.float32), 1, 16).render(pad(x, 6))
    assert np.array_equal(y, g["mono_small"])
    xs = np.array([[1, 10], [2, 20], [3, 30], [4, 40]], dtype=np.float32)
    assert np.array_equal(orc.OracleConvolve(np.array([1, -1], np.float32), 2, 16).render(pad(xs, 5)),
                          g["stereo_monofilter"])
    h2 = np.stack([np.array([1.0, 0.5], np.float32), np.array([-1.0, 0.5], np.float32)], axis=1)
    assert np.array_equal(orc.OracleConvolve(h2, 1, 16).render(pad(x, 5)), g["fanout"])
    xr = np.random.default_rng(0).normal(size=200).astype(np.float32)
    c = orc.OracleConvolve(np.array([0.25, 0.5, 0.25], np.float32), 1, 64)
    assert np.array_equal(_pulls(c, xr.reshape(-1, 1), (17, 23, 19, 41, 7, 93, 2)), g["chunked"])


def test_golden_ragged_and_reset():
    g = golden("convolve_ragged.npz")
    c = orc.OracleConvolve(g["h"], 2)
    ya = _pulls(c, g["x"], tuple(g["pulls_a"]))
    assert np.array_equal(ya, g["ya"])
    c.reset()
    s = int(g["start_b"])
    yb = _pulls(c, g["x"][s:], tuple(g["pulls_b"]))
    assert np.array_equal(yb, g["yb"])


def test_golden_c1():
    g = golden("c1_sine_fir4096.npz")
    n = int(g["n"])
    y = orc.OracleConvolve(wl.c1_fir(), 1).render(wl.c1_sine(n))
    assert np.array_equal(y, g["y"])
    # and the independent direct-form contract
    d = orc.direct_convolve_f64(wl.c1_sine(n), wl.c1_fir())[:n]
    assert np.max(np.abs(d - g["y"][:, 0])) <= 1e-6 * np.max(np.abs(d))


def test_golden_c2():
    g = golden("c2_stereo_reverb.npz")
    n_pulls = int(g["n_pulls"])
    c = orc.OracleConvolve(wl.c2_ir(), 2)
    y = _pulls(c, wl.c2_input(n_pulls * wl.C2_PULL), (wl.C2_PULL,) * n_pulls)
    assert np.array_equal(y, g["y"])


def _kemar_table():
    import os
    from conftest import ROOT
    z = np.load(os.path.join(ROOT, "pygmu2_b200", "assets", "kemar_compact_i16.npz"))
    return z["ir_i16"].astype(np.float32) / 32768.0


def test_golden_hrtf_lookup():
    g = golden("hrtf_lookup.npz")
    ent = orc.kemar_entries()
    idx = np.array([orc.hrtf_nearest_index(a, e, ent) for a, e in zip(g["az"], g["el"])])
    assert np.array_equal(idx, g["idx"])


def test_golden_c3():
    g = golden("c3_hrtf_mix.npz")
    table = _kemar_table()
    ns, npull = int(g["n_sources"]), int(g["n_pulls"])
    n = npull * wl.C3_PULL
    el = wl.c3_elevations()[:ns]
    hs = [orc.OracleHRTF(table, wl.c3_azimuth(s, 0, npull, ns), float(el[s])) for s in range(ns)]
    xs = [wl.c3_source(n, s, ns) for s in range(ns)]
    outs = []
    for b in range(npull):
        per = []
        for s, h in enumerate(hs):
            h.azimuth = wl.c3_azimuth(s, b, npull, ns)
            per.append(h.render(xs[s][b * 512:(b + 1) * 512], b * 512))
        outs.append(orc.oracle_mix(per))
    assert np.array_equal(np.concatenate(outs), g["y"])
    # single source, ragged pulls, swap, reset
    h = orc.OracleHRTF(table, -37.0, 12.0)
    x = wl.c3_source(4000, 3, 1)
    segs, pos = [], 0
    for i, d in enumerate(g["single_pulls"]):
        if i == 3:
            h.azimuth = 100.0
        if i == 5:
            h.elevation = -35.0
        segs.append(h.render(x[pos:pos + d], pos))
        pos += int(d)
    segs.append(h.render(x[3000:3400], 3000))
    assert np.array_equal(np.concatenate(segs), g["single"])
    st = np.stack([wl.c3_source(1500, 1, 1), wl.c3_source(1500, 2, 1)], axis=1)
    h2 = orc.OracleHRTF(table, 60.0, -20.0)
    y = np.concatenate([h2.render(st[0:512], 0), h2.render(st[512:1024], 512), h2.render(st[1024:1500], 1024)])
    assert np.array_equal(y, g["stereo_src"])


def test_golden_c4():
    g = golden("c4_streams_mix.npz")
    ns, npull, L = int(g["n_streams"]), int(g["n_pulls"]), int(g["L"])
    n = npull * wl.C4_PULL
    per = []
    for s in range(ns):
        c = orc.OracleConvolve(wl.c4_ir(s, L), 1)
        per.append(_pulls(c, wl.c4_input(n, s).reshape(-1, 1), (wl.C4_PULL,) * npull))
    assert np.array_equal(np.stack([p[:, 0] for p in per]), g["per_stream"])
    assert np.array_equal(orc.oracle_mix(per), g["mix"])


def test_golden_c5():
    g = golden("c5_voicebank_longir.npz")
    nv, npull, L = int(g["n_voices"]), int(g["n_pulls"]), int(g["L"])
    n = npull * wl.C5_PULL
    v = wl.c5_voices(n, nv)
    vm = orc.oracle_mix([v[i].reshape(-1, 1) for i in range(nv)])
    assert np.array_equal(vm, g["voice_mix"])
    c = orc.OracleConvolve(wl.c5_ir(L), 1)
    assert np.array_equal(_pulls(c, vm, (wl.C5_PULL,) * npull), g["y"])
