"""
Parity of the CUDA path (through the C ABI: ctypes -> libpgx.so) against the oracle and
against golden outputs of the real reference.  Needs a B200: run with ``-m gpu``.

Tolerance (BASELINE.json north_star): max-abs error <= 1e-5 of full scale of the reference
output, fp32 device arithmetic vs the reference's float64 (ConvolvePE) / float32 (HRTF).
MixPE's K5 sum is bit-exact.
"""
import numpy as np
import pytest

import pygmu2_b200 as pg
import pygmu2_oracle as orc
from pygmu2_b200 import workloads as wl
from conftest import golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _pull_pe(pe, pulls, start=0):
    out, pos = [], start
    for d in pulls:
        out.append(pe.render(pos, int(d)).data.copy())
        pos += int(d)
    return np.concatenate(out, axis=0)


def _oracle_pulls(conv, x, pulls):
    out, pos = [], 0
    for d in pulls:
        d = int(d)
        seg = x[pos:pos + d]
        if seg.shape[0] < d:
            seg = np.concatenate([seg, np.zeros((d - seg.shape[0],) + seg.shape[1:], np.float32)])
        out.append(conv.render(seg))
        pos += d
    return np.concatenate(out, axis=0)


# ---------------------------------------------------------------------------
# the reference's own hot-path tests, run against the device PEs
# (reference tests/test_convolve_pe.py:49-162)
class TestReferenceConvolveCases:
    def setup_method(self):
        self.renderer = pg.NullRenderer(sample_rate=10_000)

    def test_matches_numpy_convolve_mono(self):
        x = np.array([1, 2, 3, 4], dtype=np.float32)
        h = np.array([1, 0.5, -1], dtype=np.float32)
        pe = pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(h), fft_size=16)
        self.renderer.set_source(pe)
        y_expected = np.convolve(x, h, mode="full").astype(np.float32)
        y = pe.render(0, len(y_expected)).data[:, 0]
        np.testing.assert_allclose(y, y_expected, atol=1e-5, rtol=0.0)

    def test_dirac_impulse_is_identity(self):
        x = np.random.default_rng(0).normal(size=64).astype(np.float32)
        pe = pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE([1.0]), fft_size=64)
        self.renderer.set_source(pe)
        y = pe.render(0, len(x)).data[:, 0]
        np.testing.assert_allclose(y, x, atol=1e-6, rtol=0.0)

    def test_filter_mono_applies_to_all_channels(self):
        x = np.array([[1, 10], [2, 20], [3, 30], [4, 40]], dtype=np.float32)
        h = np.array([1, -1], dtype=np.float32)
        pe = pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(h), fft_size=16)
        self.renderer.set_source(pe)
        y = pe.render(0, 5).data
        np.testing.assert_allclose(y[:, 0], np.convolve(x[:, 0], h), atol=1e-5, rtol=0.0)
        np.testing.assert_allclose(y[:, 1], np.convolve(x[:, 1], h), atol=1e-5, rtol=0.0)

    def test_mono_src_multi_channel_filter_fans_out(self):
        x = np.array([1.0, 2.0, 3.0, 4.0], dtype=np.float32)
        h = np.stack([np.array([1.0, 0.5], np.float32), np.array([-1.0, 0.5], np.float32)], axis=1)
        pe = pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(h), fft_size=16)
        self.renderer.set_source(pe)
        y = pe.render(0, 5).data
        assert y.shape[1] == 2
        np.testing.assert_allclose(y[:, 0], np.convolve(x, h[:, 0]), atol=1e-5, rtol=0.0)
        np.testing.assert_allclose(y[:, 1], np.convolve(x, h[:, 1]), atol=1e-5, rtol=0.0)

    def test_chunked_render_matches_full(self):
        x = np.random.default_rng(0).normal(size=200).astype(np.float32)
        h = np.array([0.25, 0.5, 0.25], dtype=np.float32)
        full = pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(h), fft_size=64)
        self.renderer.set_source(full)
        total = len(x) + len(h) - 1
        y_full = full.render(0, total).data[:, 0]
        chunked = pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(h), fft_size=64)
        self.renderer.set_source(chunked)
        y_chunk = _pull_pe(chunked, (17, 23, 19, 41, 7, 93, 2))[:, 0]
        np.testing.assert_allclose(y_chunk[:total], y_full, atol=1e-5, rtol=0.0)

    def test_fft_size_smaller_than_filter_raises(self):
        pe = pg.ConvolvePE(pg.ArrayPE(np.ones(8)), pg.ArrayPE(np.ones(20)), fft_size=16)
        with pytest.raises(ValueError):
            pe.render(0, 4)

    def test_channel_mismatch_raises(self):
        pe = pg.ConvolvePE(pg.ArrayPE(np.ones((8, 2))), pg.ArrayPE(np.ones((4, 3))))
        with pytest.raises(ValueError):
            pe.render(0, 4)

    def test_default_fft_size_reported(self):
        pe = pg.ConvolvePE(pg.ArrayPE(np.ones(8)), pg.ArrayPE(np.ones(3000)))
        pe.render(0, 4)
        assert pe.fft_size == 4096

    def test_restart_after_stop_works(self):
        # the reference asserts here; the device PE simply starts a new run
        x = np.arange(1, 9, dtype=np.float32)
        pe = pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE([1.0, 1.0]))
        r = pg.NullRenderer(sample_rate=10_000)
        r.set_source(pe)
        r.start()
        a = pe.render(0, 8).data.copy()
        r.stop()
        r.start()
        b = pe.render(0, 8).data
        np.testing.assert_allclose(a, b, atol=1e-6)


# ---------------------------------------------------------------------------
# golden outputs of the real reference
def test_golden_unit_vectors():
    g = golden("convolve_unit.npz")
    pg.set_sample_rate(10_000)
    x = np.array([1, 2, 3, 4], dtype=np.float32)
    y = pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(np.array([1, 0.5, -1], np.float32)), fft_size=16).render(0, 6).data
    np.testing.assert_allclose(y, g["mono_small"], atol=1e-5)
    xs = np.array([[1, 10], [2, 20], [3, 30], [4, 40]], dtype=np.float32)
    y = pg.ConvolvePE(pg.ArrayPE(xs), pg.ArrayPE(np.array([1, -1], np.float32)), fft_size=16).render(0, 5).data
    np.testing.assert_allclose(y, g["stereo_monofilter"], atol=1e-5)
    h2 = np.stack([np.array([1.0, 0.5], np.float32), np.array([-1.0, 0.5], np.float32)], axis=1)
    y = pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(h2), fft_size=16).render(0, 5).data
    np.testing.assert_allclose(y, g["fanout"], atol=1e-5)
    xr = np.random.default_rng(0).normal(size=200).astype(np.float32)
    pe = pg.ConvolvePE(pg.ArrayPE(xr), pg.ArrayPE(np.array([0.25, 0.5, 0.25], np.float32)), fft_size=64)
    np.testing.assert_allclose(_pull_pe(pe, (17, 23, 19, 41, 7, 93, 2)), g["chunked"], atol=1e-5)


@pytest.mark.parametrize("block", [None, 16, 64, 256, 1024])
def test_golden_ragged_pulls_and_reset(block):
    g = golden("convolve_ragged.npz")
    pe = pg.ConvolvePE(pg.ArrayPE(g["x"]), pg.ArrayPE(g["h"]), block_size=block)
    ya = _pull_pe(pe, g["pulls_a"])
    assert rel_err(ya, g["ya"]) <= TOL
    yb = _pull_pe(pe, g["pulls_b"], start=int(g["start_b"]))  # jump back: history must be cleared
    assert rel_err(yb, g["yb"]) <= TOL


def test_golden_c1_sine_fir4096():
    g = golden("c1_sine_fir4096.npz")
    n = int(g["n"])
    pe = pg.ConvolvePE(pg.SinePE(frequency=440.0), pg.ArrayPE(wl.c1_fir()))
    with pg.NullRenderer(sample_rate=wl.SR_441) as r:
        r.set_source(pe)
        r.start()
        y = pe.render(0, n).data
    assert pe.fft_size == 4096
    assert rel_err(y, g["y"]) <= TOL


def test_golden_c2_stereo_reverb():
    g = golden("c2_stereo_reverb.npz")
    pg.set_sample_rate(wl.SR_48)
    npull = int(g["n_pulls"])
    pe = pg.ConvolvePE(pg.ArrayPE(wl.c2_input(npull * wl.C2_PULL)), pg.ArrayPE(wl.c2_ir()))
    y = _pull_pe(pe, (wl.C2_PULL,) * npull)
    assert pe.bank.block == 512 and pe.bank.partitions == 259
    assert rel_err(y, g["y"]) <= TOL


def test_golden_c3_hrtf_mix_fused_and_per_pe():
    g = golden("c3_hrtf_mix.npz")
    ns, npull = int(g["n_sources"]), int(g["n_pulls"])
    n = npull * wl.C3_PULL
    el = wl.c3_elevations()[:ns]
    for fuse in (True, False):
        methods = [pg.SpatialHRTF(azimuth=wl.c3_azimuth(s, 0, npull, ns), elevation=float(el[s])) for s in range(ns)]
        pes = [pg.SpatialPE(pg.ArrayPE(wl.c3_source(n, s, ns)), method=m) for s, m in enumerate(methods)]
        mix = pg.MixPE(*pes, fuse=fuse)
        outs = []
        for b in range(npull):
            for s, m in enumerate(methods):
                m.azimuth = wl.c3_azimuth(s, b, npull, ns)
            outs.append(mix.render(b * wl.C3_PULL, wl.C3_PULL).data.copy())
        assert rel_err(np.concatenate(outs), g["y"]) <= TOL, f"fuse={fuse}"


def test_golden_c3_single_source_ragged_swap_reset():
    g = golden("c3_hrtf_mix.npz")
    m = pg.SpatialHRTF(azimuth=-37.0, elevation=12.0)
    sp = pg.SpatialPE(pg.ArrayPE(wl.c3_source(4000, 3, 1)), method=m)
    segs, pos = [], 0
    for i, d in enumerate(g["single_pulls"]):
        if i == 3:
            m.azimuth = 100.0
        if i == 5:
            m.elevation = -35.0
        segs.append(sp.render(pos, int(d)).data.copy())
        pos += int(d)
    segs.append(sp.render(3000, 400).data.copy())
    assert rel_err(np.concatenate(segs), g["single"]) <= TOL
    st = np.stack([wl.c3_source(1500, 1, 1), wl.c3_source(1500, 2, 1)], axis=1)
    sp2 = pg.SpatialPE(pg.ArrayPE(st), method=pg.SpatialHRTF(azimuth=60.0, elevation=-20.0))
    assert rel_err(_pull_pe(sp2, (512, 512, 476)), g["stereo_src"]) <= TOL


def test_hrtf_sample_rate_mismatch_raises_in_strict():
    pg.set_sample_rate(48_000)
    sp = pg.SpatialPE(pg.ArrayPE(np.ones(64)), method=pg.SpatialHRTF(azimuth=10.0))
    with pytest.raises(RuntimeError):
        sp.render(0, 32)


def test_golden_c4_streams_and_fused_mix():
    g = golden("c4_streams_mix.npz")
    ns, npull, L = int(g["n_streams"]), int(g["n_pulls"]), int(g["L"])
    n = npull * wl.C4_PULL
    irs = np.stack([wl.c4_ir(s, L) for s in range(ns)])
    x = np.stack([wl.c4_input(n, s) for s in range(ns)])[:, None, :]
    bank = pg.ConvolveBank(irs, ns, 1, block=512)
    ys = np.concatenate([bank.process(x[:, :, b * 512:(b + 1) * 512]) for b in range(npull)], axis=2)
    assert rel_err(ys[:, 0, :], g["per_stream"]) <= TOL
    # the drop-in graph: MixPE over ConvolvePEs is adopted into one bank and mixed on the device
    pes = [pg.ConvolvePE(pg.ArrayPE(wl.c4_input(n, s)), pg.ArrayPE(wl.c4_ir(s, L))) for s in range(ns)]
    mix = pg.MixPE(*pes)
    y = _pull_pe(mix, (wl.C4_PULL,) * npull)
    assert mix._fused not in (None, False)
    assert rel_err(y, g["mix"]) <= TOL
    pes = [pg.ConvolvePE(pg.ArrayPE(wl.c4_input(n, s)), pg.ArrayPE(wl.c4_ir(s, L))) for s in range(ns)]
    y2 = _pull_pe(pg.MixPE(*pes, fuse=False), (wl.C4_PULL,) * npull)
    assert rel_err(y2, g["mix"]) <= TOL


def test_golden_c5_voicebank_long_ir_64_blocks():
    g = golden("c5_voicebank_longir.npz")
    nv, npull, L = int(g["n_voices"]), int(g["n_pulls"]), int(g["L"])
    n = npull * wl.C5_PULL
    v = wl.c5_voices(n, nv)
    vm = pg.MixPE(*[pg.ArrayPE(v[i]) for i in range(nv)])
    assert np.array_equal(vm.render(0, n).data, g["voice_mix"])  # K5 is bit-exact
    # block_size=64: the NAMED low-latency partitioning (the default would pick B=256 for a 64-sample first pull)
    pe = pg.ConvolvePE(pg.MixPE(*[pg.ArrayPE(v[i]) for i in range(nv)]), pg.ArrayPE(wl.c5_ir(L)), block_size=64)
    y = _pull_pe(pe, (wl.C5_PULL,) * npull)
    assert pe.bank.block == 64 and pe.bank.partitions == 6891
    assert rel_err(y, g["y"]) <= TOL


def test_golden_mix_extents_bit_exact():
    g = golden("mix_extents.npz")
    a = [g[f"a{i}"] for i in range(7)]
    pes = [pg.ArrayPE(a[0]), pg.DelayPE(pg.ArrayPE(a[1]), 100), pg.ArrayPE(a[2][:50]),
           pg.DelayPE(pg.ArrayPE(a[3]), 250)] + [pg.ArrayPE(t) for t in a[4:]]
    assert np.array_equal(pg.MixPE(*pes).render(0, 600).data, g["y"])
    assert np.array_equal(pg.MixPE(*pes).render(560, 100).data, g["y_late"])


# ---------------------------------------------------------------------------
# seeded random cases against the oracle
@pytest.mark.parametrize("L,B,c_src,c_f", [
    (1, 16, 1, 1), (15, 16, 2, 1), (16, 16, 1, 2), (17, 16, 2, 2), (100, 32, 1, 1), (1000, 128, 3, 3),
    (5000, 512, 2, 1), (4096, 4096, 1, 1), (9000, 2048, 1, 2), (20000, 8192, 1, 1),
])
def test_oracle_parity_shapes(L, B, c_src, c_f):
    rng = np.random.default_rng(1000 + L + B)
    h = (rng.standard_normal((L, c_f)) / np.sqrt(L)).astype(np.float32)
    n = max(3 * B + 37, 2 * L + 11)
    x = rng.uniform(-1, 1, (n, c_src)).astype(np.float32)
    pe = pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(h), block_size=B)
    pulls = []
    left = n + L  # run past the source end to flush the tail
    for d in (1, B - 1, B, B + 1, 7, 2 * B + 3):
        if left <= 0:
            break
        pulls.append(min(d, left))
        left -= pulls[-1]
    if left > 0:
        pulls.append(left)
    y = _pull_pe(pe, pulls)
    ref = _oracle_pulls(orc.OracleConvolve(h, c_src), x, pulls)
    assert y.shape == ref.shape
    assert rel_err(y, ref) <= TOL


def test_bank_many_streams_distinct_and_shared_filters():
    rng = np.random.default_rng(5)
    N, L, B, n = 37, 3000, 256, 1500
    x = rng.uniform(-1, 1, (N, 2, n)).astype(np.float32)
    for shared in (True, False):
        F = 1 if shared else N
        h = (rng.standard_normal((F, L, 2)) / np.sqrt(L)).astype(np.float32)
        bank = pg.ConvolveBank(h, N, 2, block=B)
        pulls = (100, 256, 300, 844)
        ys = np.concatenate([bank.process(np.ascontiguousarray(x[:, :, a:a + d]))
                             for a, d in zip(np.cumsum((0,) + pulls[:-1]), pulls)], axis=2)
        for s in (0, 1, N // 2, N - 1):
            ref = orc.OracleConvolve(h[0 if shared else s], 2).render(x[s].T)
            assert rel_err(ys[s].T, ref) <= TOL
        # fused mix == sum of per-stream outputs
        bank2 = pg.ConvolveBank(h, N, 2, block=B)
        ymix = bank2.process_mix(x)
        ref_mix = np.sum(ys.astype(np.float64), axis=0)
        assert rel_err(ymix, ref_mix) <= TOL


def test_bank_per_stream_reset():
    rng = np.random.default_rng(6)
    N, L, B = 4, 700, 64
    h = (rng.standard_normal((N, L)) / np.sqrt(L)).astype(np.float32)
    x = rng.uniform(-1, 1, (N, 1, 1000)).astype(np.float32)
    bank = pg.ConvolveBank(h, N, 1, block=B)
    bank.process(np.ascontiguousarray(x[:, :, :333]))
    bank.reset(streams=[1, 3])
    y = bank.process(np.ascontiguousarray(x[:, :, 333:]))
    for s in range(N):
        o = orc.OracleConvolve(h[s], 1)
        if s in (1, 3):
            ref = o.render(x[s, 0, 333:])
        else:
            o.render(x[s, 0, :333])
            ref = o.render(x[s, 0, 333:])
        assert rel_err(y[s, 0], ref[:, 0]) <= TOL, s


def test_mix_sum_bit_exact_random():
    rng = np.random.default_rng(9)
    a = [rng.standard_normal((1000, 2)).astype(np.float32) * 10 ** rng.uniform(-3, 3) for _ in range(33)]
    assert np.array_equal(pg.device_mix_sum(a), orc.oracle_mix(a))


# ---------------------------------------------------------------------------
# size-independent properties at BASELINE sizes (the oracle would take minutes there)
def test_full_size_c2_impulse_and_linearity():
    """C2 at the named size (L=132300, B=512, P=259): a unit impulse reproduces the IR exactly
    (to fp32 FFT rounding), and the response is linear in the input."""
    pg.set_sample_rate(wl.SR_48)
    ir = wl.c2_ir()
    L = ir.shape[0]
    bank = pg.ConvolveBank(ir, 3, 2, block=512, single_filter_dims=True, max_pull=8192)
    n = 8192 * 17  # > L
    rng = np.random.default_rng(12)
    a = rng.uniform(-1, 1, (2, 4096)).astype(np.float32)
    x = np.zeros((3, 2, n), np.float32)
    x[0, :, 0] = 1.0                     # impulse
    x[1, :, :4096] = a                   # signal
    x[2, :, :4096] = 0.5 * a
    x[2, :, 0] += 0.25                   # 0.5*signal + 0.25*impulse
    y = bank.process(x)
    assert rel_err(y[0, :, :L].T, ir) <= TOL
    assert np.max(np.abs(y[0, :, L:])) <= TOL * np.max(np.abs(ir))
    lin = 0.5 * y[1].astype(np.float64) + 0.25 * y[0].astype(np.float64)
    assert rel_err(y[2], lin) <= TOL


def _fftconv64(x, h, n):
    """Exact linear convolution in float64 (the definition both the reference and the device path implement,
    SURVEY.md Appendix A), first n samples."""
    from scipy.signal import fftconvolve
    return fftconvolve(np.asarray(x, np.float64), np.asarray(h, np.float64))[:n]


def test_full_plan_c2_256_streams_shared_ir_every_partition():
    """The BENCHMARKED launch plan (256 stereo streams, one shared 132300-tap IR, B=512, P=259: k_fdl_mac<false,4,4>,
    5 term splits) with noise over n > L samples, so every one of the 259 partitions of every stream carries data
    when the last pulls are produced; streams {0, 1, N/2, N-1} against the float64 convolution."""
    pg.set_sample_rate(wl.SR_48)
    ir = wl.c2_ir()
    L, N, pull = ir.shape[0], 256, 512
    bank = pg.ConvolveBank(ir, N, 2, block=512, single_filter_dims=True, max_pull=pull)
    info = bank.info()
    assert info.partitions == 259 and info.mac_stream_tile == (4 if info.mac_tile == 1 else 2)   # the plan bench.py times
    pulls = 259 + 3
    n = pulls * pull
    rng = np.random.default_rng(77)
    sample = (0, 1, N // 2, N - 1)
    ys = {s: [] for s in sample}
    xs = {s: [] for s in sample}
    for i in range(pulls):
        x = rng.uniform(-1, 1, (N, 2, pull)).astype(np.float32)
        y = bank.process(x)
        for s in sample:
            xs[s].append(x[s].copy())
            ys[s].append(y[s].copy())
    for s in sample:
        x = np.concatenate(xs[s], axis=1)
        y = np.concatenate(ys[s], axis=1)
        for c in range(2):
            ref = _fftconv64(x[c], ir[:, c], n)
            assert rel_err(y[c], ref) <= TOL, (s, c)
            assert rel_err(y[c, -pull:], ref[-pull:]) <= TOL * max(1.0, np.max(np.abs(ref)) / np.max(np.abs(ref[-pull:])))
    bank.close()


def test_full_plan_c2_distinct_irs_64_streams():
    """Distinct-filter plan (k_fdl_mac<false,1,8>) at the named filter length, n > L."""
    pg.set_sample_rate(wl.SR_48)
    N, pull = 64, 2048
    irs = np.stack([wl.c2_ir(stream=s) for s in range(N)])
    L = irs.shape[1]
    bank = pg.ConvolveBank(irs, N, 2, block=512, max_pull=pull)
    assert bank.info().mac_stream_tile == 1
    pulls = -(-(L + 1024) // pull)
    n = pulls * pull
    rng = np.random.default_rng(78)
    sample = (0, 31, N - 1)
    x = rng.uniform(-1, 1, (N, 2, n)).astype(np.float32)
    y = np.concatenate([bank.process(np.ascontiguousarray(x[:, :, i * pull:(i + 1) * pull])) for i in range(pulls)], axis=2)
    for s in sample:
        for c in range(2):
            assert rel_err(y[s, c], _fftconv64(x[s, c], irs[s, :, c], n)) <= TOL, (s, c)
    bank.close()


def test_full_plan_c4_512_streams_all_173_partitions_fused_mix():
    """C4 per-GPU size (512 mono streams x distinct 88200-tap IRs, fused mix, the per-stream mix layout) over
    n > L samples.  Every stream gets a scaled impulse at its own offset -- its spectrum marches through all 173
    partitions of THAT stream's filter, and its exact contribution is the shifted IR -- and 8 streams get noise on
    top (float64 convolution).  The fused mix must equal the sum."""
    N, pull = 512, 512
    irs = np.stack([wl.c4_ir(s) for s in range(N)])
    L = irs.shape[1]
    bank = pg.ConvolveBank(irs, N, 1, block=512, max_pull=pull)
    assert bank.info().partitions == 173
    pulls = 173 + 3
    n = pulls * pull
    rng = np.random.default_rng(79)
    amp = rng.uniform(0.5, 1.0, N) * rng.choice([-1.0, 1.0], N)
    off = rng.integers(0, 2048, N)
    noisy = rng.choice(N, 8, replace=False)
    x = np.zeros((N, 1, n), np.float32)
    x[np.arange(N), 0, off] = amp.astype(np.float32)
    ref = np.zeros(n)
    for s in range(N):
        a = float(np.float32(amp[s]))
        m = min(L, n - off[s])
        ref[off[s]:off[s] + m] += a * irs[s, :m].astype(np.float64)
    for s in noisy:
        nz = (rng.uniform(-1, 1, n) / 8).astype(np.float32)
        nz[off[s]] = 0.0
        x[s, 0] += nz
        ref += _fftconv64(nz, irs[s], n)
    y = np.concatenate([bank.process_mix(np.ascontiguousarray(x[:, :, i * pull:(i + 1) * pull])) for i in range(pulls)], axis=1)
    assert rel_err(y[0], ref) <= TOL
    tail = slice(n - 4 * pull, n)                          # produced when all 173 partitions are live
    assert np.max(np.abs(y[0, tail] - ref[tail])) <= TOL * np.max(np.abs(ref))
    bank.close()


def test_full_plan_c5_single_stream_all_6891_partitions():
    """C5 as named (one stream, 441000 taps, B=64: 6891 partitions), noise over n > L samples pulled 4096 at a
    time (64 block steps per pull), against the float64 convolution."""
    ir = wl.c5_ir()
    L = ir.shape[0]
    bank = pg.ConvolveBank(ir, 1, 1, block=64, single_filter_dims=True, max_pull=4096)
    assert bank.info().partitions == 6891
    pulls = -(-(L + 4096) // 4096)
    n = pulls * 4096
    x = (np.random.default_rng(80).uniform(-1, 1, n) / 4).astype(np.float32)
    y = np.concatenate([bank.process(x[None, None, i * 4096:(i + 1) * 4096])[0, 0] for i in range(pulls)])
    ref = _fftconv64(x, ir, n)
    assert rel_err(y, ref) <= TOL
    assert np.max(np.abs(y[-4096:] - ref[-4096:])) <= TOL * np.max(np.abs(ref))
    assert bank.info().block_steps >= 6891
    bank.close()


@pytest.mark.parametrize("P,B", [(2, 64), (7, 128), (16, 64)])
def test_conv_to_mix_switch_back_to_back_on_a_block_boundary(P, B):
    """Banks with 2..16 partitions run conv pulls as ONE fused kernel on the critical stream; a mix pull queued
    right behind (no host synchronisation: pipelined submits, resident input) must see what that kernel wrote."""
    rng = np.random.default_rng(P)
    N, L = 5, P * B - 3
    h = (rng.standard_normal((N, L)) / np.sqrt(L)).astype(np.float32)
    n_conv, n_mix = 3 * B, 4 * B
    x = rng.uniform(-1, 1, (N, 1, n_conv + n_mix)).astype(np.float32)
    for trial in range(8):                       # a race needs several tries to show
        bank = pg.ConvolveBank(h, N, 1, block=B, max_pull=B)
        y1 = [np.empty((N, 1, B), np.float32) for _ in range(3)]
        y2 = [np.empty((1, B), np.float32) for _ in range(4)]
        tks = [bank.submit(np.ascontiguousarray(x[:, :, i * B:(i + 1) * B]), y1[i]) for i in range(3)]
        tks += [bank.submit(np.ascontiguousarray(x[:, :, n_conv + i * B:n_conv + (i + 1) * B]), y2[i], mix=True)
                for i in range(4)]
        for t in tks:
            bank.wait(t)
        ref = np.stack([orc.OracleConvolve(h[s], 1).render(x[s, 0])[:, 0] for s in range(N)])
        assert rel_err(np.concatenate(y1, axis=2)[:, 0], ref[:, :n_conv]) <= TOL
        assert rel_err(np.concatenate(y2, axis=1)[0], ref[:, n_conv:].astype(np.float64).sum(axis=0)) <= TOL, trial
        bank.close()


def test_adopted_convolve_mix_gates_on_extents_like_the_reference():
    """MixPE over ConvolvePEs adopts them into one bank; inputs whose extent misses the request are not rendered
    (mix_pe.py:81-85) -- a HOLD_LAST source therefore stops contributing where its ConvolvePE's extent ends."""
    rng = np.random.default_rng(3)
    a = rng.uniform(-1, 1, 700).astype(np.float32)
    b = rng.uniform(-1, 1, 300).astype(np.float32)
    h = (rng.standard_normal(40) / 6).astype(np.float32)

    def graph(fuse):
        return pg.MixPE(pg.ConvolvePE(pg.ArrayPE(a), pg.ArrayPE(h)),
                        pg.ConvolvePE(pg.ArrayPE(b, extend_mode=pg.ExtendMode.HOLD_LAST), pg.ArrayPE(h)), fuse=fuse)
    pulls = (256,) * 4
    y_f, y_u = _pull_pe(graph(True), pulls), _pull_pe(graph(False), pulls)
    assert rel_err(y_f, y_u) <= TOL
    assert np.max(np.abs(y_u[512:, 0])) > 0          # stream a still sounds ...
    # ... and past its extent (300 + 39) stream b is silent although its source holds its last value forever
    ca = orc.OracleConvolve(h, 1).render(np.concatenate([a, np.zeros(324, np.float32)]))[:, 0]
    assert rel_err(y_f[512:, 0], ca[512:]) <= TOL


@pytest.mark.parametrize("variant", ["0", "1", "2", "3", "4"])
def test_c1_block4096_fft_kernel_variants_agree_with_oracle(variant, monkeypatch):
    """B = 4096 single-partition step: the radix-8 kernel (0), the radix-16 kernel with shared-memory exchanges (1)
    and with the half-warp exchange by warp shuffles (2), and both with split / product / merge on mirror pairs in
    registers (3, 4) against the oracle -- whole blocks (the
    radix-16 fast path), then ragged pulls (general kernel), with a wet/dry output stage, mono and stereo."""
    monkeypatch.setenv("PGX_FFT16", variant)
    rng = np.random.default_rng(int(variant) + 40)
    for c in (1, 2):
        N, L, B = 5, 4096, 4096
        h = (rng.standard_normal((N, L, 1)) / 64).astype(np.float32)
        x = rng.uniform(-1, 1, (N, c, 6 * B + 777)).astype(np.float32)
        bank = pg.ConvolveBank(h, N, c, block=B, max_pull=B)
        if c == 2:
            bank.set_output_gains(0.25, 0.5)
        pulls = (B,) * 4 + (100, B - 100, 777, B)
        ys, pos = [], 0
        for d in pulls:
            ys.append(bank.process(np.ascontiguousarray(x[:, :, pos:pos + d])))
            pos += d
        y = np.concatenate(ys, axis=2)
        for s_ in (0, N - 1):
            ref = orc.OracleConvolve(h[s_], c).render(x[s_].T).astype(np.float64)
            if c == 2:
                ref = 0.5 * x[s_].T + 0.25 * ref
            assert rel_err(y[s_].T, ref) <= TOL, (variant, c, s_)
        bank.close()


@pytest.mark.parametrize("L,B,c", [(5000, 64, 1), (700, 64, 2), (100, 128, 1), (132300, 512, 2)])
def test_graph_replay_of_whole_block_pulls_is_bit_identical_to_the_streamed_schedule(L, B, c, monkeypatch):
    """Small whole-block host pulls run as ONE CUDA-graph replay (H2D, kernels, D2H; kernel-node arguments refreshed
    per step).  Same kernels, same arguments: the output must equal the multi-stream schedule's BIT FOR BIT -- through
    whole blocks, ragged pulls in between (which leave and re-enter graph mode), resets and a gain change -- and both
    must match the oracle."""
    rng = np.random.default_rng(L)
    h = (rng.standard_normal((L, c)) / np.sqrt(L)).astype(np.float32)
    pulls = (B,) * 9 + (B // 2, B // 2 - 3, 3) + (B,) * 6 + (2 * B,) + (B,) * 4
    n = sum(pulls)
    x = rng.uniform(-1, 1, (n, c)).astype(np.float32)
    outs = {}
    for mode in ("0", "1", "1f"):          # streamed schedule, graph replay, graph replay with K1+K2 as one launch
        monkeypatch.setenv("PGX_GRAPH", mode[0])
        monkeypatch.setenv("PGX_GRAPH_FUSE", "1" if mode == "1f" else "0")
        bank = pg.ConvolveBank(h, 1, c, block=B, single_filter_dims=True, max_pull=2 * B)
        ys, pos = [], 0
        for rnd in range(2):
            for i, d in enumerate(pulls):
                if rnd == 1 and i == 5:
                    bank.set_output_gains(0.5, 0.25)
                ys.append(bank.process_interleaved(x[pos % n:pos % n + d]))
                pos += d
            if rnd == 0:
                bank.reset()
        outs[mode] = np.concatenate(ys)
        info = bank.info()
        import os
        if (os.environ.get("PGX_MAC") == "tma" and info.partitions > 16) or info.mac_tile > 1:
            assert mode != "0" or info.graph_pulls == 0     # (the bulk-async and the time-tiled pass have no graph form: either way is fine)
        else:
            assert (info.graph_pulls > 0) == (mode != "0"), (mode, info.graph_pulls)
        bank.close()
    assert np.array_equal(outs["0"], outs["1f"])
    assert np.array_equal(outs["0"], outs["1"])
    ref = orc.OracleConvolve(h, c).render(x)
    assert rel_err(outs["1"][:n], ref) <= TOL


def test_graph_replay_fused_hrtf_mix_matches_streamed(monkeypatch):
    """The one-launch HRTF mix step (k_mix1<LAST>) with resident sources as a graph replay, directions re-selected
    between pulls: bit-identical to the streamed schedule."""
    rng = np.random.default_rng(11)
    srcs = [rng.uniform(-1, 1, 512 * 12).astype(np.float32) / 8 for _ in range(24)]
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("PGX_GRAPH", mode)
        methods = [pg.SpatialHRTF(-170.0 + 14 * i, 10.0 * (i % 5)) for i in range(24)]
        mix = pg.MixPE(*[pg.SpatialPE(pg.ArrayPE(s_), method=m) for s_, m in zip(srcs, methods)])
        ys = []
        for p in range(12):
            if p % 3 == 2:
                for i, m in enumerate(methods):
                    m.azimuth = ((m.azimuth + 25.0 + i + 180.0) % 360.0) - 180.0
            ys.append(mix.render(p * 512, 512).data.copy())
        outs[mode] = np.concatenate(ys)
        gp = mix._fused.bank.info().graph_pulls
        assert (gp > 0) == (mode == "1"), (mode, gp)
    assert np.array_equal(outs["0"], outs["1"])


def test_trajectory_table_equals_per_pull_attribute_assignment():
    """MixPE.set_trajectory (directions of every pull resolved once, one device-resident int32 table) gives exactly
    what assigning method.azimuth / .elevation before every pull gives -- including sources that start late or end
    early (extent gating) and a non-contiguous pull."""
    rng = np.random.default_rng(21)
    N, pull, n_p = 20, 256, 14
    lens = rng.integers(pull * 6, pull * n_p, N)
    data = [rng.uniform(-1, 1, int(l)).astype(np.float32) / 4 for l in lens]
    delays = [0] * (N - 3) + [300, 700, 1500]
    az = rng.uniform(-180, 180, (n_p, N))
    el = rng.uniform(-40, 90, (n_p, N))

    def graph():
        ms = [pg.SpatialHRTF(az[0, i], el[0, i]) for i in range(N)]
        return pg.MixPE(*[pg.DelayPE(pg.SpatialPE(pg.ArrayPE(d), method=m), delay=dl) if dl else
                          pg.SpatialPE(pg.ArrayPE(d), method=m) for d, m, dl in zip(data, ms, delays)]), ms
    order = list(range(n_p)) + [3, 4, 5]                     # ... then jump back: a non-contiguous pull
    mix_a, ms = graph()
    ya = []
    for p in order:
        for i, m in enumerate(ms):
            m.azimuth, m.elevation = az[p, i], el[p, i]
        ya.append(mix_a.render(p * pull, pull).data.copy())
    mix_b, _ = graph()
    mix_b.set_trajectory(az, el, hop=pull)
    yb = [mix_b.render(p * pull, pull).data.copy() for p in order]
    assert np.array_equal(np.concatenate(ya), np.concatenate(yb))
    assert np.max(np.abs(np.concatenate(ya))) > 0
    mix_b.set_trajectory(None, hop=pull)                     # back to the attributes (still those of construction)
    yc = mix_b.render(0, pull).data
    mix_c, _ = graph()
    assert np.array_equal(yc, mix_c.render(0, pull).data)


def test_mix_with_one_unbankable_input_still_fuses_the_rest():
    """A MixPE whose inputs are ConvolvePEs except one wrapped in a PE-VALUED GainPE (its per-sample gain applies after
    the convolution, gain_pe.py:92-127, so it cannot ride in the frequency-domain sum) and one plain ArrayPE: the
    bank-able inputs are adopted into one bank, the others are rendered as they are and added -- equal to the
    unfused graph, and the bank really exists."""
    rng = np.random.default_rng(8)
    n = 2048
    xs = [rng.uniform(-1, 1, n).astype(np.float32) / 3 for _ in range(5)]
    hs = [(rng.standard_normal(200) / 14).astype(np.float32) for _ in range(5)]
    gain_ctl = (0.5 + 0.5 * np.sin(np.arange(n) / 50.0)).astype(np.float32)

    def graph(fuse):
        convs = [pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(h)) for x, h in zip(xs, hs)]
        ins = convs[:3] + [pg.GainPE(convs[3], pg.ArrayPE(gain_ctl)), pg.DelayPE(pg.GainPE(convs[4], 0.5), delay=100),
                           pg.ArrayPE(xs[0])]
        return pg.MixPE(*ins, fuse=fuse)
    mf, mu = graph(True), graph(False)
    pulls = (512, 100, 412, 1024)
    yf, yu = _pull_pe(mf, pulls), _pull_pe(mu, pulls)
    assert rel_err(yf, yu) <= TOL
    assert isinstance(mf.fused_bank, pg.ConvolveBank) and mf.fused_bank.n_streams == 4 and len(mf._rest) == 2
    assert mu.fused_bank is None


def test_full_size_c5_impulse_response():
    ir = wl.c5_ir()
    L = ir.shape[0]
    pe = pg.ConvolvePE(pg.ArrayPE(np.array([1.0], np.float32)), pg.ArrayPE(ir), block_size=64)
    y = _pull_pe(pe, (64,) * 40)[:, 0]   # 40 low-latency pulls through P = 6891 partitions
    assert rel_err(y, ir[:y.shape[0]]) <= TOL
    del L


# ---------------------------------------------------------------------------
# callers of the path: ReverbPE, AudioRenderer pull loop, BankRenderer
class _CountingPE(pg.ProcessingElement):
    def __init__(self, source):
        self._source, self.render_calls = source, 0

    def inputs(self):
        return [self._source]

    def channel_count(self):
        return self._source.channel_count()

    def _compute_extent(self):
        return self._source.extent()

    def _render(self, start, duration):
        self.render_calls += 1
        return self._source.render(start, duration)


def test_reverb_uses_single_source_pull():
    # reference tests/test_convolve_pe.py:210-221
    pg.set_sample_rate(10_000)
    src = _CountingPE(pg.ArrayPE([1.0, 2.0, 3.0, 4.0]))
    pe = pg.ReverbPE(src, pg.ArrayPE([1.0]), mix=0.5, normalize_ir=False)
    pg.NullRenderer(sample_rate=10_000).set_source(pe)
    _ = pe.render(0, 4)
    assert src.render_calls == 1


def test_reverb_wet_dry_matches_oracle():
    rng = np.random.default_rng(31)
    x = rng.uniform(-1, 1, (3000, 2)).astype(np.float32)
    ir = (rng.standard_normal(2500) * np.exp(-np.arange(2500) / 400.0)).astype(np.float32)
    pe = pg.ReverbPE(pg.ArrayPE(x), pg.ArrayPE(ir), mix=0.3)
    y = _pull_pe(pe, (512,) * 5 + (440,))
    wet = orc.OracleConvolve(ir, 2).render(x)
    e = orc.ir_energy_norm(ir)
    ref = orc.oracle_mix([x * np.float32(0.7), wet * np.float32(0.3 / e)])
    assert pe.ir_energy == pytest.approx(e)
    assert rel_err(y, ref) <= TOL


class _Sink:
    def __init__(self, sr, ch, bs):
        self.meta, self.blocks, self.state = (sr, ch, bs), [], []

    def start(self):
        self.state.append("start")

    def write(self, data):
        assert data.dtype == np.float32 and data.ndim == 2
        self.blocks.append(data.copy())

    def stop(self):
        self.state.append("stop")

    def close(self):
        self.state.append("close")


def test_audio_renderer_play_extent_pulls_in_chunks():
    rng = np.random.default_rng(32)
    x = rng.uniform(-1, 1, 5000).astype(np.float32)
    h = (rng.standard_normal(300) / 17).astype(np.float32)
    sinks = []

    def factory(sr, ch, bs):
        sinks.append(_Sink(sr, ch, bs))
        return sinks[-1]

    r = pg.AudioRenderer(sample_rate=44_100, blocksize=64, stream_factory=factory)
    r.set_source(pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(h)))
    r.start()
    r.play_extent()                       # chunks of blocksize*16 = 1024 over extent [0, 5299)
    r.stop()
    got = np.concatenate(sinks[0].blocks)[:, 0]
    assert [b.shape[0] for b in sinks[0].blocks] == [1024] * 5 + [179]
    assert sinks[0].state == ["start", "stop", "close"] and sinks[0].meta == (44_100, 1, 64)
    ref = orc.OracleConvolve(h, 1).render(np.concatenate([x, np.zeros(299, np.float32)]))[:, 0]
    assert rel_err(got, ref) <= TOL


def test_bank_renderer_lockstep_pull():
    rng = np.random.default_rng(33)
    N, L = 5, 1200
    xs = [rng.uniform(-1, 1, 2048).astype(np.float32) for _ in range(N)]
    hs = (rng.standard_normal((N, L)) / 30).astype(np.float32)
    bank = pg.ConvolveBank(hs, N, 1, pull_hint=256)
    bank.attach_sources([pg.ArrayPE(x) for x in xs])
    outs = []
    with pg.BankRenderer(bank, sink=outs.append) as r:
        r.start()
        for p in range(0, 2048, 256):
            r.render(p, 256)
    y = np.concatenate(outs, axis=2)
    for s in range(N):
        ref = orc.OracleConvolve(hs[s], 1).render(xs[s])[:, 0]
        assert rel_err(y[s, 0], ref) <= TOL


# ---------------------------------------------------------------------------
# device-resident queue: pulls enqueued back to back with no host synchronisation, ingest of pull i+1
# allowed to overtake the output stage of pull i (PGX_PULL_INPUT_RESIDENT) -- exercises every
# cross-stream hazard of the three-stream schedule
@pytest.mark.parametrize("mix", [False, True])
@pytest.mark.parametrize("L,B", [(3000, 256), (700, 64), (100, 128)])
def test_device_queue_back_to_back(mix, L, B):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(77 + L)
    N, C = 3, 2
    h = (rng.standard_normal((N, L, C)) / np.sqrt(L)).astype(np.float32)
    pulls = [B] * 9 + [1, B - 1, 17, B, 2 * B + 5, B // 2, B // 2, B] + [B] * 9
    total = sum(pulls)
    x = rng.uniform(-1, 1, (N, C, total)).astype(np.float32)
    bank = pg.ConvolveBank(h, N, C, block=B, max_pull=4 * B)
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(device=dev)
    xs, ys, pos = [], [], 0
    for d in pulls:                      # every pull has its own resident input / output buffer
        xs.append(torch.from_numpy(np.ascontiguousarray(x[:, :, pos:pos + d])).to(dev))
        ys.append(torch.empty((C, d) if mix else (N, C, d), dtype=torch.float32, device=dev))
        pos += d
    torch.cuda.synchronize(dev)
    for rep in range(3):                 # repeat the whole queue: different interleavings, same answer
        bank.reset()
        for xt, yt, d in zip(xs, ys, pulls):
            bank.process_device(xt.data_ptr(), yt.data_ptr(), d, mix=mix, cuda_stream=st.cuda_stream,
                                input_resident=True)
        st.synchronize()
        y = np.concatenate([t.cpu().numpy() for t in ys], axis=-1)
        per = np.stack([orc.OracleConvolve(h[s], C).render(x[s].T).T for s in range(N)])
        ref = per.astype(np.float64).sum(axis=0) if mix else per
        assert rel_err(y, ref) <= TOL, f"rep {rep}"
    bank.synchronize()


# ---------------------------------------------------------------------------
# pipelined host-buffer pulls: pgx_bank_submit / pgx_bank_wait with up to three pulls in flight must give
# exactly what the synchronous pgx_bank_process gives, in submission order, ragged pulls included
@pytest.mark.parametrize("mix", [False, True])
def test_submit_wait_pipeline_matches_synchronous(mix):
    from pygmu2_b200._lib import PinnedArray
    rng = np.random.default_rng(91)
    N, C, L, B = 4, 2, 2000, 128
    h = (rng.standard_normal((N, L, C)) / np.sqrt(L)).astype(np.float32)
    pulls = [B] * 6 + [5, B - 5, 3 * B, 77, B] + [B] * 6
    x = rng.uniform(-1, 1, (N, C, sum(pulls))).astype(np.float32)
    bank = pg.ConvolveBank(h, N, C, block=B, max_pull=4 * B)
    sync = []
    pos = 0
    for d in pulls:
        xc = np.ascontiguousarray(x[:, :, pos:pos + d])
        sync.append(bank.process_mix(xc) if mix else bank.process(xc))
        pos += d
    bank.reset()
    xin = [PinnedArray((N, C, d)) for d in pulls]
    yout = [PinnedArray((C, d) if mix else (N, C, d)) for d in pulls]
    pos = 0
    for a, d in zip(xin, pulls):
        a.array[...] = x[:, :, pos:pos + d]
        pos += d
    tickets = []
    for i, d in enumerate(pulls):
        tickets.append(bank.submit(xin[i].array, yout[i].array, mix=mix))
        if i >= 2:
            bank.wait(tickets[i - 2])
    for t in tickets[-2:]:
        bank.wait(t)
    for i in range(len(pulls)):
        np.testing.assert_array_equal(yout[i].array, sync[i])
    per = np.stack([orc.OracleConvolve(h[s], C).render(x[s].T).T for s in range(N)])
    ref = per.astype(np.float64).sum(axis=0) if mix else per
    assert rel_err(np.concatenate([a.array for a in yout], axis=-1), ref) <= TOL
    with pytest.raises(ValueError):
        bank.wait(10_000)
    for a in xin + yout:
        a.free()
    bank.close()


def test_submit_pipeline_with_filter_map_reselected_every_pull():
    """Moving sources: the filter map changes before every pull while earlier pulls are still in flight
    (the map ring must neither race nor drain).  Reference semantics: the filter of the current pull is
    applied to the whole carried history (spatial_pe.py:446-449,499-504)."""
    from pygmu2_b200._lib import PinnedArray
    rng = np.random.default_rng(92)
    N, F, L, B, n_pulls = 6, 5, 128, 128, 40
    h = (rng.standard_normal((F, L, 2)) / np.sqrt(L)).astype(np.float32)
    x = rng.uniform(-1, 1, (N, 1, n_pulls * B)).astype(np.float32)
    maps = rng.integers(0, F, (n_pulls, N)).astype(np.int32)
    bank = pg.ConvolveBank(h, N, 1, block=B, max_pull=B, filter_of_stream=maps[0])
    xin = [PinnedArray((N, 1, B)) for _ in range(n_pulls)]
    yout = [PinnedArray((2, B)) for _ in range(n_pulls)]
    for i, a in enumerate(xin):
        a.array[...] = x[:, :, i * B:(i + 1) * B]
    tickets = []
    for i in range(n_pulls):
        bank.set_filter_map(maps[i])
        tickets.append(bank.submit(xin[i].array, yout[i].array, mix=True))
    for t in tickets:
        bank.wait(t)
    # oracle: per pull, each stream's current filter convolved with its last 2B samples of history
    for i in range(n_pulls):
        ref = np.zeros((2, B))
        for s in range(N):
            seg = x[s, 0, max(0, (i - 1) * B):(i + 1) * B].astype(np.float64)
            for c in range(2):
                full = np.convolve(seg, h[maps[i, s], :, c].astype(np.float64))
                ref[c] += full[seg.shape[0] - B:seg.shape[0]]
        assert rel_err(yout[i].array, ref) <= TOL, f"pull {i}"
    for a in xin + yout:
        a.free()
    bank.close()


# ---------------------------------------------------------------------------
# ReverbPE: fused wet/dry epilogue (pgx_bank_set_output_gains) vs the reference's composite graph
def test_reverb_fused_is_bit_identical_to_composite_graph():
    rng = np.random.default_rng(41)
    x = rng.uniform(-1, 1, (4000, 2)).astype(np.float32)
    ir = (rng.standard_normal((1800, 2)) * np.exp(-np.arange(1800)[:, None] / 300.0)).astype(np.float32)
    pulls = (512, 17, 512, 1000, 959, 1000)
    fused = pg.ReverbPE(pg.ArrayPE(x), pg.ArrayPE(ir), mix=0.35)
    assert fused._fused
    e = fused.ir_energy
    src = pg.CachePE(pg.ArrayPE(x))
    comp = pg.MixPE(pg.GainPE(src, gain=1.0 - 0.35),
                    pg.GainPE(pg.ConvolvePE(src, pg.ArrayPE(ir)), gain=0.35 / e), fuse=False)
    np.testing.assert_array_equal(_pull_pe(fused, pulls), _pull_pe(comp, pulls))


def test_reverb_pe_valued_mix_matches_oracle():
    rng = np.random.default_rng(42)
    n = 3000
    x = rng.uniform(-1, 1, (n, 1)).astype(np.float32)
    ir = (rng.standard_normal(700) / 20).astype(np.float32)
    mixv = np.linspace(0.0, 1.0, n + 699, dtype=np.float32)
    pe = pg.ReverbPE(pg.ArrayPE(x), pg.ArrayPE(ir), mix=pg.ArrayPE(mixv))
    assert not pe._fused
    y = _pull_pe(pe, (512,) * 5 + (440,))
    wet = orc.OracleConvolve(ir, 1).render(x)
    e = np.float32(1.0 / orc.ir_energy_norm(ir))
    m = mixv[:n, None]
    dry_g = np.float32(1.0) + m * np.float32(-1.0)
    ref = x * dry_g + wet * (m * e)
    assert rel_err(y, ref) <= TOL


def test_output_gains_contract():
    bank = pg.ConvolveBank(np.ones((1, 8), np.float32), 2, 1, block=16)
    bank.set_output_gains(0.5, 0.25)
    x = np.random.default_rng(5).uniform(-1, 1, (2, 1, 40)).astype(np.float32)
    y = bank.process(x)
    wet = np.stack([np.convolve(x[s, 0].astype(np.float64), np.ones(8))[:40] for s in range(2)])[:, None, :]
    assert rel_err(y, 0.25 * x + 0.5 * wet) <= TOL
    fan = pg.ConvolveBank(np.ones((1, 8, 2), np.float32), 1, 1, block=16)   # mono source, stereo filter
    with pytest.raises(ValueError):
        fan.set_output_gains(1.0, 0.5)
    fan.set_output_gains(0.5, 0.0)


# ---------------------------------------------------------------------------
# AudioRenderer.stream_start (audio_renderer.py:183-248): the sink's own thread pulls the device PE
class _CallbackSink:
    """Stand-in for sounddevice.OutputStream(callback=...): a thread that asks for `frames` samples at a time."""

    def __init__(self, sr, ch, bs, callback=None):
        import threading
        self.ch, self.bs, self.cb = ch, bs, callback
        self.blocks, self.done = [], threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.done.is_set():
            out = np.full((self.bs, self.ch), np.nan, dtype=np.float32)
            try:
                self.cb(out, self.bs, None, None)
            except pg.CallbackStop:
                self.done.set()
                break
            self.blocks.append(out)

    def start(self):
        self.thread.start()

    def stop(self):
        self.done.set()
        self.thread.join(timeout=10)

    def close(self):
        pass


def test_audio_renderer_stream_start_pulls_from_the_callback_thread():
    rng = np.random.default_rng(51)
    x = rng.uniform(-1, 1, 3000).astype(np.float32)
    h = (rng.standard_normal(400) / 10).astype(np.float32)
    pe = pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(h))
    sinks = []

    def factory(sr, ch, bs, callback=None):
        sinks.append(_CallbackSink(sr, ch, bs, callback))
        return sinks[-1]

    r = pg.AudioRenderer(sample_rate=44_100, blocksize=256, stream_factory=factory)
    r.set_source(pe)
    r.start()
    r.stream_start(0, 3399)                 # extent end: 3000 + 400 - 1
    assert sinks[0].done.wait(timeout=30)
    r.stream_stop()
    r.stop()
    got = np.concatenate(sinks[0].blocks)[:, 0]
    assert got.shape[0] == 14 * 256          # 13 full blocks + a zero-padded one of 71 samples
    ref = orc.OracleConvolve(h, 1).render(np.concatenate([x, np.zeros(399, np.float32)]))[:, 0]
    assert rel_err(got[:3399], ref) <= TOL and not np.any(got[3399:])


# ---------------------------------------------------------------------------
# DelayPE / GainPE / pan laws folded into the fused mix, and MixPE's extent gating of
# SpatialPE inputs -- against the REAL reference's output for the same graph (oracle/gen_golden_mixfold.py)
def _mixfold_graph(fuse=True):
    lengths, chans = [3000, 4000, 2000, 5000, 700, 1500], [1, 2, 1, 1, 1, 1]
    s = []
    for i, (n, c) in enumerate(zip(lengths, chans)):
        x = np.random.default_rng(700 + i).uniform(-1, 1, (n, c)).astype(np.float32) / 4
        s.append(pg.ArrayPE(x if c > 1 else x[:, 0]))
    return pg.MixPE(
        pg.DelayPE(pg.SpatialPE(s[0], method=pg.SpatialHRTF(30.0, 0.0)), 300),
        pg.GainPE(pg.SpatialPE(s[1], method=pg.SpatialHRTF(-100.0, 20.0)), 0.5),
        pg.DelayPE(pg.GainPE(pg.SpatialPE(s[2], method=pg.SpatialLinear(-45.0)), 0.8), 1000),
        pg.SpatialPE(s[3], method=pg.SpatialConstantPower(60.0)),
        pg.SpatialPE(s[4], method=pg.SpatialHRTF(170.0, -10.0)),
        pg.DelayPE(pg.SpatialPE(s[5], method=pg.SpatialHRTF(0.0, 90.0)), 2500),
        fuse=fuse)


@pytest.mark.parametrize("fuse", [True, False])
def test_golden_mix_fold_delay_gain_pan_and_extent_gating(fuse):
    g = golden("mix_fold.npz")
    mix = _mixfold_graph(fuse)
    ext = mix.extent()
    assert [ext.start, ext.end] == list(g["extent"])
    y = _pull_pe(mix, g["pulls"])
    assert y.shape == g["y"].shape
    assert rel_err(y, g["y"]) <= TOL
    if fuse:
        from pygmu2_b200.hrtf_bank import HrtfMixBank
        assert isinstance(mix._fused, HrtfMixBank) and len(mix._fused.pan_index) == 2
        # source 4 ends at 700: the reference's MixPE stops rendering it, cutting its HRTF tail (pull 2 onwards)
        assert not mix._fused._was_active[4]


@pytest.mark.parametrize("fuse", [True, False])
def test_golden_mix_with_pe_valued_gain_and_pan_azimuth(fuse):
    """A MixPE over HRTF sources where one input is scaled by a PE-VALUED GainPE (gain_pe.py:105-121) and one is panned
    by a PE-valued azimuth (spatial_pe.py:179-214), against outputs of the REAL reference
    (oracle/gen_golden_mixfold.py::graph_pe_controls).  Fused: the four bank-able inputs share one HrtfMixBank, the two
    per-sample-controlled ones are rendered beside it and added."""
    import os
    from conftest import ROOT
    g = golden("mix_fold_pe_controls.npz")
    # the generator's graph builder is module-agnostic (it takes the package as an argument): reuse its text here
    # without importing the reference
    src = open(os.path.join(ROOT, "oracle", "gen_golden_mixfold.py")).read()
    ns = {"np": np, "SR": 44_100}
    body = src[src.index("LENGTHS ="):src.index("def main():")]
    exec(body, ns)
    mix = ns["graph_pe_controls"](pg, fuse=fuse)
    ext = mix.extent()
    assert [ext.start, ext.end] == list(g["extent"])
    y = _pull_pe(mix, g["pulls"])
    assert y.shape == g["y"].shape and rel_err(y, g["y"]) <= TOL
    if fuse:
        from pygmu2_b200.hrtf_bank import HrtfMixBank
        assert isinstance(mix.fused_bank, HrtfMixBank) and len(mix.fused_bank.sources) == 4 and len(mix._rest) == 2


def test_mix_of_delayed_scaled_convolves_folds_into_one_bank():
    rng = np.random.default_rng(61)
    N, L = 4, 900
    xs = [rng.uniform(-1, 1, 2000 + 100 * i).astype(np.float32) for i in range(N)]
    hs = (rng.standard_normal((N, L)) / 30).astype(np.float32)
    delays, gains = [0, 128, 777, 1500], [None, 0.5, 2.0, 0.25]

    def chain(i):
        pe = pg.ConvolvePE(pg.ArrayPE(xs[i]), pg.ArrayPE(hs[i]))
        if gains[i] is not None:
            pe = pg.GainPE(pe, gains[i])
        return pg.DelayPE(pe, delays[i]) if delays[i] else pe

    mix = pg.MixPE(*[chain(i) for i in range(N)])
    y = _pull_pe(mix, [256] * 20)[:, 0]
    assert isinstance(mix._fused, pg.ConvolveBank)
    ref = np.zeros(256 * 20)
    for i in range(N):
        full = np.convolve(xs[i].astype(np.float64), hs[i].astype(np.float64)) * (gains[i] or 1.0)
        n = min(full.shape[0], ref.shape[0] - delays[i])
        ref[delays[i]:delays[i] + n] += full[:n]
    assert rel_err(y, ref) <= TOL


# ---------------------------------------------------------------------------
# more size-independent properties at BASELINE.json's full sizes
def test_full_size_c4_512_streams_fused_mix_impulse_sum_and_oracle_sample():
    """C4 per-GPU shard at the named size (512 mono streams x 88200-tap distinct IRs, B=512, P=173, fused mix):
    an impulse into every stream makes the mix the SUM of all 512 IRs; an impulse into one stream isolates its IR;
    and the mix is linear in the inputs."""
    pg.set_sample_rate(wl.SR_441)
    N, L = 512, wl.C4_L
    irs = np.stack([wl.c4_ir(s) for s in range(N)])
    bank = pg.ConvolveBank(irs, N, 1, block=512, max_pull=4096)
    n = 4096 * 2
    x = np.zeros((N, 1, n), np.float32)
    x[:, 0, 0] = 1.0
    y_all = bank.process_mix(x)[0]
    ref = irs[:, :n].astype(np.float64).sum(axis=0)
    assert rel_err(y_all, ref) <= TOL
    bank.reset()
    x[:] = 0.0
    x[137, 0, 5] = 0.5                                   # one stream, delayed and scaled
    y_one = bank.process_mix(x)[0]
    ref1 = np.zeros(n)
    ref1[5:] = 0.5 * irs[137, :n - 5]
    assert rel_err(y_one, ref1) <= TOL
    bank.reset()
    rng = np.random.default_rng(44)
    xa = (rng.uniform(-1, 1, (N, 1, 1024)) / N).astype(np.float32)
    ya = bank.process_mix(xa)[0]
    bank.reset()
    yb = bank.process_mix(2.0 * xa)[0]
    assert rel_err(yb, 2.0 * ya.astype(np.float64)) <= 1e-6
    sample = [3, 200, 511]                                # oracle on a sample of streams: the rest silent
    bank.reset()
    xs = np.zeros((N, 1, 1024), np.float32)
    xs[sample] = xa[sample] * N
    ysamp = bank.process_mix(xs)[0]
    refs = sum(orc.OracleConvolve(irs[s], 1).render(xs[s].T)[:, 0].astype(np.float64) for s in sample)
    assert rel_err(ysamp, refs) <= TOL
    bank.close()


def test_full_size_c3_256_moving_sources_512_taps_vs_direct_convolution():
    """C3 at the named shape (256 sources x 512-tap synthetic HRTF pairs, filter re-selected every 512-pull,
    fused stereo mix): every pull against the float64 direct form of the reference's contract (Appendix A:
    the CURRENT pull's IR applied to the carried history)."""
    pg.set_sample_rate(wl.SR_441)
    N, taps, pulls = 256, 512, 6
    table = wl.c3_synthetic_hrtf_table(taps)
    both = np.concatenate([table, table[:, :, ::-1]], axis=0)
    rng = np.random.default_rng(6)
    traj = rng.integers(0, both.shape[0], (pulls, N)).astype(np.int32)
    x = np.stack([wl.c3_source(pulls * 512, s) for s in range(N)])[:, None, :]
    bank = pg.ConvolveBank(both, N, 1, block=512, max_pull=512, mixdown_input=True, filter_of_stream=traj[0])
    for p in range(pulls):
        bank.set_filter_map(traj[p])
        y = bank.process_mix(np.ascontiguousarray(x[:, :, p * 512:(p + 1) * 512]))
        ref = np.zeros((2, 512))
        lo = max(0, p * 512 - (taps - 1))
        for s in range(N):
            seg = x[s, 0, lo:(p + 1) * 512].astype(np.float64)
            for c in range(2):
                ref[c] += np.convolve(seg, both[traj[p, s], :, c].astype(np.float64))[seg.shape[0] - 512:seg.shape[0]]
        assert rel_err(y, ref) <= TOL, f"pull {p}"
    bank.close()


def test_full_size_c1_many_streams_distinct_fir4096_block4096():
    """C1's throughput shape (B = 4096, P = 1, distinct 4096-tap FIRs): 64 streams, impulse -> own FIR, plus the
    reference's sine input against the oracle for a sample of streams."""
    pg.set_sample_rate(wl.SR_441)
    N = 64
    firs = np.stack([wl.c1_fir(s) for s in range(N)])
    bank = pg.ConvolveBank(firs, N, 1, block=4096, max_pull=4096)
    x = np.zeros((N, 1, 8192), np.float32)
    x[:, 0, 0] = 1.0
    y = np.concatenate([bank.process(np.ascontiguousarray(x[:, :, :4096])),
                        bank.process(np.ascontiguousarray(x[:, :, 4096:]))], axis=2)
    assert rel_err(y[:, 0, :4096], firs) <= TOL and np.max(np.abs(y[:, 0, 4096:])) <= TOL
    bank.reset()
    xs = np.stack([wl.c1_sine(8192, s, N) for s in range(N)])[:, None, :]
    ys = np.concatenate([bank.process(np.ascontiguousarray(xs[:, :, :4096])),
                         bank.process(np.ascontiguousarray(xs[:, :, 4096:]))], axis=2)
    for s in (0, 31, 63):
        ref = orc.OracleConvolve(firs[s], 1).render(xs[s].T)[:, 0]
        assert rel_err(ys[s, 0], ref) <= TOL
    bank.close()


# ---------------------------------------------------------------------------
# WAV staging -- PCM16 converted on the device on the way in and on the way out
def test_wav_reader_feeds_convolve_with_raw_pcm16_frames(tmp_path):
    import wave
    rng = np.random.default_rng(71)
    pcm = rng.integers(-32768, 32768, (5000, 2), dtype=np.int64).astype(np.int16)
    path = str(tmp_path / "in.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(44_100)
        w.writeframes(pcm.astype("<i2").tobytes())
    ir = (rng.standard_normal((600, 2)) / 25).astype(np.float32)
    pulls = [512, 100, 2000, 512, 2475]
    y = _pull_pe(pg.ConvolvePE(pg.WavReaderPE(path), pg.ArrayPE(ir)), pulls)        # int16 H2D, /32768 in HBM
    x = pcm.astype(np.float32) / np.float32(32768.0)
    y_host = _pull_pe(pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(ir)), pulls)          # float32 H2D
    np.testing.assert_array_equal(y, y_host)                                        # the device conversion is exact
    xz = np.concatenate([x, np.zeros((599, 2), np.float32)])
    assert rel_err(y, orc.OracleConvolve(ir, 2).render(xz)) <= TOL


def test_render_to_file_takes_pcm16_off_the_device(tmp_path):
    import wave
    from pygmu2_b200.wav_pe import f32_to_pcm16
    rng = np.random.default_rng(72)
    x = rng.uniform(-1, 1, (9000, 2)).astype(np.float32)
    ir = (rng.standard_normal(500) / 6).astype(np.float32)                          # loud enough to clip sometimes
    path = str(tmp_path / "out.wav")
    pg.render_to_file(pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(ir)), path, chunk=4096)
    with wave.open(path, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (2, 2, 44_100, 9499)
        got = np.frombuffer(w.readframes(9499), dtype="<i2").reshape(-1, 2)
    y = _pull_pe(pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(ir)), [4096, 4096, 1307])
    want = f32_to_pcm16(y)                                                           # host formula on the float32 output
    assert np.any(np.abs(want) == 32767) or np.any(want == -32768)                   # the clip branch is exercised
    np.testing.assert_array_equal(got, want)
    # fused ReverbPE -> file, same route
    path2 = str(tmp_path / "rev.wav")
    pg.render_to_file(pg.ReverbPE(pg.ArrayPE(x), pg.ArrayPE(ir), mix=0.4), path2, chunk=5000)
    yr = _pull_pe(pg.ReverbPE(pg.ArrayPE(x), pg.ArrayPE(ir), mix=0.4), [5000, 4499])
    with wave.open(path2, "rb") as w:
        got2 = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").reshape(-1, 2)
    np.testing.assert_array_equal(got2, f32_to_pcm16(yr))


# ---------------------------------------------------------------------------
# randomised operation sequences: conv and mix pulls of random sizes interleaved with full / per-stream resets
# and filter-map changes, all enqueued without host synchronisation in between (pipelined submit), against a
# direct float64 model of the contract (Appendix A: the CURRENT filter applied to the history since the last reset)
@pytest.mark.parametrize("seed,L,B", [(1, 300, 64), (2, 70, 32), (3, 1000, 128), (4, 40, 64)])
def test_random_operation_sequences_match_direct_model(seed, L, B):
    from pygmu2_b200._lib import PinnedArray
    rng = np.random.default_rng(1000 + seed)
    N, F, C = 3, 4, 2
    h = (rng.standard_normal((F, L, C)) / np.sqrt(L)).astype(np.float32)
    fmap = rng.integers(0, F, N).astype(np.int32)
    bank = pg.ConvolveBank(h, N, 1, block=B, max_pull=4 * B, filter_of_stream=fmap)
    hist = [np.zeros(0, np.float32) for _ in range(N)]
    pending = []          # (ticket, y pinned, reference, label)
    keep = []

    def model(n, mix):
        per = np.zeros((N, C, n))
        for s in range(N):
            seg = hist[s][-(n + L - 1):].astype(np.float64)
            for c in range(C):
                full = np.convolve(seg, h[fmap[s], :, c].astype(np.float64))
                per[s, c] = full[seg.shape[0] - n:seg.shape[0]]
        return per.sum(axis=0) if mix else per

    def drain():
        for tk, yp, ref, label in pending:
            bank.wait(tk)
            # full scale: inputs are uniform(-1,1), filters unit-energy -> outputs O(1); a 1-sample pull can have a
            # tiny maximum by chance, which is not the scale of the signal
            err = float(np.max(np.abs(yp.array.astype(np.float64) - ref)) / max(float(np.max(np.abs(ref))), 0.25))
            assert err <= TOL, label
        pending.clear()

    for step in range(70):
        op = rng.choice(["pull", "pull", "pull", "pull", "mix", "mix", "reset_all", "reset_some", "map"])
        if op in ("pull", "mix"):
            n = int(rng.choice([1, 7, B - 1, B, B + 1, 2 * B, 3 * B + 5, int(rng.integers(1, 4 * B))]))
            x = rng.uniform(-1, 1, (N, 1, n)).astype(np.float32)
            for s in range(N):
                hist[s] = np.concatenate([hist[s], x[s, 0]])
            xp = PinnedArray((N, 1, n))
            xp.array[...] = x
            yp = PinnedArray((C, n) if op == "mix" else (N, C, n))
            tk = bank.submit(xp.array, yp.array, mix=(op == "mix"))
            pending.append((tk, yp, model(n, op == "mix"), f"seed {seed} step {step} {op} n={n}"))
            keep += [xp, yp]
            if len(pending) >= 3:
                drain()
        else:
            drain()   # resets and map changes are ordered against the queue by the library; results checked first
            if op == "reset_all":
                bank.reset()
                hist = [np.zeros(0, np.float32) for _ in range(N)]
            elif op == "reset_some":
                ids = [int(s) for s in range(N) if rng.random() < 0.5] or [0]
                # a per-stream reset clears that stream's history but keeps the block grid: the model is the same
                # because zero history is zero history wherever the grid is
                bank.reset(ids)
                for s in ids:
                    hist[s] = np.zeros(0, np.float32)
            else:
                fmap = rng.integers(0, F, N).astype(np.int32)
                bank.set_filter_map(fmap)
    drain()
    for a in keep:
        a.free()
    bank.close()


# ---------------------------------------------------------------------------
# two-level partitioning (pgx_bank_config.tail_block): head taps at the small block, the rest at a big block.
# Same contract as the uniform bank -- exact linear convolution, zero latency for any pull size.
@pytest.mark.parametrize("L,B,TB", [(3000, 64, 512), (700, 32, 128), (5000, 128, 1024), (513, 64, 512)])
@pytest.mark.parametrize("mix", [False, True])
def test_two_level_partitioning_matches_oracle_and_uniform(L, B, TB, mix):
    rng = np.random.default_rng(300 + L)
    N, C = 3, 2
    h = (rng.standard_normal((N, L, C)) / np.sqrt(L)).astype(np.float32)
    pulls = [B] * 5 + [1, B - 1, 17, TB, TB + 3, 2 * B + 5, B // 2, B // 2, 3 * TB - 7] + [B] * 9
    x = rng.uniform(-1, 1, (N, C, sum(pulls))).astype(np.float32)
    two = pg.ConvolveBank(h, N, C, block=B, tail_block=TB, max_pull=4 * TB)
    uni = pg.ConvolveBank(h, N, C, block=B, max_pull=4 * TB)
    info = two.info()
    assert (info.tail_block, info.block, info.filter_len) == (TB, B, L)
    assert info.partitions == TB // B and info.tail_partitions == -(-(L - TB) // TB)
    y2, y1, pos = [], [], 0
    for d in pulls:
        xc = np.ascontiguousarray(x[:, :, pos:pos + d])
        y2.append(two.process_mix(xc) if mix else two.process(xc))
        y1.append(uni.process_mix(xc) if mix else uni.process(xc))
        pos += d
    y2, y1 = np.concatenate(y2, axis=-1), np.concatenate(y1, axis=-1)
    per = np.stack([orc.OracleConvolve(h[s], C).render(x[s].T).T for s in range(N)])
    ref = per.astype(np.float64).sum(axis=0) if mix else per
    assert rel_err(y2, ref) <= TOL
    assert rel_err(y2, y1) <= 2e-6
    # a reset starts a new run on both levels
    two.reset()
    xc = np.ascontiguousarray(x[:, :, :TB + 50])
    y = two.process_mix(xc) if mix else two.process(xc)
    assert rel_err(y, ref[..., :TB + 50]) <= TOL
    two.close()
    uni.close()


def test_two_level_contract_restrictions_and_c5_shape():
    rng = np.random.default_rng(77)
    ir = (rng.standard_normal(20_000) * np.exp(-np.arange(20_000) / 4000.0) / 30).astype(np.float32)
    x = rng.uniform(-1, 1, 64 * 400).astype(np.float32)
    pe = pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(ir), block_size=64, tail_block=4096)   # C5's shape in miniature
    y = _pull_pe(pe, [64] * 400)[:, 0]
    ref = orc.OracleConvolve(ir, 1).render(x[:, None])[:, 0]
    assert rel_err(y, ref) <= TOL
    bank = pe.bank
    assert bank.info().tail_block == 4096 and bank.info().partitions == 64
    with pytest.raises(ValueError):
        bank.set_filter_map(np.zeros(1, np.int32))
    with pytest.raises(ValueError):
        bank.load_filter(0, ir)
    with pytest.raises(ValueError):                       # mode switch without a reset
        bank.process_mix(np.zeros((1, 1, 64), np.float32))
    short = pg.ConvolveBank(ir[:1000], 1, 1, block=64, tail_block=4096, single_filter_dims=True)
    assert short.info().tail_block == 0                    # the whole filter fits the head: plain uniform bank
    with pytest.raises(ValueError):
        pg.ConvolveBank(ir, 1, 1, block=64, tail_block=3000, single_filter_dims=True)


# ---------------------------------------------------------------------------
# cross-feature checks: two-level banks under output gains / PCM16 / device-resident input, and the host-gather
# path of the fused HRTF mix against the resident-source path
def test_two_level_with_output_gains_pcm16_and_device_voices():
    rng = np.random.default_rng(404)
    L, B, TB, n = 2500, 64, 512, 64 * 30
    h = (rng.standard_normal((1, L, 2)) / np.sqrt(L)).astype(np.float32)
    x = rng.uniform(-1, 1, (1, 2, n)).astype(np.float32)
    two = pg.ConvolveBank(h, 1, 2, block=B, tail_block=TB)
    uni = pg.ConvolveBank(h, 1, 2, block=B)
    for bank in (two, uni):
        bank.set_output_gains(0.6, 0.3)
    y2 = np.concatenate([two.process(np.ascontiguousarray(x[:, :, p:p + 96])) for p in range(0, n, 96)], axis=2)
    y1 = np.concatenate([uni.process(np.ascontiguousarray(x[:, :, p:p + 96])) for p in range(0, n, 96)], axis=2)
    wet = orc.OracleConvolve(h[0], 2).render(x[0].T).T
    assert rel_err(y2[0], 0.3 * x[0] + 0.6 * wet) <= TOL and rel_err(y2, y1) <= 2e-6
    # int16 out of a two-level single-stream bank
    two.reset()
    from pygmu2_b200.wav_pe import f32_to_pcm16
    xi = np.ascontiguousarray(x[0].T[:1000])
    pcm = two.process_interleaved(xi, pcm16_out=True)
    two.reset()
    np.testing.assert_array_equal(pcm, f32_to_pcm16(two.process_interleaved(xi)))
    # device-resident voices into a two-level ConvolvePE (the c5v --tail-block route), against the uniform one
    ir = (rng.standard_normal(3000) * np.exp(-np.arange(3000) / 600.0) / 12).astype(np.float32)
    def graph(tail):
        voices = [pg.SuperSawPE(frequency=80.0 * 2 ** (i / 12.0), amplitude=1.0 / 8, seed=i) for i in range(8)]
        return pg.ConvolvePE(pg.MixPE(*voices), pg.ArrayPE(ir), block_size=64, tail_block=tail)
    ya = _pull_pe(graph(512), [64] * 40)
    yb = _pull_pe(graph(None), [64] * 40)
    assert rel_err(ya, yb) <= 2e-6 and np.max(np.abs(yb)) > 0.01


class _Opaque(pg.ProcessingElement):
    """Hides an ArrayPE behind a custom PE so the fused mix has to gather on the host."""

    def __init__(self, inner):
        self._inner = inner

    def inputs(self):
        return [self._inner]

    def is_pure(self):
        return True

    def channel_count(self):
        return self._inner.channel_count()

    def _compute_extent(self):
        return self._inner.extent()

    def _render(self, start, duration):
        return self._inner.render(start, duration)


def test_fused_hrtf_mix_host_gather_equals_resident_sources():
    rng = np.random.default_rng(405)
    xs = [rng.uniform(-1, 1, 1500 + 200 * i).astype(np.float32) / 4 for i in range(5)]
    az = [-150.0, -40.0, 0.0, 75.0, 180.0]

    def mix(wrap):
        ins = []
        for i, (x, a) in enumerate(zip(xs, az)):
            src = pg.ArrayPE(x)
            sp = pg.SpatialPE(_Opaque(src) if wrap else src, method=pg.SpatialHRTF(a, 10.0 * i))
            ins.append(pg.DelayPE(sp, 100 * i) if i % 2 else sp)
        return pg.MixPE(*ins)

    res, host = mix(False), mix(True)
    yr, yh = _pull_pe(res, [512] * 6), _pull_pe(host, [512] * 6)
    assert res._fused._resident is not None and host._fused._resident is None
    np.testing.assert_array_equal(yr, yh)


# ---------------------------------------------------------------------------
# time-tiled accumulate passes (PGX_TILE): one pass over the delay line per `tile` blocks
@pytest.mark.gpu
@pytest.mark.parametrize("tile", ["2", "4"])
@pytest.mark.parametrize("shared", [True, False])
def test_time_tiled_pass_matches_per_block_pass_and_float64(tile, shared, monkeypatch):
    """The same bank with the per-block pass and with the tiled pass (forced on a small bank), pulled in ragged
    chunks: the tiled outputs against the float64 convolution, and against the per-block schedule far below the
    parity tolerance."""
    pg.set_sample_rate(wl.SR_48)
    rng = np.random.default_rng(int(tile) * 10 + shared)
    B, P, N, c = 256, 37, 5, 2
    L = B * P - 19
    nf = 1 if shared else 3
    h = (rng.standard_normal((nf, L, c)) * np.exp(-np.arange(L) / (L / 5))[None, :, None]).astype(np.float32)
    fmap0 = np.zeros(N, np.int32) if shared else (np.arange(N) % nf).astype(np.int32)
    pulls = [256, 256, 100, 156, 512, 700, 68, 256] + [256] * 44 + [1024, 33, 223] + [256] * 12
    n = sum(pulls)
    x = rng.uniform(-1, 1, (N, c, n)).astype(np.float32)

    def run(env_tile):
        monkeypatch.setenv("PGX_TILE", env_tile)
        monkeypatch.setenv("PGX_TILE_MIN", "0")
        bank = pg.ConvolveBank(h[0] if shared else h, N, c, block=B, single_filter_dims=shared, max_pull=1024,
                               filter_of_stream=None if shared else fmap0)
        assert bank.info().mac_tile == int(env_tile)
        out, pos = [], 0
        for d in pulls:
            out.append(bank.process(np.ascontiguousarray(x[:, :, pos:pos + d])).copy())
            pos += d
        bank.close()
        return np.concatenate(out, axis=2)

    y1 = run("1")
    yt = run(tile)
    for s in range(N):
        for ch in range(c):
            ref = _fftconv64(x[s, ch], h[fmap0[s], :, ch], n)
            assert rel_err(yt[s, ch], ref) <= TOL, (s, ch)
    assert rel_err(yt, y1) <= 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("tile", ["2", "4"])
def test_time_tiled_pass_random_operation_sequence(tile, monkeypatch):
    """Resets, filter-map changes, filter reloads and conv/mix switches at arbitrary points of a tiled bank: the
    outputs equal those of the per-block schedule driven through the same sequence."""
    pg.set_sample_rate(wl.SR_48)
    B, P, N, c = 256, 33, 4, 1
    L = B * P
    rng0 = np.random.default_rng(5)
    h = (rng0.standard_normal((3, L, c)) * np.exp(-np.arange(L) / (L / 4))[None, :, None]).astype(np.float32)
    h2 = (rng0.standard_normal((L, c)) * 0.05).astype(np.float32)

    def run(env_tile):
        monkeypatch.setenv("PGX_TILE", env_tile)
        monkeypatch.setenv("PGX_TILE_MIN", "0")
        rng = np.random.default_rng(99)
        bank = pg.ConvolveBank(h, N, c, block=B, max_pull=600, filter_of_stream=np.arange(N, dtype=np.int32) % 3)
        assert bank.info().mac_tile == int(env_tile)
        outs = []
        for step in range(140):
            op = rng.integers(0, 12)
            if op == 0:
                bank.reset([int(rng.integers(0, N))])
            elif op == 1:
                bank.set_filter_map(rng.integers(0, 3, N).astype(np.int32))
            elif op == 2 and step > 60:
                bank.load_filter(1, h2)
            d = int(rng.choice([256, 256, 256, 512, 600, 17, 239, 100]))
            xx = rng.uniform(-1, 1, (N, c, d)).astype(np.float32)
            if op == 3:
                outs.append(bank.process_mix(xx).copy().ravel())
            else:
                outs.append(bank.process(xx).copy().ravel())
        bank.close()
        return np.concatenate(outs)

    y1 = run("1")
    yt = run(tile)
    assert rel_err(yt, y1) <= 2e-6
