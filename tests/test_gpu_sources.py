"""
Device-resident sources through the C ABI (pgx_osc_*): SinePE, BlitSawPE, SuperSawPE
against golden outputs of the REAL reference (tests/golden/src_oscillators.npz) and the oracle restatement,
and the device-to-device hand-over into ConvolvePE / MixPE / ConvolveBank.  Needs a B200: ``-m gpu``.

Tolerance: the kernels use float64 like the reference and round to float32 at the same places; what differs is
the libm (CUDA sin vs glibc) and the order of the phase / integrator sums, so outputs agree to ~1 float32 ulp.
The bar stays the path's: max-abs error <= 1e-5 of full scale.
"""
import numpy as np
import pytest

import pygmu2_b200 as pg
import pygmu2_oracle as orc
import pygmu2_oracle_sources as osrc
from conftest import golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5
SR = 44_100
MIX = {0: "center_heavy", 1: "linear", 2: "equal"}


def _pull(pe, pulls, start=0):
    out, pos = [], start
    for d in pulls:
        out.append(pe.render(pos, int(d)).data.copy())
        pos += int(d)
    return np.concatenate(out, axis=0)


def test_sine_matches_reference_goldens():
    g = golden("src_oscillators.npz")
    pulls = g["pulls"]
    (f0, a0, p0), (f1, a1, p1), (f2, a2, p2) = g["sine_params"]
    y = _pull(pg.SinePE(frequency=f0), pulls)
    assert y.shape == g["sine_440"].shape and rel_err(y, g["sine_440"]) <= 2e-7
    yb = _pull(pg.SinePE(frequency=f1, amplitude=a1, phase=p1, channels=2), pulls, start=4_000_000)
    assert yb.shape == g["sine_b"].shape and rel_err(yb, g["sine_b"]) <= 2e-7
    yc = _pull(pg.SinePE(frequency=f2, amplitude=a2, phase=p2), pulls, start=-300)
    assert rel_err(yc, g["sine_c"]) <= 2e-7


def test_blit_saw_matches_reference_goldens_including_restart():
    g = golden("src_oscillators.npz")
    pulls = g["pulls"]
    for i, (f, a, p, m, lk) in enumerate(g["blit_cases"]):
        pe = pg.BlitSawPE(frequency=f, amplitude=a, initial_phase=p, m=None if m < 0 else int(m), leak=lk)
        y = np.concatenate([_pull(pe, pulls), _pull(pe, [64, 64], start=10_000)])
        ref = g[f"blit_{i}"]
        assert y.shape == ref.shape
        # full scale of a BlitSaw is its amplitude (case 5, 10 kHz: M = 1, the output is pure rounding noise ~1e-15)
        err = float(np.max(np.abs(y.astype(np.float64) - ref)) / max(float(np.max(np.abs(ref))), a))
        assert err <= TOL, f"case {i}: {err:.3e}"


def test_supersaw_matches_reference_goldens():
    g = golden("src_oscillators.npz")
    pulls = g["pulls"]
    for i, (f, a, v, d, mm, rp, sd) in enumerate(g["ssaw_cases"]):
        pe = pg.SuperSawPE(frequency=f, amplitude=a, voices=int(v), detune_cents=d, mix_mode=MIX[int(mm)],
                           randomize_phase=bool(rp), seed=int(sd))
        y = _pull(pe, pulls)
        assert rel_err(y, g[f"ssaw_{i}"]) <= TOL, f"case {i}: {rel_err(y, g[f'ssaw_{i}']):.3e}"


def test_supersaw_long_pull_and_start_stop_reset():
    """4096-sample pulls (32 warp tiles per launch) and the on_start reset, against the oracle."""
    pe = pg.SuperSawPE(frequency=523.25, amplitude=0.8, voices=9, detune_cents=25.0, seed=11, channels=2)
    o = osrc.OracleSuperSaw(523.25, 0.8, 9, 25.0, seed=11, sample_rate=SR)
    with pg.NullRenderer(sample_rate=SR) as r:
        r.set_source(pe)
        r.start()
        y = _pull(pe, [4096, 4096, 1000])
        ref = np.concatenate([o.render(0, 4096), o.render(4096, 4096), o.render(8192, 1000)])
        assert y.shape == (9192, 2)
        assert rel_err(y[:, 0], ref) <= TOL and np.array_equal(y[:, 0], y[:, 1])
        r.stop()
        r.start()                                   # restart: phases and integrators back to their initial values
        o.reset()
        assert rel_err(_pull(pe, [300])[:, 0], o.render(0, 300)) <= TOL


def test_voice_mix_fused_matches_reference_golden():
    """C5 front end in miniature: MixPE of 16 SuperSaw voices = one VoiceBank launch + the float32 voice sum."""
    g = golden("src_oscillators.npz")
    voices = [pg.SuperSawPE(frequency=110.0 * 2 ** (i / 12.0), amplitude=1.0 / 16, seed=i) for i in range(16)]
    mix = pg.MixPE(*voices)
    y = _pull(mix, [64] * 12)
    assert rel_err(y, g["c5_voicemix16"]) <= TOL
    from pygmu2_b200.mix_pe import _VoiceMix
    assert isinstance(mix._fused, _VoiceMix) and mix._fused.vb.bank.launches == 2 * 12


def test_pe_valued_parameters_are_inputs_like_in_the_reference():
    assert not pg.SinePE(frequency=pg.ConstantPE(440.0)).is_pure()
    assert pg.SuperSawPE(frequency=pg.ConstantPE(440.0)).inputs()
    m = pg.ConstantPE(5.0)
    assert pg.BlitSawPE(frequency=440.0, m=m).inputs() == [m]


def test_modulated_blit_and_supersaw_match_reference_goldens():
    """BlitSawPE / SuperSawPE with PE-valued frequency and / or amplitude (blit_saw_pe.py:161-262,
    super_saw_pe.py:223-246,287-303) against outputs of the REAL reference: vibrato, a six-octave glide (the harmonic
    count and the period change from sample to sample), an amplitude envelope, fixed m, stereo, and a non-contiguous
    pull.  Same tolerance as the constant-parameter oscillators (1e-5 of full scale)."""
    g = golden("src_modulated.npz")
    pulls = [int(d) for d in g["pulls"]]
    A = lambda k: pg.ArrayPE(g["ctl_" + k])  # noqa: E731

    def check(y, ref, scale, what):
        assert y.shape == ref.shape, what
        err = float(np.max(np.abs(y.astype(np.float64) - ref)) / max(float(np.max(np.abs(ref))), scale))
        assert err <= TOL, f"{what}: {err:.3e}"

    check(_pull(pg.BlitSawPE(frequency=A("vib_freq"), amplitude=0.7, initial_phase=0.3), pulls), g["blit_vib"], 0.7, "blit_vib")
    check(_pull(pg.BlitSawPE(frequency=A("glide_freq"), amplitude=A("env_amp")), pulls), g["blit_glide_env"], 0.8, "blit_glide_env")
    check(_pull(pg.BlitSawPE(frequency=330.0, amplitude=A("env_amp"), m=12, leak=0.995), pulls), g["blit_amp_only_m12"], 0.8, "m12")
    check(_pull(pg.SuperSawPE(frequency=A("vib_freq"), amplitude=0.5, seed=3), pulls), g["ssaw_vib"], 0.5, "ssaw_vib")
    check(_pull(pg.SuperSawPE(frequency=A("glide_freq"), amplitude=A("env_amp"), voices=5, detune_cents=35.0,
                              mix_mode="linear", channels=2, seed=4), pulls), g["ssaw_glide_env_stereo"], 0.8, "ssaw_glide")
    check(_pull(pg.SuperSawPE(frequency=110.0, amplitude=A("env_amp"), seed=5), pulls), g["ssaw_amp_only"], 0.8, "ssaw_amp")
    check(_pull(pg.BlitSawPE(frequency=A("vib_freq"), amplitude=0.6, m=A("m_steps")), pulls), g["blit_m_pe"], 0.6, "blit_m_pe")
    pe = pg.BlitSawPE(frequency=A("vib_freq"), amplitude=A("env_amp"))
    y = np.concatenate([_pull(pe, [256, 256]), _pull(pe, [256, 256], start=1024)])
    check(y, g["blit_gap"], 0.8, "blit_gap")
    assert pe.extent().end == g["ctl_vib_freq"].shape[0] and not pe.is_pure()


def test_c1_sine_source_stays_on_device_through_convolve():
    """SinePE -> ConvolvePE with the samples handed over in HBM (no H2D): same answer as the host hand-over."""
    rng = np.random.default_rng(3)
    h = (rng.standard_normal(700) / 20).astype(np.float32)
    pe = pg.ConvolvePE(pg.SinePE(frequency=440.0), pg.ArrayPE(h))
    pulls = [512, 17, 512, 1000, 259]
    y = _pull(pe, pulls)
    x = osrc.sine(440.0, 1.0, 0.0, SR, 0, sum(pulls))
    ref = orc.OracleConvolve(h, 1).render(x[:, None])
    assert rel_err(y, ref) <= TOL
    host = pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(h))
    assert rel_err(y, _pull(host, pulls)) <= 1e-6


def test_c5_pipeline_voices_to_long_ir_all_on_device():
    """MixPE(64 SuperSaw voices) -> ConvolvePE(2000-tap IR) at 64-sample pulls: voices, voice sum and the
    convolution input never leave the device; checked against oracle voices -> float32 mix -> oracle convolve."""
    V, n_pulls = 64, 40
    rng = np.random.default_rng(8)
    ir = (rng.standard_normal(2000) * np.exp(-np.arange(2000) / 333.0) / 10).astype(np.float32)
    voices = [pg.SuperSawPE(frequency=55.0 * 2 ** (i / 24.0), amplitude=1.0 / 32, seed=i) for i in range(V)]
    pe = pg.ConvolvePE(pg.MixPE(*voices), pg.ArrayPE(ir), block_size=64)
    y = _pull(pe, [64] * n_pulls)
    ov = [osrc.OracleSuperSaw(55.0 * 2 ** (i / 24.0), 1.0 / 32, seed=i, sample_rate=SR) for i in range(V)]
    xs = []
    for p in range(n_pulls):
        acc = ov[0].render(p * 64, 64).copy()
        for o in ov[1:]:
            acc += o.render(p * 64, 64)
        xs.append(acc)
    ref = orc.OracleConvolve(ir, 1).render(np.concatenate(xs)[:, None])
    assert rel_err(y, ref) <= TOL


def test_bank_with_device_voice_source_lockstep():
    """N independent SinePE streams -> N distinct FIRs: one VoiceBank feeds the ConvolveBank in HBM."""
    N, L = 6, 300
    rng = np.random.default_rng(12)
    hs = (rng.standard_normal((N, L)) / 15).astype(np.float32)
    sines = [pg.SinePE(frequency=200.0 * (i + 1), amplitude=0.5, phase=0.1 * i) for i in range(N)]
    bank = pg.ConvolveBank(hs, N, 1, pull_hint=256)
    bank.attach_device_source(pg.VoiceBank(sines))
    ys = [bank.render(p, 256) for p in range(0, 1024, 256)]
    y = np.concatenate(ys, axis=2)
    for s in range(N):
        x = osrc.sine(200.0 * (s + 1), 0.5, 0.1 * s, SR, 0, 1024)
        assert rel_err(y[s, 0], orc.OracleConvolve(hs[s], 1).render(x[:, None])[:, 0]) <= TOL


def test_modulated_sine_fm_am_pm_matches_reference_golden():
    """SinePE with PE-valued frequency / amplitude / phase (sine_pe.py:134-232) on the device against outputs of the
    REAL reference: the float64 phase is walked left to right like np.cumsum and carried between pulls, so the result
    may differ from the reference only where CUDA's and the host libm's float64 sin differ in the last bit before the
    float32 rounding: <= 1 float32 ulp, and exact almost everywhere."""
    g = golden("src_modulated.npz")
    pulls = [int(d) for d in g["pulls"]]
    A = lambda k: pg.ArrayPE(g["ctl_" + k])  # noqa: E731

    def close(y, ref):
        assert y.shape == ref.shape
        ulp = np.spacing(np.maximum(np.abs(ref), np.float32(1e-3)).astype(np.float32))
        assert np.all(np.abs(y.astype(np.float64) - ref.astype(np.float64)) <= ulp), float(np.max(np.abs(y - ref)))
        assert np.mean(y == ref) > 0.98

    close(_pull(pg.SinePE(frequency=A("fm_freq")), pulls), g["fm"])
    close(_pull(pg.SinePE(frequency=A("fm_freq"), amplitude=0.5, phase=0.7), pulls), g["fm_phase0p7_amp0p5"])
    close(_pull(pg.SinePE(frequency=1000.0, amplitude=A("am_amp"), phase=0.25), pulls), g["am"])
    close(_pull(pg.SinePE(frequency=220.0, phase=A("pm_phase")), pulls), g["pm"])
    close(_pull(pg.SinePE(frequency=A("sweep_freq"), amplitude=A("am_amp"), phase=A("pm_phase"), channels=2), pulls),
          g["all3_stereo"])
    pe = pg.SinePE(frequency=A("fm_freq"), phase=A("pm_phase"))
    assert not pe.is_pure() and len(pe.inputs()) == 2 and pe.extent().end == g["ctl_fm_freq"].shape[0]
    r = pg.NullRenderer(sample_rate=44_100)
    r.set_source(pe)
    r.start()
    a = _pull(pe, [256, 256])
    r.stop()
    r.start()
    b = _pull(pe, [256, 256])
    r.stop()
    close(np.concatenate([a, b]), g["restart"])


def test_modulated_sine_feeds_convolve_on_the_device():
    """An FM SinePE as the source of a ConvolvePE: the modulated block is produced in HBM on the bank's stream and
    consumed there (device_block protocol); against the oracle chain."""
    import pygmu2_oracle as orc
    import pygmu2_oracle_sources as osrc
    g = golden("src_modulated.npz")
    rng = np.random.default_rng(5)
    h = (rng.standard_normal(300) / 17).astype(np.float32)
    pe = pg.ConvolvePE(pg.SinePE(frequency=pg.ArrayPE(g["ctl_fm_freq"]), amplitude=0.5), pg.ArrayPE(h), block_size=256)
    y = _pull(pe, [256] * 8)
    o, c = osrc.OracleSineModulated(amplitude=0.5), orc.OracleConvolve(h, 1)
    ref = np.concatenate([c.render(o.render(256, g["ctl_fm_freq"][i * 256:(i + 1) * 256])) for i in range(8)])
    assert np.max(np.abs(y - ref)) <= 1e-5 * np.max(np.abs(ref))


def test_speculative_voice_blocks_are_invisible_in_the_output(monkeypatch):
    """ConvolvePE over a MixPE of SuperSaw voices renders the NEXT voice block behind every pull (so that the 20 us
    voice front end is off the latency chain of the pull that asks for it).  The oscillator state is snapshotted and
    rolled back when the next pull is a different one: outputs must equal, BIT FOR BIT, those with speculation off --
    through hits, a pull of another size, a non-contiguous pull, a stop/start -- and match the oracle chain."""
    import pygmu2_oracle as orc
    import pygmu2_oracle_sources as osrc
    rng = np.random.default_rng(2)
    h = (rng.standard_normal(900) / 30).astype(np.float32)
    seq = [(0, 64), (64, 64), (128, 64), (192, 32), (224, 64), (288, 64), (1000, 64), (1064, 64), (1128, 17), (1145, 64)]

    def run(spec):
        monkeypatch.setenv("PGX_SPECULATE", spec)
        voices = [pg.SuperSawPE(frequency=110.0 * 2 ** (i / 12.0), amplitude=1.0 / 8, seed=i) for i in range(8)]
        pe = pg.ConvolvePE(pg.MixPE(*voices), pg.ArrayPE(h), block_size=64)
        r = pg.NullRenderer(sample_rate=SR)
        r.set_source(pe)
        r.start()
        ys = [pe.render(s_, d).data.copy() for s_, d in seq]
        r.stop()
        r.start()
        ys += [pe.render(s_, d).data.copy() for s_, d in seq[:4]]
        r.stop()
        return ys, pe

    y1, pe1 = run("1")
    y0, pe0 = run("0")
    for a, b in zip(y1, y0):
        assert np.array_equal(a, b)
    assert pe1._speculate and not pe0._speculate
    assert pe1._spec is None            # stop() dropped the outstanding block and rolled the oscillators back
    # the oracle chain for the first contiguous run
    vo = [osrc.OracleSuperSaw(110.0 * 2 ** (i / 12.0), 1.0 / 8, seed=i, sample_rate=SR) for i in range(8)]
    c = orc.OracleConvolve(h, 1)
    ref = []
    for s_, d in seq[:6]:
        x = osrc.oracle_mix([v.render(s_, d)[:, None] for v in vo]) if hasattr(osrc, "oracle_mix") else \
            orc.oracle_mix([v.render(s_, d)[:, None] for v in vo])
        ref.append(c.render(x))
    got, want = np.concatenate(y1[:6]), np.concatenate(ref)
    assert np.max(np.abs(got - want)) <= 1e-5 * np.max(np.abs(want))
