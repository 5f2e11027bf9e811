"""
The source-PE oracle (oracle/pygmu2_oracle_sources.py) against golden outputs of the REAL reference
(tests/golden/src_oscillators.npz, made by oracle/gen_golden_sources.py).  CPU only.
Sine / BlitSaw / SuperSaw restatements use the same numpy / scipy calls as the reference, so the
float32 outputs must be bit-identical.
"""
import numpy as np

import pygmu2_oracle_sources as osrc
from conftest import golden

SR = 44_100
MIX = {0: "center_heavy", 1: "linear", 2: "equal"}


def _pull(render, pulls, start=0):
    out, pos = [], start
    for d in pulls:
        out.append(render(pos, int(d)))
        pos += int(d)
    return np.concatenate(out)


def test_sine_bit_exact():
    g = golden("src_oscillators.npz")
    pulls = g["pulls"]
    (f0, a0, p0), (f1, a1, p1), (f2, a2, p2) = g["sine_params"]
    np.testing.assert_array_equal(_pull(lambda s, d: osrc.sine(f0, a0, p0, SR, s, d), pulls), g["sine_440"][:, 0])
    yb = _pull(lambda s, d: osrc.sine(f1, a1, p1, SR, s, d), pulls, start=4_000_000)
    np.testing.assert_array_equal(yb, g["sine_b"][:, 0])
    np.testing.assert_array_equal(yb, g["sine_b"][:, 1])
    np.testing.assert_array_equal(_pull(lambda s, d: osrc.sine(f2, a2, p2, SR, s, d), pulls, start=-300), g["sine_c"][:, 0])


def test_blit_saw_bit_exact_including_restart():
    g = golden("src_oscillators.npz")
    pulls = g["pulls"]
    for i, (f, a, p, m, lk) in enumerate(g["blit_cases"]):
        o = osrc.OracleBlitSaw(f, a, p, None if m < 0 else int(m), lk, SR)
        y = np.concatenate([_pull(o.render, pulls), _pull(o.render, [64, 64], start=10_000)])
        np.testing.assert_array_equal(y, g[f"blit_{i}"][:, 0], err_msg=f"case {i}")


def test_supersaw_bit_exact():
    g = golden("src_oscillators.npz")
    pulls = g["pulls"]
    for i, (f, a, v, d, mm, rp, sd) in enumerate(g["ssaw_cases"]):
        o = osrc.OracleSuperSaw(f, a, int(v), d, MIX[int(mm)], bool(rp), int(sd), SR)
        np.testing.assert_array_equal(_pull(o.render, pulls), g[f"ssaw_{i}"][:, 0], err_msg=f"case {i}")


def test_c5_voice_mix_bit_exact():
    g = golden("src_oscillators.npz")
    voices = [osrc.OracleSuperSaw(110.0 * 2 ** (i / 12.0), 1.0 / 16, seed=i, sample_rate=SR) for i in range(16)]
    out, pos = [], 0
    for _ in range(12):
        acc = voices[0].render(pos, 64).copy()
        for v in voices[1:]:
            acc += v.render(pos, 64)            # MixPE: float32 left-to-right (mix_pe.py:92-94)
        out.append(acc)
        pos += 64
    np.testing.assert_array_equal(np.concatenate(out), g["c5_voicemix16"][:, 0])


def test_oracle_modulated_sine_is_the_reference_bit_for_bit():
    """OracleSineModulated against tests/golden/src_modulated.npz (real reference, oracle/gen_golden_modulated.py)."""
    import pygmu2_oracle_sources as osrc
    g = golden("src_modulated.npz")
    pulls = [int(d) for d in g["pulls"]]

    def run(o, f=None, a=None, p=None, pulls=pulls):
        out, pos = [], 0
        for d in pulls:
            sl = slice(pos, pos + d)
            out.append(o.render(d, None if f is None else g[f][sl], None if a is None else g[a][sl],
                                None if p is None else g[p][sl]))
            pos += d
        return np.concatenate(out)
    assert np.array_equal(run(osrc.OracleSineModulated(), f="ctl_fm_freq"), g["fm"])
    assert np.array_equal(run(osrc.OracleSineModulated(amplitude=0.5, phase=0.7), f="ctl_fm_freq"), g["fm_phase0p7_amp0p5"])
    assert np.array_equal(run(osrc.OracleSineModulated(frequency=1000.0, phase=0.25), a="ctl_am_amp"), g["am"])
    assert np.array_equal(run(osrc.OracleSineModulated(frequency=220.0), p="ctl_pm_phase"), g["pm"])
    assert np.array_equal(run(osrc.OracleSineModulated(channels=2), f="ctl_sweep_freq", a="ctl_am_amp", p="ctl_pm_phase"),
                          g["all3_stereo"])
    o = osrc.OracleSineModulated()
    a = run(o, f="ctl_fm_freq", p="ctl_pm_phase", pulls=[256, 256])
    o.reset()
    b = run(o, f="ctl_fm_freq", p="ctl_pm_phase", pulls=[256, 256])
    assert np.array_equal(np.concatenate([a, b]), g["restart"])


def test_oracle_modulated_blit_and_supersaw_are_the_reference_bit_for_bit():
    import pygmu2_oracle_sources as osrc
    g = golden("src_modulated.npz")
    pulls = [int(d) for d in g["pulls"]]

    def run(o, f=None, a=None, pulls=pulls, start=0, off=0):
        out, pos = [], 0
        for d in pulls:
            sl = slice(off + pos, off + pos + d)
            out.append(o.render(start + pos, d, None if f is None else g[f][sl], None if a is None else g[a][sl]))
            pos += d
        return np.concatenate(out)
    assert np.array_equal(run(osrc.OracleBlitSaw(1.0, 0.7, 0.3), f="ctl_vib_freq"), g["blit_vib"][:, 0])
    assert np.array_equal(run(osrc.OracleBlitSaw(1.0, 1.0), f="ctl_glide_freq", a="ctl_env_amp"), g["blit_glide_env"][:, 0])
    assert np.array_equal(run(osrc.OracleBlitSaw(330.0, 1.0, m=12, leak=0.995), a="ctl_env_amp"), g["blit_amp_only_m12"][:, 0])
    assert np.array_equal(run(osrc.OracleSuperSaw(1.0, 0.5, seed=3), f="ctl_vib_freq"), g["ssaw_vib"][:, 0])
    y = run(osrc.OracleSuperSaw(1.0, 1.0, voices=5, detune_cents=35.0, mix_mode="linear", seed=4),
            f="ctl_glide_freq", a="ctl_env_amp")
    assert np.array_equal(np.stack([y, y], axis=1), g["ssaw_glide_env_stereo"])
    assert np.array_equal(run(osrc.OracleSuperSaw(110.0, 1.0, seed=5), a="ctl_env_amp"), g["ssaw_amp_only"][:, 0])
    ob, out, pos = osrc.OracleBlitSaw(1.0, 0.6), [], 0
    for d in pulls:
        out.append(ob.render(pos, d, g["ctl_vib_freq"][pos:pos + d], None, g["ctl_m_steps"][pos:pos + d]))
        pos += d
    assert np.array_equal(np.concatenate(out), g["blit_m_pe"][:, 0])
    o = osrc.OracleBlitSaw(1.0, 1.0)
    a = run(o, f="ctl_vib_freq", a="ctl_env_amp", pulls=[256, 256])
    b = run(o, f="ctl_vib_freq", a="ctl_env_amp", pulls=[256, 256], start=1024, off=1024)
    assert np.array_equal(np.concatenate([a, b]), g["blit_gap"][:, 0])
