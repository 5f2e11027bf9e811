#!/usr/bin/env python
"""
Generate tests/golden/*.npz by running the REAL Python reference (rdpoor/pygmu2
at /root/reference) in the build container.  Test infrastructure only.

    python oracle/gen_golden.py            # needs /root/reference; rewrites tests/golden/

The reference is pure Python and cannot travel to the GPU box, so its outputs
on the seeded inputs of ``pygmu2_b200.workloads`` are committed here as small
fixtures.  Inputs are never stored: tests regenerate them from the same seeds.
The script also packs the reference's KEMAR compact HRTF WAVs (data, MIT Media
Lab, Gardner & Martin 1994) into ``pygmu2_b200/assets/kemar_compact_i16.npz``
so SpatialHRTF has its table on a machine without pygmu2 installed, and checks
that the programmatic KEMAR grid equals the reference's literal table.
"""
from __future__ import annotations

import os
import sys
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("PYGMU2_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "stubs"))
sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import pygmu2 as pg  # noqa: E402  (the real reference)

from pygmu2_b200 import workloads as wl  # noqa: E402
import pygmu2_oracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
ASSETS = os.path.join(ROOT, "pygmu2_b200", "assets")


def _pull_all(pe, pulls, start=0):
    out = []
    pos = start
    for d in pulls:
        out.append(pe.render(pos, d).data.copy())
        pos += d
    return np.concatenate(out, axis=0)


def gen_kemar():
    entries = pg.SpatialHRTF.KEMAR_HRTF_ENTRIES
    mine = orc.kemar_entries()
    assert len(entries) == len(mine) == 368
    for a, b in zip(entries, mine):
        assert (int(a[0]), int(a[1]), a[2]) == b, (a, b)
    from pygmu2.assets import get_kemar_dir
    kdir = get_kemar_dir()
    table = np.zeros((368, 128, 2), dtype=np.int16)
    for i, (_, _, fn) in enumerate(entries):
        with wave.open(str(kdir / fn), "rb") as w:
            assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (2, 2, 44100, 128)
            table[i] = np.frombuffer(w.readframes(128), dtype="<i2").reshape(128, 2)
    os.makedirs(ASSETS, exist_ok=True)
    np.savez_compressed(os.path.join(ASSETS, "kemar_compact_i16.npz"), ir_i16=table,
                        sample_rate=np.int32(44100))
    # nearest-neighbour lookups incl. ties / clamping / negative azimuth
    rng = np.random.default_rng(11)
    az = np.concatenate([rng.uniform(-200, 200, 400), np.array([0, 45, -45, 90, 180, -180, 2.5, 7.5, 177.5, 3.0])])
    el = np.concatenate([rng.uniform(-60, 100, 400), np.array([0, 0, 0, 0, 0, 0, 0, 5.0, 85.0, 45.0])])
    names = [pg.SpatialHRTF.hrtf_filename_for(a, e) for a, e in zip(az, el)]
    idx = np.array([[e[2] for e in entries].index(n) for n in names], dtype=np.int32)
    np.savez_compressed(os.path.join(GOLD, "hrtf_lookup.npz"), az=az, el=el, idx=idx)
    return table.astype(np.float32) / 32768.0


def gen_unit_vectors():
    """The reference's own small known-answer cases (tests/test_convolve_pe.py) run through the reference."""
    pg.set_sample_rate(10_000)
    out = {}
    x = np.array([1, 2, 3, 4], dtype=np.float32)
    h = np.array([1, 0.5, -1], dtype=np.float32)
    out["mono_small"] = pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(h), fft_size=16).render(0, 6).data
    xs = np.array([[1, 10], [2, 20], [3, 30], [4, 40]], dtype=np.float32)
    out["stereo_monofilter"] = pg.ConvolvePE(pg.ArrayPE(xs), pg.ArrayPE(np.array([1, -1], np.float32)),
                                             fft_size=16).render(0, 5).data
    h2 = np.stack([np.array([1.0, 0.5], np.float32), np.array([-1.0, 0.5], np.float32)], axis=1)
    out["fanout"] = pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(h2), fft_size=16).render(0, 5).data
    rng = np.random.default_rng(0)
    xr = rng.normal(size=200).astype(np.float32)
    hr = np.array([0.25, 0.5, 0.25], dtype=np.float32)
    pe = pg.ConvolvePE(pg.ArrayPE(xr), pg.ArrayPE(hr), fft_size=64)
    out["chunked"] = _pull_all(pe, (17, 23, 19, 41, 7, 93, 2))
    np.savez_compressed(os.path.join(GOLD, "convolve_unit.npz"), **out)


def gen_ragged():
    """Ragged pulls, a non-contiguous jump (history reset), a pull past the extent; stereo src x stereo filter."""
    pg.set_sample_rate(44_100)
    rng = np.random.default_rng(77)
    x = rng.uniform(-1, 1, (6000, 2)).astype(np.float32)
    h = (rng.standard_normal((1000, 2)) * np.exp(-np.arange(1000)[:, None] / 200.0) / 10).astype(np.float32)
    pe = pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(h))
    pulls_a = (1, 17, 511, 512, 513, 64, 1000, 3)
    ya = _pull_all(pe, pulls_a, start=0)
    # non-contiguous: jump back to 2000 -> tail cleared (convolve_pe.py:255-256)
    pulls_b = (700, 5, 2048, 1500, 900)  # runs past src end (6000) and past extent end (6999)
    yb = _pull_all(pe, pulls_b, start=2000)
    np.savez_compressed(os.path.join(GOLD, "convolve_ragged.npz"), x=x, h=h,
                        pulls_a=np.array(pulls_a), pulls_b=np.array(pulls_b), start_b=np.int64(2000), ya=ya, yb=yb)


def gen_c1(n=4410):
    pg.set_sample_rate(wl.SR_441)
    src = pg.SinePE(frequency=440.0)
    ref_x = src.render(0, n).data[:, 0]
    assert np.array_equal(ref_x, wl.c1_sine(n)), "workloads.c1_sine != reference SinePE"
    pe = pg.ConvolvePE(pg.SinePE(frequency=440.0), pg.ArrayPE(wl.c1_fir()))
    r = pg.NullRenderer(sample_rate=wl.SR_441)
    r.set_source(pe)
    r.start()
    y = _pull_all(pe, (n,))
    r.stop()
    assert pe.fft_size == 4096
    np.savez_compressed(os.path.join(GOLD, "c1_sine_fir4096.npz"), y=y, n=np.int64(n))


def gen_c2(n_pulls=24):
    pg.set_sample_rate(wl.SR_48)
    n = n_pulls * wl.C2_PULL
    pe = pg.ConvolvePE(pg.ArrayPE(wl.c2_input(n)), pg.ArrayPE(wl.c2_ir()))
    y = _pull_all(pe, (wl.C2_PULL,) * n_pulls)
    np.savez_compressed(os.path.join(GOLD, "c2_stereo_reverb.npz"), y=y, n_pulls=np.int64(n_pulls))


def gen_c3(table_f32, n_sources=8, n_pulls=10):
    pg.set_sample_rate(wl.SR_441)
    n = n_pulls * wl.C3_PULL
    el = wl.c3_elevations()[:n_sources]
    methods = [pg.SpatialHRTF(azimuth=wl.c3_azimuth(s, 0, n_pulls, n_sources), elevation=float(el[s]))
               for s in range(n_sources)]
    pes = [pg.SpatialPE(pg.ArrayPE(wl.c3_source(n, s, n_sources)), method=m) for s, m in enumerate(methods)]
    mix = pg.MixPE(*pes)
    outs = []
    per_source0 = []
    for b in range(n_pulls):
        for s, m in enumerate(methods):
            m.azimuth = wl.c3_azimuth(s, b, n_pulls, n_sources)  # attribute mutation between pulls
        outs.append(mix.render(b * wl.C3_PULL, wl.C3_PULL).data.copy())
    y = np.concatenate(outs, axis=0)
    # one source alone, ragged pulls, with a negative-azimuth (L/R swap) segment and a reset
    m = pg.SpatialHRTF(azimuth=-37.0, elevation=12.0)
    sp = pg.SpatialPE(pg.ArrayPE(wl.c3_source(4000, 3, 1)), method=m)
    pulls = (100, 1, 127, 128, 129, 700, 45)
    segs = []
    pos = 0
    for i, d in enumerate(pulls):
        if i == 3:
            m.azimuth = 100.0
        if i == 5:
            m.elevation = -35.0
        segs.append(sp.render(pos, d).data.copy())
        pos += d
    segs.append(sp.render(3000, 400).data.copy())  # non-contiguous -> tail dropped
    # stereo source mixdown path (mean over channels, spatial_pe.py:483)
    st = np.stack([wl.c3_source(1500, 1, 1), wl.c3_source(1500, 2, 1)], axis=1)
    sp2 = pg.SpatialPE(pg.ArrayPE(st), method=pg.SpatialHRTF(azimuth=60.0, elevation=-20.0))
    y_st = _pull_all(sp2, (512, 512, 476))
    np.savez_compressed(os.path.join(GOLD, "c3_hrtf_mix.npz"), y=y, n_sources=np.int64(n_sources),
                        n_pulls=np.int64(n_pulls), single=np.concatenate(segs, axis=0),
                        single_pulls=np.array(pulls), stereo_src=y_st)


def gen_c4(n_streams=4, n_pulls=8, L=wl.C4_L):
    pg.set_sample_rate(wl.SR_441)
    n = n_pulls * wl.C4_PULL
    pes = [pg.ConvolvePE(pg.ArrayPE(wl.c4_input(n, s)), pg.ArrayPE(wl.c4_ir(s, L))) for s in range(n_streams)]
    per = np.stack([_pull_all(p, (wl.C4_PULL,) * n_pulls)[:, 0] for p in pes])
    pes2 = [pg.ConvolvePE(pg.ArrayPE(wl.c4_input(n, s)), pg.ArrayPE(wl.c4_ir(s, L))) for s in range(n_streams)]
    mix = _pull_all(pg.MixPE(*pes2), (wl.C4_PULL,) * n_pulls)
    np.savez_compressed(os.path.join(GOLD, "c4_streams_mix.npz"), per_stream=per, mix=mix,
                        n_streams=np.int64(n_streams), n_pulls=np.int64(n_pulls), L=np.int64(L))


def gen_c5(n_voices=16, n_pulls=24, L=wl.C5_L):
    pg.set_sample_rate(wl.SR_441)
    n = n_pulls * wl.C5_PULL
    v = wl.c5_voices(n, n_voices)
    voice_mix = pg.MixPE(*[pg.ArrayPE(v[i]) for i in range(n_voices)])
    pe = pg.ConvolvePE(voice_mix, pg.ArrayPE(wl.c5_ir(L)))
    y = _pull_all(pe, (wl.C5_PULL,) * n_pulls)
    vm = pg.MixPE(*[pg.ArrayPE(v[i]) for i in range(n_voices)]).render(0, n).data
    np.savez_compressed(os.path.join(GOLD, "c5_voicebank_longir.npz"), y=y, voice_mix=vm,
                        n_voices=np.int64(n_voices), n_pulls=np.int64(n_pulls), L=np.int64(L))


def gen_mix():
    pg.set_sample_rate(44_100)
    rng = np.random.default_rng(21)
    a = [rng.uniform(-1, 1, (300, 2)).astype(np.float32) for _ in range(7)]
    # inputs with different extents: only intersecting ones are summed (mix_pe.py:81-85)
    pes = [pg.ArrayPE(a[0]), pg.DelayPE(pg.ArrayPE(a[1]), 100), pg.ArrayPE(a[2][:50]),
           pg.DelayPE(pg.ArrayPE(a[3]), 250)] + [pg.ArrayPE(t) for t in a[4:]]
    y = pg.MixPE(*pes).render(0, 600).data
    y2 = pg.MixPE(*pes).render(560, 100).data  # only the delayed ones intersect... or none
    np.savez_compressed(os.path.join(GOLD, "mix_extents.npz"), y=y, y_late=y2, **{f"a{i}": t for i, t in enumerate(a)})


def main():
    os.makedirs(GOLD, exist_ok=True)
    table = gen_kemar()
    gen_unit_vectors()
    gen_ragged()
    gen_mix()
    gen_c1()
    gen_c2()
    gen_c3(table)
    gen_c4()
    gen_c5()
    tot = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print(f"golden written: {sorted(os.listdir(GOLD))}  total {tot/1e6:.2f} MB")


if __name__ == "__main__":
    main()
