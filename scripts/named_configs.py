#!/usr/bin/env python
"""The five BASELINE.json configurations exactly as named (single graphs through the PE API, NullRenderer pull
loop), wall-clock on one B200: what a pygmu2 user who switches the import sees.  Prints one JSON line per config.

    python scripts/named_configs.py [--seconds S]     (S = audio seconds rendered per config, default 2)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pygmu2_b200 as pg  # noqa: E402
from pygmu2_b200 import workloads as wl  # noqa: E402


def run(name, pe, sr, pull, seconds, channels, before_pull=None):
    n_pulls = int(seconds * sr / pull)
    with pg.NullRenderer(sample_rate=sr) as r:
        r.set_source(pe)
        r.start()
        for p in range(4):                      # warm-up: builds banks, filter spectra, oscillator handles
            r.render(p * pull, pull)
        t0 = time.perf_counter()
        for p in range(4, 4 + n_pulls):
            if before_pull is not None:
                before_pull(p)
            r.render(p * pull, pull)
        dt = time.perf_counter() - t0
    audio = n_pulls * pull / sr
    print(json.dumps({"config": name, "pull": pull, "pulls": n_pulls, "audio_s": audio, "wall_s": dt,
                      "ms_per_pull": 1e3 * dt / n_pulls, "x_realtime": audio / dt,
                      "audio_s_ch_per_s": audio * channels / dt}), flush=True)


def run_paced(name, pe, sr, pull, seconds):
    """The same pull loop PACED at real time (what AudioRenderer's callback does): the latency of each render() call
    when the caller is idle in between, not the throughput of a back-to-back loop."""
    n_pulls = int(seconds * sr / pull)
    period = pull / sr
    lat = []
    with pg.NullRenderer(sample_rate=sr) as r:
        r.set_source(pe)
        r.start()
        for p in range(4):
            r.render(p * pull, pull)
        nxt = time.perf_counter()
        for p in range(4, 4 + n_pulls):
            nxt += period
            while time.perf_counter() < nxt:
                pass
            t0 = time.perf_counter()
            r.render(p * pull, pull)
            lat.append(time.perf_counter() - t0)
    lat = np.array(lat) * 1e6
    print(json.dumps({"config": name + " [paced at real time]", "pull": pull, "pulls": n_pulls,
                      "latency_us_median": float(np.median(lat)), "latency_us_p99": float(np.percentile(lat, 99)),
                      "ms_per_pull": float(np.median(lat)) / 1e3, "x_realtime": period / (float(np.median(lat)) * 1e-6)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--paced", action="store_true", help="only the paced C5 / C2 latency measurement")
    a = ap.parse_args()
    if a.paced:
        pg.set_sample_rate(wl.SR_441)
        voices = [pg.SuperSawPE(frequency=55.0 * 2.0 ** (i / 128.0), amplitude=1.0 / 32.0, seed=i) for i in range(wl.C5_VOICES)]
        run_paced("C5 1024 SuperSaw -> MixPE -> 441000-tap IR", pg.ConvolvePE(pg.MixPE(*voices), pg.ArrayPE(wl.c5_ir()),
                                                                             block_size=64), wl.SR_441, 64, min(a.seconds, 1.0))
        pg.set_sample_rate(wl.SR_48)
        x = wl.c2_input(int((a.seconds + 1) * wl.SR_48))
        run_paced("C2 stereo reverb 132300 taps", pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(wl.c2_ir())), wl.SR_48, 512, a.seconds)
        return
    S = a.seconds
    # C1: SinePE 440 Hz -> ConvolvePE with a 4096-tap FIR, mono, 44.1 kHz, default fft_size
    pg.set_sample_rate(wl.SR_441)
    run("C1 SinePE->ConvolvePE 4096-tap", pg.ConvolvePE(pg.SinePE(440.0), pg.ArrayPE(wl.c1_fir())), wl.SR_441, 4096, S, 1)
    # C2: stereo convolution reverb, 3 s IR @48 kHz, 512-sample pulls
    pg.set_sample_rate(wl.SR_48)
    x = wl.c2_input(int((S + 1) * wl.SR_48))
    run("C2 stereo reverb 132300 taps", pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(wl.c2_ir())), wl.SR_48, 512, S, 2)
    run("C2 as ReverbPE (fused wet/dry)", pg.ReverbPE(pg.ArrayPE(x), pg.ArrayPE(wl.c2_ir()), mix=0.3), wl.SR_48, 512, S, 2)
    # C3: 256 SpatialPE(HRTF) sources -> MixPE, KEMAR table, 512-sample pulls (azimuths fixed here)
    pg.set_sample_rate(wl.SR_441)
    el = wl.c3_elevations()
    n3 = int((S + 1) * wl.SR_441)
    srcs = [pg.SpatialPE(pg.ArrayPE(wl.c3_source(n3, i)), method=pg.SpatialHRTF(wl.c3_azimuth(i, 0, 2), el[i]))
            for i in range(wl.C3_SOURCES)]
    run("C3 256 HRTF sources -> MixPE (static)", pg.MixPE(*srcs), wl.SR_441, 512, S, 2)
    srcs = [pg.SpatialPE(pg.ArrayPE(wl.c3_source(n3, i)), method=pg.SpatialHRTF(wl.c3_azimuth(i, 0, 2), el[i]))
            for i in range(wl.C3_SOURCES)]
    n_total = int(S * wl.SR_441 / 512) + 4
    az_tab = np.array([[wl.c3_azimuth(i, p, n_total) for i in range(wl.C3_SOURCES)] for p in range(n_total + 1)])

    def move(p):                                # the user's own per-pull code: every source gets a new azimuth
        for i, sp in enumerate(srcs):
            sp.method.azimuth = az_tab[p, i]

    run("C3 256 MOVING HRTF sources -> MixPE", pg.MixPE(*srcs), wl.SR_441, 512, S, 2, before_pull=move)
    # the same motion handed over once as a table (extension): no per-pull host work at all
    srcs = [pg.SpatialPE(pg.ArrayPE(wl.c3_source(n3, i)), method=pg.SpatialHRTF(wl.c3_azimuth(i, 0, 2), el[i]))
            for i in range(wl.C3_SOURCES)]
    mix = pg.MixPE(*srcs)
    mix.set_trajectory(az_tab, el, hop=512)
    run("C3 256 MOVING HRTF sources -> MixPE, trajectory table", mix, wl.SR_441, 512, S, 2)
    # C5: 1024 SuperSawPE voices -> MixPE -> 10 s IR at 64-sample pulls
    voices = [pg.SuperSawPE(frequency=55.0 * 2.0 ** (i / 128.0), amplitude=1.0 / 32.0, seed=i) for i in range(wl.C5_VOICES)]
    run("C5 1024 SuperSaw -> MixPE -> 441000-tap IR", pg.ConvolvePE(pg.MixPE(*voices), pg.ArrayPE(wl.c5_ir()), block_size=64),
        wl.SR_441, 64, min(S, 1.0), 1)
    voices = [pg.SuperSawPE(frequency=55.0 * 2.0 ** (i / 128.0), amplitude=1.0 / 32.0, seed=i) for i in range(wl.C5_VOICES)]
    run("C5 same, two-level partitions (64 / 4096)",
        pg.ConvolvePE(pg.MixPE(*voices), pg.ArrayPE(wl.c5_ir()), block_size=64, tail_block=4096), wl.SR_441, 64, min(S, 1.0), 1)


if __name__ == "__main__":
    main()
