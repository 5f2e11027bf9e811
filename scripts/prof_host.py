#!/usr/bin/env python
"""Where the host time of a single-graph pull goes: cProfile of the C2 and C3 named graphs through the PE API."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pygmu2_b200 as pg  # noqa: E402
from pygmu2_b200 import workloads as wl  # noqa: E402


def loop(r, n, pull, start=4):
    for p in range(start, start + n):
        r.render(p * pull, pull)


def run(name, pe, sr, pull, n=3000):
    with pg.NullRenderer(sample_rate=sr) as r:
        r.set_source(pe)
        r.start()
        loop(r, 4, pull, 0)
        t0 = time.perf_counter()
        loop(r, n, pull)
        dt = time.perf_counter() - t0
        print(f"== {name}: {1e6 * dt / n:.1f} us per pull (no profiler)")
        pr = cProfile.Profile()
        pr.enable()
        loop(r, n, pull, 4 + n)
        pr.disable()
        st = pstats.Stats(pr)
        st.sort_stats("tottime").print_stats(14)


pg.set_sample_rate(wl.SR_48)
x = wl.c2_input(int(40 * wl.SR_48))
run("C2", pg.ConvolvePE(pg.ArrayPE(x), pg.ArrayPE(wl.c2_ir())), wl.SR_48, 512)
pg.set_sample_rate(wl.SR_441)
el = wl.c3_elevations()
n3 = int(40 * wl.SR_441)
srcs = [pg.SpatialPE(pg.ArrayPE(wl.c3_source(n3, i)), method=pg.SpatialHRTF(wl.c3_azimuth(i, 0, 2), el[i]))
        for i in range(wl.C3_SOURCES)]
run("C3 static", pg.MixPE(*srcs), wl.SR_441, 512)
