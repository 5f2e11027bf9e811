#!/usr/bin/env python
"""Soak test of the multi-stream engine (compute-sanitizer is not available on this pool): many random operation
sequences -- conv / mix pulls of random sizes, resets, filter-map changes, uniform and two-level banks, all
pipelined three deep -- each checked against a direct float64 model.  Prints one line per configuration.

    python scripts/soak.py [--seeds 24] [--steps 160]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pygmu2_b200 as pg  # noqa: E402
from pygmu2_b200._lib import PinnedArray  # noqa: E402

TOL = 1e-5


def rel_err(y, ref):
    """Max-abs error over FULL SCALE: inputs are uniform(-1, 1) and filters have unit energy, so outputs are O(1);
    a pull of one or two samples can have a tiny maximum by chance, which is not the scale of the signal."""
    return float(np.max(np.abs(np.asarray(y, np.float64) - ref)) / max(np.max(np.abs(ref)), 0.25))


def one(seed, L, B, TB, steps):
    rng = np.random.default_rng(seed)
    N, F, C = int(rng.integers(1, 6)), 4, int(rng.choice([1, 2]))
    h = (rng.standard_normal((F, L, C)) / np.sqrt(L)).astype(np.float32)
    fmap = rng.integers(0, F, N).astype(np.int32)
    bank = pg.ConvolveBank(h, N, 1, block=B, max_pull=4 * max(B, TB or B), filter_of_stream=fmap, tail_block=TB)
    two = bank.info().tail_block > 0
    hist = [np.zeros(0, np.float32) for _ in range(N)]
    pending, keep, worst, mode = [], [], 0.0, None
    fill = [0]   # samples in the open block: whole-block pulls on a block boundary are what runs as a graph replay

    def model(n, mix):
        per = np.zeros((N, C, n))
        for s in range(N):
            seg = hist[s][-(n + L - 1):].astype(np.float64)
            for c in range(C):
                full = np.convolve(seg, h[fmap[s], :, c].astype(np.float64))
                per[s, c] = full[seg.shape[0] - n:seg.shape[0]]
        return per.sum(axis=0) if mix else per

    def drain():
        nonlocal worst
        for tk, yp, ref, label in pending:
            bank.wait(tk)
            e = rel_err(yp.array, ref)
            worst = max(worst, e)
            assert e <= TOL, f"{label}: {e:.3e}"
        pending.clear()

    def pull(op, n, step):
        top = 4 * max(B, TB or B)
        x = rng.uniform(-1, 1, (N, 1, n)).astype(np.float32)
        for s in range(N):
            hist[s] = np.concatenate([hist[s], x[s, 0]])[-(L + top):]
        xp = PinnedArray((N, 1, n))
        xp.array[...] = x
        yp = PinnedArray((C, n) if op == "mix" else (N, C, n))
        tk = bank.submit(xp.array, yp.array, mix=(op == "mix"))
        pending.append((tk, yp, model(n, op == "mix"), f"seed {seed} L={L} B={B} TB={TB} step {step} {op} n={n}"))
        keep.extend([xp, yp])
        fill[0] = (fill[0] + n) % B
        if len(pending) >= 3:
            drain()

    ops = ["pull"] * 5 + ["mix"] * 3 + ["burst"] * 3 + ["reset_all", "reset_some", "map"]
    for step in range(steps):
        op = str(rng.choice(ops))
        if op == "burst":   # align to a block boundary, then a run of whole-block pulls (graph replays), conv or mix
            kind = mode if (two and mode) else str(rng.choice(["pull", "mix"]))
            if two and mode is None:
                mode = kind
            if fill[0]:
                pull(kind, B - fill[0], step)
            for _ in range(int(rng.integers(2, 7))):
                pull(kind, B, step)
            continue
        if two:  # two-level banks keep their map and their mode between resets
            if op == "map":
                op = "pull"
            if op in ("pull", "mix"):
                if mode is None:
                    mode = op
                op = mode
            if op == "reset_some" and mode == "mix":
                op = "reset_all"
        if op in ("pull", "mix"):
            top = 4 * max(B, TB or B)
            n = int(rng.choice([1, 7, B - 1, B, B + 1, 2 * B, 3 * B + 5, int(rng.integers(1, top))]))
            pull(op, n, step)
        else:
            drain()
            if op == "reset_all":
                bank.reset()
                hist = [np.zeros(0, np.float32) for _ in range(N)]
                mode = None
                fill[0] = 0
            elif op == "reset_some":
                ids = [int(s) for s in range(N) if rng.random() < 0.5] or [0]
                bank.reset(ids)
                for s in ids:
                    hist[s] = np.zeros(0, np.float32)
            else:
                fmap = rng.integers(0, F, N).astype(np.int32)
                bank.set_filter_map(fmap)
    drain()
    gp = int(bank.info().graph_pulls)
    for a in keep:
        a.free()
    bank.close()
    return worst, gp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=24)
    ap.add_argument("--steps", type=int, default=160)
    a = ap.parse_args()
    shapes = [(40, 64, None), (300, 64, None), (1000, 128, None), (70, 32, None), (5000, 256, None), (513, 512, None),
              (3000, 64, 512), (700, 32, 128), (5000, 128, 1024), (9000, 512, 2048), (20000, 16, 256), (2000, 1024, None)]
    t0 = time.time()
    total = 0
    for L, B, TB in shapes:
        worst, graphs = 0.0, 0
        for seed in range(a.seeds):
            w, gp = one(1000 * B + seed, L, B, TB, a.steps)
            worst, graphs = max(worst, w), graphs + gp
            total += 1
        print(f"L={L:6d} B={B:5d} tail={TB}: {a.seeds} sequences x {a.steps} operations ok, worst error {worst:.2e}, "
              f"{graphs} pulls ran as graph replays", flush=True)
    print(f"soak ok: {total} sequences in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
