#!/usr/bin/env python
"""
bench.py -- ConvolvePE hot-path throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--variant shared|distinct]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm on the host CPU cores

Workload (BASELINE.json configs[1]): stereo convolution reverb, 3 s IR
(132 300 taps) @ 48 kHz, 512-sample blocks, uniform partitions; 256 independent stereo
streams per GPU (weak scaling: streams are sharded, there is no data-path collective).
A *step* is one 512-sample pull of every stream: K1 (ingest + R2C) -> K3 (multiply-accumulate
over the 259 partitions) -> K2 (C2R + emit).

  value     audio-seconds x channels per second, inputs resident in HBM (CUDA events, max over ranks)
  e2e       same metric through the public host API (ConvolveBank.process: pinned host x -> H2D ->
            step -> D2H -> pinned host y, every step inside the timed region)
  roofline  dominant kernel (k_fdl_mac): algorithmic bytes / mean launch duration vs measured HBM peak
  cpu_baseline  the oracle port of the reference's numpy algorithm timed on this box's host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from pygmu2_b200 import workloads as wl  # noqa: E402

METRIC = "ConvolvePE audio-sec x channels/sec"
UNIT = "audio-s*ch/s"
SR = wl.SR_48
L, B, PULL = wl.C2_L, 512, wl.C2_PULL
STREAMS_PER_GPU = 256
CH = 2
N_INPUT_BLOCKS = 8  # distinct resident input blocks cycled through


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled via NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def finish(self):
        self._halt.set()
        self.join(timeout=2.0)
        med = int(statistics.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pygmu2_oracle as orc  # timed CPU baseline / checker only
    return orc


def _cpu_worker(args):
    """One host core: one stereo stream through the reference algorithm (float64 numpy overlap-save,
    one 262144-point rfft/irfft pair per 512-sample pull), `pulls` pulls.  Returns seconds."""
    stream, pulls, warm = args
    orc = _oracle()
    conv = orc.OracleConvolve(wl.c2_ir(), CH)                   # filter preparation excluded, as for the GPU
    x = wl.c2_input((pulls + warm) * PULL, stream)
    for p in range(warm):
        conv.render(x[p * PULL:(p + 1) * PULL])
    t0 = time.perf_counter()
    for p in range(warm, warm + pulls):
        conv.render(x[p * PULL:(p + 1) * PULL])
    return time.perf_counter() - t0


def cpu_baseline_single(seconds_budget: float = 12.0):
    """Bounded sample on 1 core: one stereo stream, ~budget seconds of pulls."""
    t_probe = _cpu_worker((0, 4, 1)) / 4
    pulls = int(max(8, min(600, seconds_budget / max(t_probe, 1e-4))))
    t = _cpu_worker((0, pulls, 1))
    audio = pulls * PULL / SR * CH
    return {"value": audio / t, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"1 stereo stream x {pulls} pulls of {PULL} @48k, L={L} (oracle port of the numpy "
                      f"overlap-save, nfft=262144), {t:.1f} s on 1 host core; filter prep excluded"}


_REF = {}


def _ref_init():
    """Per worker process: one stereo stream with its filter prepared once (untimed, like the GPU arm)."""
    orc = _oracle()
    _REF["conv"] = orc.OracleConvolve(wl.c2_ir(), CH)
    _REF["x"] = wl.c2_input(64 * PULL, os.getpid() % 1000)
    _REF["pos"] = 0
    _REF["conv"].render(_REF["x"][:PULL])


def _ref_step(pulls):
    conv, x = _REF["conv"], _REF["x"]
    t0 = time.perf_counter()
    for _ in range(pulls):
        p = _REF["pos"] % 64
        conv.render(x[p * PULL:(p + 1) * PULL])
        _REF["pos"] += 1
    return time.perf_counter() - t0


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port: float64 numpy overlap-save, one
    262144-point rfft/irfft pair per 512-sample pull and channel pair) on all host cores, one independent
    stereo stream per worker process.  A step = every worker renders `pulls_per_step` pulls."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    t_probe = _cpu_worker((0, 3, 1)) / 3
    budget = 150.0
    pulls_per_step = int(max(1, min(16, budget / max((K + W) * t_probe * 2.0, 1e-6))))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_ref_init) as pool:
        for _ in range(max(W, 1)):
            pool.map(_ref_step, [pulls_per_step] * cores, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(K):
            pool.map(_ref_step, [pulls_per_step] * cores, chunksize=1)
        dt = time.perf_counter() - t0
    audio_per_step = cores * CH * pulls_per_step * PULL / SR
    value = audio_per_step * K / dt
    sample = (f"{cores} worker processes x 1 stereo stream x {pulls_per_step} pulls of {PULL} per step "
              f"(oracle port of the reference: float64 numpy rfft/irfft nfft=262144 per pull), "
              f"{K} steps in {dt:.1f} s wall; filter prep excluded")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(args, cores_note=f"{cores} host cores"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def _config(args, cores_note=None):
    P = -(-L // B)
    c = {
        "workload": "C2 stereo convolution reverb: 3 s IR (132300 taps) @48 kHz, 512-sample blocks, uniform partitions",
        "streams_per_gpu": args.streams, "channels": CH, "filter": args.variant, "filter_len": L, "block": B,
        "partitions": P, "pull": PULL, "sample_rate": SR,
        "l2": "inputs larger than L2: the delay-line state streamed every step is "
              f"{args.streams * CH * P * B * 8 / 1e6:.0f} MB (+ filter spectra) vs 126 MB L2",
    }
    if cores_note:
        c["host"] = cores_note
    return c


# ---------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import pygmu2_b200 as pg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg.set_sample_rate(SR)

    N, K, W = args.streams, args.steps, args.warmup
    distinct = args.variant == "distinct"
    P = -(-L // B)
    if distinct:
        irs = np.stack([wl.c2_ir(stream=rank * N + s) for s in range(N)])
        bank = pg.ConvolveBank(irs, N, CH, block=B, max_pull=PULL, device=local)
    else:
        bank = pg.ConvolveBank(wl.c2_ir(), N, CH, block=B, max_pull=PULL, device=local, single_filter_dims=True)

    # synthetic inputs, resident in HBM: N_INPUT_BLOCKS pulls of uniform(-1,1), planar [blk][N][CH][PULL]
    rng = np.random.default_rng(1000 + rank)
    x_host = rng.uniform(-1.0, 1.0, (N_INPUT_BLOCKS, N, CH, PULL)).astype(np.float32)
    x_dev = torch.from_numpy(x_host).to(dev)
    y_dev = torch.empty((N, CH, PULL), dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    sh = stream.cuda_stream
    blk_bytes = N * CH * PULL * 4

    def step(i):
        bank.process_device(x_dev.data_ptr() + (i % N_INPUT_BLOCKS) * blk_bytes, y_dev.data_ptr(), PULL,
                            cuda_stream=sh)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # fill the delay line so the timed steps stream real (non-zero) spectra: P pulls, then W warm-up steps
    for i in range(P):
        step(i)
    for i in range(max(W, 3)):
        step(i)
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    l0 = bank.info().kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(K):
        step(i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    info = bank.info()
    launches = int(info.kernel_launches - l0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # instrumented pass of the same loop: CUDA events around each kernel on the launching stream
    bank.profile_begin()
    kp = min(K, 500)
    for i in range(kp):
        step(i)
    prof = bank.profile_end()
    clocks = sampler.finish()

    # e2e: public host API, pinned host buffers, H2D + step + D2H every step
    from pygmu2_b200._lib import PinnedArray
    xp = PinnedArray((N_INPUT_BLOCKS, N, CH, PULL))
    yp = PinnedArray((N, CH, PULL))
    xp.array[...] = x_host
    ke = min(K, 400)
    for i in range(3):
        bank.process(xp.array[i % N_INPUT_BLOCKS], out=yp.array)
    barrier()
    t0 = time.perf_counter()
    for i in range(ke):
        bank.process(xp.array[i % N_INPUT_BLOCKS], out=yp.array)
    torch.cuda.synchronize(dev)
    dt_e2e = time.perf_counter() - t0
    te = torch.tensor([dt_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    dt_e2e = float(te.item())
    checksum = float(np.abs(yp.array).mean())

    audio_per_step = world * N * CH * PULL / SR            # audio-seconds x channels, all ranks
    value = audio_per_step * K / (ms_max * 1e-3)
    e2e_value = audio_per_step * ke / dt_e2e

    peak, peak_src = measured_hbm_peak()
    Kbins = B + 1
    mac_bytes = N * CH * P * Kbins * 8 * (2 if distinct else 1) + N * CH * Kbins * 8   # rows streamed + Y written
    step_bytes = wl.bytes_per_block_step(N, CH, CH, L, B, distinct)
    mac_ms = prof.ms_mac / max(prof.steps, 1)
    step_ms_prof = (prof.ms_r2c + prof.ms_mac + prof.ms_c2r) / max(prof.steps, 1)
    achieved = mac_bytes / (mac_ms * 1e-3) / 1e9
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "mac_traffic.json")
    if os.path.exists(tr_path):
        try:
            traffic = json.load(open(tr_path)).get(args.variant)
        except Exception:
            traffic = None

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": _config(args),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": blk_bytes, "d2h_bytes_per_step": blk_bytes,
                    "steps": ke, "api": "ConvolveBank.process (pgx_bank_process, pinned host buffers)",
                    "checksum_mean_abs_y": checksum},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "k_fdl_mac", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": mac_bytes, "mean_launch_ms": mac_ms,
                         "launch_plan": {"grid": info.mac_grid, "term_splits": info.mac_split,
                                         "streams_per_cta": info.mac_stream_tile, "ctas_per_sm": info.mac_occupancy},
                         "share_of_step": prof.ms_mac / max(prof.ms_r2c + prof.ms_mac + prof.ms_c2r, 1e-12),
                         "timing": f"CUDA events around each kernel on the launching stream, {prof.steps} steps "
                                   "(instrumented pass of the same loop)",
                         "step": {"algorithmic_bytes": step_bytes, "ms": ms_max / K,
                                  "achieved": step_bytes / (ms_max / K * 1e-3) / 1e9,
                                  "frac": step_bytes / (ms_max / K * 1e-3) / 1e9 / peak,
                                  "kernel_ms": {"k_r2c_ingest": prof.ms_r2c / max(prof.steps, 1), "k_fdl_mac": mac_ms,
                                                "k_c2r_emit": prof.ms_c2r / max(prof.steps, 1),
                                                "sum": step_ms_prof}}},
        }
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline_single(args.cpu_seconds)
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    xp.free()
    yp.free()
    bank.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="shared", choices=["shared", "distinct"],
                    help="one IR shared by all streams (the named reverb) or one IR per stream")
    ap.add_argument("--streams", type=int, default=STREAMS_PER_GPU, help="stereo streams per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
