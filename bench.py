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

  value     audio-seconds x channels per second, inputs resident in HBM (CUDA events, max over ranks); the
            K-step loop is repeated (>= 50 times when short) and the MEDIAN repetition is reported, spread in "reps"
  e2e       same metric through the public host API (ConvolveBank.submit/wait: pinned host x -> H2D ->
            step -> D2H -> pinned host y, every step inside the timed region)
  roofline  dominant kernel (k_fdl_mac): algorithmic bytes / time per launch over the UN-instrumented timed loop
            vs measured HBM peak; the per-kernel CUDA-event pass is reported beside it ("instrumented")
  parity    untimed leg on the SAME bank and launch plan: reset, P+2 random pulls, streams {0,1,N/2,N-1} (and
            the mix) against a float64 scipy.signal.fftconvolve of the same inputs; > 1e-5 exits non-zero
  cpu_baseline  the reference itself (oracle/_ref, else the oracle port) timed on this box's host cores
  c4        (--gpus N > 1) the sharded-mix configuration: 512 streams per GPU, fused mix, pgx_mix_reduce over
            NVLink peer memory onto rank 0 every pull -- value, e2e, reduce cost, parity of the reduced mix
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


def _reference_pkg():
    """The real reference package from oracle/_ref (built by oracle/build_ref.py where /root/reference exists and
    shipped with the snapshot), or None."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import build_ref
        return build_ref.import_ref()
    except Exception:
        return None


class _RefStream:
    """One stereo C2 stream on the CPU: the reference's own ConvolvePE pulled through its NullRenderer
    (kind "reference"), or -- when oracle/_ref is absent -- the oracle port of the same algorithm (kind "port").
    Either way: float64 numpy overlap-save, one 262144-point rfft/irfft pair per 512-sample pull."""

    def __init__(self, stream: int, n_pulls: int):
        self.x = wl.c2_input(n_pulls * PULL, stream)
        self.n_pulls, self.pos = n_pulls, 0
        ref = _reference_pkg()
        if ref is not None:
            self.kind = "reference"
            ref.set_sample_rate(SR)
            self.pe = ref.ConvolvePE(ref.ArrayPE(self.x), ref.ArrayPE(wl.c2_ir()))
            self.renderer = ref.NullRenderer(sample_rate=SR)
            self.renderer.set_source(self.pe)
            self.renderer.start()
        else:
            self.kind = "port"
            self.conv = _oracle().OracleConvolve(wl.c2_ir(), CH)

    def pull(self):
        p = self.pos % self.n_pulls
        if p == 0 and self.pos:      # wrap: a non-contiguous pull, history := 0 (convolve_pe.py:255-256) - same cost
            pass
        if self.kind == "reference":
            self.renderer.render(p * PULL, PULL)
        else:
            self.conv.render(self.x[p * PULL:(p + 1) * PULL])
        self.pos += 1


def _cpu_worker(args):
    """One host core: one stereo stream through the reference, `pulls` pulls after `warm` (filter preparation
    happens in the first warm-up pull and is excluded, as for the GPU).  Returns (seconds, kind)."""
    stream, pulls, warm = args
    st = _RefStream(stream, pulls + max(warm, 1))
    for _ in range(max(warm, 1)):
        st.pull()
    t0 = time.perf_counter()
    for _ in range(pulls):
        st.pull()
    return time.perf_counter() - t0, st.kind


def cpu_baseline_single(seconds_budget: float = 12.0):
    """Bounded sample on 1 core: one stereo stream, ~budget seconds of pulls."""
    t_probe = _cpu_worker((0, 4, 1))[0] / 4
    pulls = int(max(8, min(600, seconds_budget / max(t_probe, 1e-4))))
    t, kind = _cpu_worker((0, pulls, 1))
    audio = pulls * PULL / SR * CH
    what = ("the reference's own ConvolvePE through its NullRenderer (oracle/_ref)" if kind == "reference"
            else "oracle port of the reference (oracle/_ref absent)")
    return {"value": audio / t, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"1 stereo stream x {pulls} pulls of {PULL} @48k, L={L}: {what}, float64 numpy overlap-save "
                      f"with nfft=262144 per pull, {t:.1f} s on 1 host core; filter prep excluded"}


def cpu_baseline_c5v(n_voices: int, seconds_budget: float = 12.0):
    """C5 with its front end on ONE host core, bounded sample: the oracle port of SuperSawPE (7 BlitSaw
    oscillators, float64, scipy lfilter) for a few voices x pulls, plus the oracle ConvolvePE pull
    (one 524288-point float64 rfft/irfft pair per 64-sample pull); voice time is scaled to n_voices."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pygmu2_oracle_sources as osrc  # timed CPU baseline only
    orc = _oracle()
    sv, pulls = 8, 16
    voices = [osrc.OracleSuperSaw(55.0 * 2.0 ** (i / 128.0), 1.0 / 32.0, seed=i, sample_rate=wl.SR_441) for i in range(sv)]
    for v in voices:
        v.render(0, 64)
    t0 = time.perf_counter()
    for p in range(1, pulls + 1):
        for v in voices:
            v.render(p * 64, 64)
    t_voice = (time.perf_counter() - t0) / (sv * pulls)           # seconds per voice per 64-sample pull
    conv = orc.OracleConvolve(wl.c5_ir(), 1)
    x = np.random.default_rng(0).uniform(-1, 1, (64, 1)).astype(np.float32)
    conv.render(x)
    cp = int(max(3, min(40, (seconds_budget - 2.0) / 0.08)))
    t0 = time.perf_counter()
    for _ in range(cp):
        conv.render(x)
    t_conv = (time.perf_counter() - t0) / cp
    per_pull = t_voice * n_voices + t_conv
    return {"value": (64 / wl.SR_441) / per_pull, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{sv} SuperSaw voices x {pulls} pulls of 64 ({t_voice * 1e6:.0f} us per voice-pull, scaled to "
                      f"{n_voices} voices) + {cp} ConvolvePE pulls with the 441000-tap IR ({t_conv * 1e3:.1f} ms each, "
                      "nfft=524288), 1 host core, oracle port of the reference"}


_REF = {}


def _ref_init():
    """Per worker process: one stereo stream with its filter prepared once (untimed, like the GPU arm)."""
    _REF["st"] = _RefStream(os.getpid() % 1000, 64)
    _REF["st"].pull()


def _ref_step(pulls):
    st = _REF["st"]
    t0 = time.perf_counter()
    for _ in range(pulls):
        st.pull()
    return time.perf_counter() - t0, st.kind


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
    t_probe = _cpu_worker((0, 3, 1))[0] / 3
    budget = 150.0
    pulls_per_step = int(max(1, min(16, budget / max((K + W) * t_probe * 2.0, 1e-6))))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_ref_init) as pool:
        for _ in range(max(W, 1)):
            pool.map(_ref_step, [pulls_per_step] * cores, chunksize=1)
        t0 = time.perf_counter()
        kind = "port"
        for _ in range(K):
            kind = pool.map(_ref_step, [pulls_per_step] * cores, chunksize=1)[0][1]
        dt = time.perf_counter() - t0
    audio_per_step = cores * CH * pulls_per_step * PULL / SR
    value = audio_per_step * K / dt
    what = ("the UNMODIFIED reference (oracle/_ref): pygmu2.ConvolvePE pulled through pygmu2.NullRenderer"
            if kind == "reference" else "oracle port of the reference (oracle/_ref absent)")
    sample = (f"{cores} worker processes x 1 stereo stream x {pulls_per_step} pulls of {PULL} per step "
              f"({what}: float64 numpy rfft/irfft nfft=262144 per pull), "
              f"{K} steps in {dt:.1f} s wall; filter prep excluded")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(args, cores_note=f"{cores} host cores; this arm renders {cores} stereo streams at a time "
                                           f"(one per core), not {args.streams or STREAMS_PER_GPU}: streams are independent, "
                                           "the metric is per stream-channel"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def _config(args, cores_note=None):
    P = -(-L // B)
    c = {
        "workload": "C2 stereo convolution reverb: 3 s IR (132300 taps) @48 kHz, 512-sample blocks, uniform partitions",
        "streams_per_gpu": args.streams or STREAMS_PER_GPU, "channels": CH, "filter": args.variant, "filter_len": L,
        "block": B, "partitions": P, "pull": PULL, "sample_rate": SR,
        "l2": "inputs larger than L2: the delay-line state streamed every step is "
              f"{(args.streams or STREAMS_PER_GPU) * CH * P * B * 8 / 1e6:.0f} MB (+ filter spectra) vs 126 MB L2",
    }
    if cores_note:
        c["host"] = cores_note
    return c


# ---------------------------------------------------------------------------
# Workloads.  c2 is the headline (BASELINE.json configs[1]); the others are the remaining BASELINE
# configs at their SURVEY.md §8d roofline-run sizes, selectable with --workload for profiling.
def make_workload(args, w, rank, local):
    import pygmu2_b200 as pg
    distinct = args.variant == "distinct"
    if w == "c2":
        N = args.streams or STREAMS_PER_GPU
        if distinct:
            irs = np.stack([wl.c2_ir(stream=rank * N + s) for s in range(N)])
            bank = pg.ConvolveBank(irs, N, CH, block=B, max_pull=PULL * max(1, args.e2e_blocks), device=local,
                                   tail_block=args.tail_block or None)
        else:
            bank = pg.ConvolveBank(wl.c2_ir(), N, CH, block=B, max_pull=PULL * max(1, args.e2e_blocks), device=local,
                                   single_filter_dims=True,
                                   tail_block=args.tail_block or None)
        cfg = _config(args)
        if args.tail_block:  # NOT the named configuration (uniform partitions): an extra data point
            cfg["partitioning"] = (f"two-level: {bank.partitions} head partitions of {B} + {bank.tail_partitions} tail "
                                   f"partitions of {args.tail_block}; roofline bytes are still SURVEY's uniform figure")
        if args.reverb:  # ReverbPE's wet/dry tail fused into the inverse-FFT kernel (reverb_pe.py:82-95)
            bank.set_output_gains(0.3, 0.7)
            cfg["output_stage"] = "fused ReverbPE wet/dry: y = 0.7*x + 0.3*conv (pgx_bank_set_output_gains)"
        ir_of = (lambda r, s_: wl.c2_ir(stream=r * N + s_)) if distinct else (lambda r, s_: wl.c2_ir())
        gains = (0.3, 0.7) if args.reverb else None
        return dict(name="c2", gains=gains, bank=bank, N=N, c_in=CH, c_out=CH, L=L, B=B, pull=PULL, sr=SR, distinct=distinct, mix=False,
                    config=cfg, fill_steps=bank.partitions, ir_of=ir_of)
    if w == "c1":   # 4096 mono streams x distinct 4096-tap FIRs, B = 4096 (P = 1): FFT-stage bound
        N = args.streams or 4096
        rng = np.random.default_rng(1234 + rank)
        irs = (rng.standard_normal((N, 4096)) / 64.0).astype(np.float32)
        bank = pg.ConvolveBank(irs, N, 1, block=4096, max_pull=4096, device=local)
        cfg = {"workload": "C1 4096-tap FIR, mono, 44.1 kHz: N independent streams x distinct filters, B=4096, P=1",
               "streams_per_gpu": N, "block": 4096, "partitions": 1, "pull": 4096, "sample_rate": wl.SR_441,
               "l2": f"per-step footprint {N * (4097 * 8 * 3 + 4096 * 8) / 1e6:.0f} MB vs 126 MB L2"}
        return dict(name="c1", bank=bank, N=N, c_in=1, c_out=1, L=4096, B=4096, pull=4096, sr=wl.SR_441, distinct=True,
                    mix=False, config=cfg, fill_steps=2, ir_of=lambda r, s_: irs[s_][:, None])
    if w == "c3":   # 256 moving mono sources x 512-tap HRTF pairs -> one stereo mix (fused)
        N = args.streams or wl.C3_SOURCES
        table = wl.c3_synthetic_hrtf_table(512)
        both = np.concatenate([table, table[:, :, ::-1]], axis=0)
        bank = pg.ConvolveBank(both, N, 1, block=512, max_pull=512, device=local, mixdown_input=True,
                               filter_of_stream=np.zeros(N, np.int32))
        cfg = {"workload": "C3 SpatialPE HRTF: 256 moving sources x 512-tap per ear (synthetic table, 736 resident "
                           "filter pairs), fused MixPE stereo sum, 512-sample pulls @44.1 kHz, filter re-selected every pull",
               "sources_per_gpu": N, "block": 512, "partitions": 1, "pull": 512, "sample_rate": wl.SR_441,
               "l2": "L2-resident by nature (5.8 MB per step): launch/FFT-bound, reported as such"}
        return dict(name="c3", bank=bank, N=N, c_in=1, c_out=2, L=512, B=512, pull=512, sr=wl.SR_441, distinct=True, mix=True,
                    config=cfg, fill_steps=2, moving=True, n_filters=both.shape[0], filters=both)
    if w == "c4":   # 512 mono streams per GPU x distinct 2 s IRs, fused mix (+ one NCCL reduce per pull when sharded)
        N = args.streams or 512
        irs = np.stack([wl.c4_ir(rank * N + s) for s in range(N)])
        bank = pg.ConvolveBank(irs, N, 1, block=512, max_pull=512, device=local)
        cfg = {"workload": "C4 independent streams x 2 s random IRs (88200 taps) @44.1 kHz, 512-sample pulls, fused MixPE "
                           "sum, sharded across GPUs with one NCCL reduce of the mix per pull",
               "streams_per_gpu": N, "block": 512, "partitions": bank.partitions, "pull": 512, "sample_rate": wl.SR_441,
               "l2": f"delay line + filter spectra streamed per step: {2 * N * bank.partitions * 512 * 8 / 1e6:.0f} MB vs 126 MB L2"}
        return dict(name="c4", bank=bank, N=N, c_in=1, c_out=1, L=wl.C4_L, B=512, pull=512, sr=wl.SR_441, distinct=True,
                    mix=True, config=cfg, fill_steps=bank.partitions, reduce=True,
                    ir_of=lambda r, s_: irs[s_][:, None])
    if w == "c5":   # 10 s IR at 64-sample blocks: N=1 is the named latency case, N=256 replicas give an HBM figure
        N = args.streams or 1
        bank = pg.ConvolveBank(wl.c5_ir(), N, 1, block=64, max_pull=64, device=local, single_filter_dims=True,
                               tail_block=args.tail_block or None)
        cfg = {"workload": "C5 10 s IR (441000 taps) @44.1 kHz at 64-sample pulls (6891 uniform partitions), "
                           f"{N} stream(s); voice generation excluded",
               "streams_per_gpu": N, "block": 64, "partitions": bank.partitions, "pull": 64, "sample_rate": wl.SR_441,
               "l2": ("state 7 MB: L2-resident, latency-bound" if N == 1 else
                      f"delay line streamed per step: {N * bank.partitions * 64 * 8 / 1e6:.0f} MB vs 126 MB L2")}
        if args.tail_block:
            cfg["partitioning"] = (f"two-level: {bank.partitions} head partitions of 64 + {bank.tail_partitions} tail "
                                   f"partitions of {args.tail_block}; roofline bytes are still SURVEY's uniform figure")
        return dict(name="c5", bank=bank, N=N, c_in=1, c_out=1, L=wl.C5_L, B=64, pull=64, sr=wl.SR_441, distinct=False,
                    mix=False, config=cfg, fill_steps=400, ir_of=lambda r, s_: wl.c5_ir()[:, None])
    if w == "c5v":  # C5 with its front end: 1024 SuperSawPE voices -> MixPE -> 10 s IR at 64-sample pulls, all in HBM
        V = args.streams or wl.C5_VOICES
        pg.set_sample_rate(wl.SR_441)
        voices = [pg.SuperSawPE(frequency=55.0 * 2.0 ** (i / 128.0), amplitude=1.0 / 32.0, seed=i) for i in range(V)]
        mixpe = pg.MixPE(*voices)
        pe = pg.ConvolvePE(mixpe, pg.ArrayPE(wl.c5_ir()), block_size=64, tail_block=args.tail_block or None)
        pe.render(0, 64)                       # builds the VoiceBank (V x 7 oscillators) and the 6891-partition bank
        bank, vb = pe.bank, mixpe._fused.vb
        cfg = {"workload": f"C5 with its front end on the device: {V} SuperSawPE voices (7 BLIT oscillators each, "
                           "float64) -> MixPE (float32 voice sum) -> ConvolvePE with a 10 s IR (441000 taps) @44.1 kHz, "
                           "64-sample pulls; voice generation INCLUDED in every step",
               "voices": V, "oscillators": V * 7, "block": 64, "partitions": bank.partitions, "pull": 64,
               "sample_rate": wl.SR_441, "l2": "state 7 MB: L2-resident, latency-bound"}
        return dict(name="c5v", bank=bank, N=1, c_in=1, c_out=1, L=wl.C5_L, B=64, pull=64, sr=wl.SR_441, distinct=False,
                    mix=False, config=cfg, fill_steps=400, voicebank=vb, pe=pe)
    raise SystemExit(f"unknown workload {w}")


class _StdoutToStderr:
    """NCCL prints its version banner on stdout when the first communicator comes up; the contract is ONE JSON
    line on stdout, so file descriptor 1 points at stderr while the process group is being set up."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)
        return False


def _setup_dist():
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        from pygmu2_b200.dist import bind_host_to_gpu
        numa = bind_host_to_gpu(local)   # host staging buffers next to this rank's GPU (best effort; reported)
        torch.cuda.set_device(local)
        with _StdoutToStderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()   # brings the communicator up (and its banner out) now
    else:
        torch.cuda.set_device(local)
    return world, rank, local, numa


def _timed_reps(K, step, barrier, stream, reps):
    """`reps` repetitions of the K-step loop, each bracketed by barrier + synchronize; ms per repetition."""
    import torch
    out, host = [], []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        th0 = time.perf_counter()
        for i in range(K):
            step(i)
        host.append((time.perf_counter() - th0) * 1e3 / K)
        e1.record(stream)
        barrier()
        out.append(e0.elapsed_time(e1))
    return out, host


def _fftconv64(x, h, n):
    from scipy.signal import fftconvolve
    return fftconvolve(np.asarray(x, np.float64), np.asarray(h, np.float64))[:n]


def parity_leg(spec, bank, dev, stream, world, rank, do_reduce, traj_dev, traj):
    """Untimed, on the SAME bank and launch plan as the timed loop: reset, P+2 random pulls through the same
    device entry point, then streams {0, 1, N/2, N-1} (for a fused mix: the mix of those streams, the others
    fed zeros) against the float64 linear convolution of the same inputs.  Returns the "parity" object."""
    import torch
    import torch.distributed as dist

    N, c_in, c_out, pull, mix = spec["N"], spec["c_in"], spec["c_out"], spec["pull"], spec["mix"]
    P = bank.partitions
    n_p = -(-spec["L"] // pull) + 2 if P > 1 else 4     # every partition (of both levels of a two-level bank) live at the end
    n = n_p * pull
    sample = sorted({0, min(1, N - 1), N // 2, N - 1})
    bank.synchronize()
    bank.reset()
    gen = torch.Generator(device=dev)
    gen.manual_seed(4242 + rank)
    scale = 1.0 / N if spec["name"] in ("c3", "c4") else 1.0
    x_all = (torch.rand((n_p, N, c_in, pull), generator=gen, device=dev, dtype=torch.float32) * 2.0 - 1.0) * scale
    moving = traj_dev is not None
    if mix and not moving:   # only the sampled streams sound: the mix is then checkable against a handful of convolutions
        keep = torch.zeros(N, dtype=torch.bool, device=dev)
        keep[sample] = True
        x_all *= keep.view(1, N, 1, 1)
    n_out_ch = c_out if mix else N * c_out
    y_all = torch.empty((n_p, n_out_ch, pull), dtype=torch.float32, device=dev)
    torch.cuda.synchronize(dev)   # x_all was produced on torch's stream; the bank's streams are not ordered with it
    sh = stream.cuda_stream
    blk = N * c_in * pull * 4
    for i in range(n_p):
        if moving:
            bank.use_filter_map_device(traj_dev.data_ptr() + (i % 64) * N * 4)
        bank.process_device(x_all.data_ptr() + i * blk, y_all.data_ptr() + i * n_out_ch * pull * 4, pull, mix=mix,
                            cuda_stream=sh, input_resident=True, reduce=do_reduce)
    torch.cuda.synchronize(dev)
    if moving:
        bank.use_filter_map_device(None)
    y = y_all.cpu().numpy()
    worst, checked = 0.0, []
    if moving:      # C3: per pull, the pull's filter pair applied to [previous block | this block] (spatial_pe.py:499-511)
        x = x_all[:, :, 0, :].cpu().numpy().astype(np.float64)            # (n_p, N, pull), mono sources
        tab = spec["filters"].astype(np.float64)                          # (F, L, 2)
        Lf = tab.shape[1]
        ref = np.zeros((n_p, 2, pull))
        prev = np.zeros((N, pull))
        for i in range(n_p):
            win = np.concatenate([prev, x[i]], axis=1)                    # (N, 2*pull)
            H = np.fft.rfft(tab[traj[i % 64]], n=4 * pull, axis=1)        # (N, bins, 2)
            X = np.fft.rfft(win, n=4 * pull, axis=1)
            yy = np.fft.irfft(X[:, :, None] * H, n=4 * pull, axis=1)[:, pull:2 * pull, :]   # (N, pull, 2)
            ref[i] = yy.sum(axis=0).T
            prev = x[i]
        del Lf
        worst = float(np.max(np.abs(y - ref)) / np.max(np.abs(ref)))
        checked = [f"mix of all {N} sources"]
    elif mix:
        xs = x_all[:, sample].cpu().numpy()                               # (n_p, S, c_in, pull)
        ref = np.zeros((c_out, n))
        for k, s_ in enumerate(sample):
            h = spec["ir_of"](rank, s_)
            for c in range(c_out):
                ref[c] += _fftconv64(xs[:, k, 0].reshape(-1), h[:, c if h.shape[1] > 1 else 0], n)
        if do_reduce:   # the reduced mix on the root is the sum over every rank's sampled streams
            t = torch.from_numpy(ref).to(dev)
            dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)   # float64 checker data over the plumbing
            ref = t.cpu().numpy()
        if not do_reduce or rank == 0:
            ym = np.transpose(y, (1, 0, 2)).reshape(c_out, n)
            worst = float(np.max(np.abs(ym - ref)) / np.max(np.abs(ref)))
        checked = [f"mix of streams {sample}" + (f" of each of the {world} ranks, reduced onto rank 0" if do_reduce else "")]
    else:
        xs = x_all[:, sample].cpu().numpy()                               # (n_p, S, c_in, pull)
        for k, s_ in enumerate(sample):
            h = spec["ir_of"](rank, s_)                                   # (L, c_f)
            for c in range(c_out):
                xc = xs[:, k, c if c_in > 1 else 0].reshape(-1)
                ref = _fftconv64(xc, h[:, c if h.shape[1] > 1 else 0], n)
                if spec.get("gains"):   # fused ReverbPE output stage: y = dry * x + wet * conv
                    ref = spec["gains"][1] * xc.astype(np.float64) + spec["gains"][0] * ref
                yc = y[:, s_ * c_out + c].reshape(-1)
                worst = max(worst, float(np.max(np.abs(yc - ref)) / np.max(np.abs(ref))))
        checked = [f"streams {sample}, every channel"]
    t = torch.tensor([worst], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    worst = float(t.item())
    bank.reset()
    return {"max_rel_err": worst, "tol": 1e-5, "ok": bool(worst <= 1e-5), "pulls": n_p, "samples": n,
            "partitions_live_at_the_end": P, "checked": checked[0],
            "against": "float64 scipy.signal.fftconvolve / numpy.fft of the same inputs (max-abs error over full scale "
                       "of the reference output); same bank, same launch plan, same device entry point as the timed loop"}


def measure(args, wname, world, rank, local, numa_bound):
    """Build one workload on this rank's GPU and measure it: value (repeated K-step loop, device-resident
    inputs), roofline, per-kernel instrumented pass, e2e (host buffers), parity.  Returns the record (rank 0) ."""
    import torch
    import torch.distributed as dist

    import pygmu2_b200 as pg
    from pygmu2_b200 import dist as pd
    from pygmu2_b200._lib import PinnedArray

    dev = torch.device("cuda", local)
    K, W = args.steps, max(args.warmup, 3)
    spec = make_workload(args, wname, rank, local)
    bank, N, c_in, c_out = spec["bank"], spec["N"], spec["c_in"], spec["c_out"]
    Lw, Bw, pull, sr = spec["L"], spec["B"], spec["pull"], spec["sr"]
    mix, distinct = spec["mix"], spec["distinct"]
    pg.set_sample_rate(sr)
    P = bank.partitions
    n_out_ch = c_out if mix else N * c_out            # output channels produced per GPU per step
    do_reduce = bool(spec.get("reduce")) and world > 1
    comm = None
    if do_reduce:   # the cross-GPU MixPE sum: pgx_mix_reduce over NVLink peer memory onto rank 0
        comm = pd.MixComm(local, rank, world, root=0, max_floats=c_out * pull)
        bank.attach_comm(comm)
    use_nccl = do_reduce and args.reduce == "nccl"    # A/B baseline: the library collective on the same partials

    # synthetic inputs, resident in HBM: N_INPUT_BLOCKS pulls of uniform(-1,1), planar [blk][N][c_in][pull]
    rng = np.random.default_rng(1000 + rank)
    x_host = rng.uniform(-1.0, 1.0, (N_INPUT_BLOCKS, N, c_in, pull)).astype(np.float32)
    if wname in ("c3", "c4"):
        x_host /= np.float32(N)
    x_dev = torch.from_numpy(x_host).to(dev)
    if spec.get("voicebank") is not None:
        spec["pe"].render(64, 64 * 8)  # leave the build pull behind; positions below continue from a round number

    y_dev = torch.empty((n_out_ch, pull), dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(device=dev, priority=-1)  # the critical (output) stream outranks the background pass
    sh = stream.cuda_stream
    blk_bytes = 0 if spec.get("voicebank") is not None else N * c_in * pull * 4
    out_bytes = n_out_ch * pull * 4
    traj = traj_dev = None
    if spec.get("moving"):  # per-pull filter choice of every source, resident as one int32 row per pull
        traj = np.random.default_rng(6).integers(0, spec["n_filters"], (64, N)).astype(np.int32)
        traj_dev = torch.from_numpy(traj).to(dev)

    vb = spec.get("voicebank")
    vpos = [64 * 9]  # running sample position of the voice front end (contiguous pulls; 0..64 was the build pull)

    def step(i):
        if vb is not None:  # voices + voice sum rendered into HBM on the same stream, then the convolution pull
            blk = vb.device_block(vpos[0], pull, mix=True, cuda_stream=sh)
            vpos[0] += pull
            bank.process_device(blk.ptr, y_dev.data_ptr(), pull, cuda_stream=sh, input_resident=False)
            return
        if traj_dev is not None:
            bank.use_filter_map_device(traj_dev.data_ptr() + (i % 64) * N * 4)
        bank.process_device(x_dev.data_ptr() + (i % N_INPUT_BLOCKS) * blk_bytes, y_dev.data_ptr(), pull,
                            mix=mix, cuda_stream=sh, input_resident=True, reduce=do_reduce and not use_nccl)
        if use_nccl:
            with torch.cuda.stream(stream):
                dist.reduce(y_dev, dst=0, op=dist.ReduceOp.SUM)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # fill the delay line so the timed steps stream real (non-zero) spectra, then W warm-up steps
    if do_reduce:
        barrier()   # banks are built at different speeds: the ranks enter the exchange together
    for i in range(spec["fill_steps"]):
        step(i)
    for i in range(W):
        step(i)
    barrier()

    # ---- value: the K-step loop, repeated; every repetition bracketed by barrier + synchronize, max over ranks
    sampler = ClockSampler(local)
    sampler.start()
    probe, _ = _timed_reps(K, step, barrier, stream, 2)
    rep_ms = max(probe[-1], 1e-3)
    reps = args.reps or int(min(max(5, 4000.0 / rep_ms), 60))     # ~4 s of timed work, 5..60 repetitions
    l0 = bank.info().kernel_launches + (vb.bank.launches if vb is not None else 0)
    ms_reps, host_reps = _timed_reps(K, step, barrier, stream, reps)
    launches = int(bank.info().kernel_launches + (vb.bank.launches if vb is not None else 0) - l0) // reps
    t = torch.tensor(ms_reps, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_reps = [float(v) for v in t.cpu()]
    ms_med = float(np.median(ms_reps))
    clocks = sampler.finish()
    info = bank.info()
    tile = int(getattr(info, "mac_tile", 1) or 1)     # > 1: the conv pass is time-tiled (PGX_TILE)

    # ---- instrumented pass of the same loop: CUDA events around each kernel on its launching stream
    bank.profile_begin()
    kp = min(max(K, 50), 500)
    for i in range(kp):
        step(i)
    prof = bank.profile_end()
    barrier()

    # ---- cost of the cross-GPU sum: the same loop without it (every rank keeps its partial mix)
    reduce_info = None
    if do_reduce:
        def step_local(i):
            bank.process_device(x_dev.data_ptr() + (i % N_INPUT_BLOCKS) * blk_bytes, y_dev.data_ptr(), pull,
                                mix=mix, cuda_stream=sh, input_resident=True)
        for i in range(W):
            step_local(i)
        ms_nr, _ = _timed_reps(K, step_local, barrier, stream, max(reps // 2, 3))
        tn = torch.tensor(ms_nr, dtype=torch.float64, device=dev)
        dist.all_reduce(tn, op=dist.ReduceOp.MAX)
        ms_nr_med = float(np.median(tn.cpu().numpy()))
        reduce_info = {"impl": "NCCL dist.reduce on the pull's stream (baseline)" if use_nccl else
                               "pgx_mix_reduce: partials stored into rank 0's mailbox over NVLink peer memory, summed in "
                               "rank order by one gather kernel (csrc/pgx_comm.cu)",
                       "bytes_per_pull": c_out * pull * 4, "ms_per_step_with": ms_med / K,
                       "ms_per_step_without": ms_nr_med / K,
                       "exposed_us_per_pull": (ms_med - ms_nr_med) / K * 1e3}

    # ---- parity of exactly this bank / plan
    parity = parity_leg(spec, bank, dev, stream, world, rank, do_reduce and not use_nccl, traj_dev, traj) \
        if vb is None else None
    if comm is not None:
        comm.check()

    # ---- e2e: public host API, pinned host buffers, H2D + step + D2H every step
    io_dt = np.int16 if args.pcm16 else np.float32   # --pcm16: WAV staging, int16 over PCIe, converted in HBM
    xp = PinnedArray((N_INPUT_BLOCKS, N, c_in, pull), io_dt, write_combined=args.wc)
    yp = PinnedArray((c_out, pull) if mix else (N, c_out, pull), io_dt)
    xp.array[...] = np.clip(np.rint(x_host * 32768.0), -32768, 32767).astype(np.int16) if args.pcm16 else x_host
    # pulls in flight: H2D of pull i+1 and D2H of pull i-1 overlap the kernels of pull i; a time-tiled bank computes the
    # past sums of `tile` blocks per pass, so it needs that many more pulls queued to keep its passes back to back
    E2E_DEPTH = 3 if (tile <= 1 or mix) else min(int(getattr(info, 'submit_depth', 0)) or 8, tile + 3)
    yps = [yp] + [PinnedArray(yp.shape, io_dt) for _ in range(E2E_DEPTH - 1)]
    # steady state of the pull loop: enough pulls that the pipeline's fill and drain (one H2D + one D2H latency)
    # do not dominate a region of a few milliseconds
    ke = int(min(max(K, 400), 4000)) if vb is None else min(K, 400)

    def e2e_step(i):
        """One pull through the public host API.  Pipelined: submit pull i, then wait for pull i-(DEPTH-1)."""
        if vb is not None:  # the PE graph itself: ConvolvePE(MixPE(voices), ir).render -> host Snippet
            yp.array[0, 0, :] = spec["pe"].render(vpos[0], pull).data[:, 0]
            vpos[0] += pull
            return None
        if traj is not None:
            bank.set_filter_map(traj[i % 64])
        tk = bank.submit(xp.array[i % N_INPUT_BLOCKS], yps[i % E2E_DEPTH].array, mix=mix, reduce=do_reduce)
        if tk >= E2E_DEPTH - 1:
            bank.wait(tk - (E2E_DEPTH - 1))
        return tk

    def e2e_drain(tk):
        if tk is not None:
            for t_ in range(max(tk - (E2E_DEPTH - 2), 0), tk + 1):
                bank.wait(t_)

    tk = None
    for i in range(6):
        tk = e2e_step(i)
    e2e_drain(tk)
    e2e_reps = 1 if vb is not None else 5
    dts = []
    for _ in range(e2e_reps):
        barrier()
        t0 = time.perf_counter()
        for i in range(ke):
            tk = e2e_step(i)
        e2e_drain(tk)
        torch.cuda.synchronize(dev)
        dts.append(time.perf_counter() - t0)
    # ---- supplementary: the same host API with M blocks per submit (--e2e-blocks M): one H2D and one D2H of M MiB
    multi = None
    M = int(args.e2e_blocks or 1)
    if M > 1 and wname == "c2" and vb is None and traj is None and not args.pcm16 and not do_reduce:
        try:
            xm = PinnedArray((2, N, c_in, M * pull), np.float32)
            for b_ in range(2):
                xm.array[b_] = np.concatenate([x_host[(b_ * M + j) % N_INPUT_BLOCKS] for j in range(M)], axis=2)
            depth_m = 3
            yms = [PinnedArray((N, c_out, M * pull), np.float32) for _ in range(depth_m)]
            kem = max(ke // M, 60)

            def mstep(i):
                tkm = bank.submit(xm.array[i % 2], yms[i % depth_m].array)
                bank.wait(tkm - (depth_m - 1))
                return tkm

            tkm = None
            for i in range(4):
                tkm = mstep(i)
            for t_ in range(tkm - (depth_m - 2), tkm + 1):
                bank.wait(t_)
            mdts = []
            for _ in range(3):
                barrier()
                t0 = time.perf_counter()
                for i in range(kem):
                    tkm = mstep(i)
                for t_ in range(tkm - (depth_m - 2), tkm + 1):
                    bank.wait(t_)
                torch.cuda.synchronize(dev)
                mdts.append(time.perf_counter() - t0)
            tm = torch.tensor(mdts, dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dtm = float(np.median(tm.cpu().numpy()))
            multi = {"blocks_per_submit": M, "submits": kem, "submits_in_flight": depth_m,
                     "seconds": dtm, "bytes_per_submit_each_way": N * c_in * M * pull * 4,
                     "copy_gbs_per_direction": N * c_in * M * pull * 4 * kem / dtm / 1e9}
            xm.free()
            for a_ in yms:
                a_.free()
        except Exception as exc:  # a supplementary leg never takes the line down
            multi = {"blocks_per_submit": M, "error": repr(exc)}
    te = torch.tensor(dts, dtype=torch.float64, device=dev)
    per_rank = [te.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, te)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    dt_e2e = float(np.median(te.cpu().numpy()))
    per_rank_s = [float(np.median(v.cpu().numpy())) for v in per_rank]
    checksum = float(np.abs(yp.array.astype(np.float32)).mean()) / (32768.0 if args.pcm16 else 1.0)
    if comm is not None:
        comm.check()

    # audio-seconds x channels per step, SURVEY.md 8d: N_streams x C_out x duration / sample_rate -- the stream-channels
    # CONVOLVED, also when they are then summed into one mix (C3: 256 sources x 2 ears, C4: 512 streams per GPU); the
    # number of mixed output channels is reported beside it in config
    units = world * N * c_out * pull / sr
    value = units * K / (ms_med * 1e-3)
    e2e_value = units * ke / dt_e2e
    io_div = 2 if args.pcm16 else 1
    d2h_step = (out_bytes if (not do_reduce or rank == 0) else 0) // io_div

    peak, peak_src = measured_hbm_peak()
    Kbins = Bw + 1
    xrows = N * (1 if info.c_x == 1 else c_in)
    mac_bytes = xrows * P * Kbins * 8 + (N * c_out * P * Kbins * 8 if distinct else 0) + n_out_ch * Kbins * 8
    per_block_mac_bytes = mac_bytes
    if tile > 1 and not mix:
        # one tiled launch: every delay-line row once, every filter row once (the T rows a slot needs are a sliding
        # register window), T result sets written -- for T output blocks
        mac_bytes = xrows * P * Kbins * 8 + (N * c_out * P * Kbins * 8 if distinct else 0) + tile * info.mac_split * n_out_ch * Kbins * 8
    step_bytes = wl.bytes_per_block_step(N, c_in, c_out, Lw, Bw, distinct)
    nprof = max(prof.steps, 1)
    mac_ms_union = prof.ms_mac_union / max(prof.n_mac, 1)   # busy time per launch (union of overlapping launches)
    mac_ms_each = prof.ms_mac / max(prof.n_mac, 1)          # mean start-to-end of one launch
    ksum = prof.ms_r2c + prof.ms_mac + prof.ms_c2r + prof.ms_fold + prof.ms_now + prof.ms_conv1
    mac_per_step = prof.n_mac / nprof                        # k_fdl_mac launches per step (1 for P > 1 banks)
    traffic, traffic_src = None, None
    tr_path = os.path.join(ROOT, "profiles", "mac_traffic.json")
    if wname in ("c2", "c4") and os.path.exists(tr_path) and not args.streams:
        try:
            tj = json.load(open(tr_path))
            key = args.variant if wname == "c2" else wname
            if tile > 1 and not mix:
                key = f"{key}_tile{tile}"
            traffic = tj.get(key)
            traffic_src = None if traffic is None else ("NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one k_fdl_mac[_tile] launch "
                           f"from the committed `ncu --set full` capture ({tj.get('source', 'profiles/')})")
        except Exception:
            traffic = None
    rec = None
    if rank == 0:
        if mix:
            spec["config"]["mixed_output"] = (f"{c_out} channel(s) per GPU per pull" +
                                              (", summed over the GPUs onto rank 0 (pgx_mix_reduce, NVLink peer memory)"
                                               if do_reduce and not use_nccl else
                                               ", reduced over the GPUs with NCCL" if do_reduce else ""))
        # the dominant kernel: one k_fdl_mac launch per step, launches back to back on the background streams, so
        # in the un-instrumented loop time per launch = step time (everything else overlaps it or adds to it:
        # a lower bound on the kernel's own bandwidth)
        if mac_per_step > 0:
            kname, kbytes, klaunch_ms = "k_fdl_mac", mac_bytes, (ms_med / K) / mac_per_step
            if tile > 1 and not mix:
                kname = f"k_fdl_mac_tile (one pass per {tile} blocks)"
        elif prof.ms_conv1 > 0:
            # one fused launch IS the block step: SURVEY 8d's per-step figure (which still counts the delay-line row the
            # fused kernel never writes or re-reads: its own minimum traffic is lower, see "fused_kernel_min_bytes")
            kname = "k_conv1 / k_conv1_r16 (fused K1+K2)" if not mix else "k_mix1 (fused ingest+FFT+HRTF+mix)"
            kbytes = step_bytes
            klaunch_ms = ms_med / K
        else:
            kname, kbytes, klaunch_ms = "block step", step_bytes, ms_med / K
        achieved = kbytes / (klaunch_ms * 1e-3) / 1e9
        rec = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_med / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": spec["config"],
            "reps": {"n": reps, "ms_per_step_median": ms_med / K, "ms_per_step_min": min(ms_reps) / K,
                     "ms_per_step_max": max(ms_reps) / K,
                     "ms_per_step_p10_p90": [float(np.percentile(ms_reps, 10)) / K, float(np.percentile(ms_reps, 90)) / K],
                     "note": f"the {K}-step timed loop repeated {reps} times, each bracketed by barrier + synchronize, "
                             "max over ranks per repetition; value and ms_per_step are the MEDIAN repetition"},
            "clocks": clocks,
            "parity": parity,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": blk_bytes // io_div,
                    "d2h_bytes_per_step": d2h_step,
                    "steps": ke, "reps": e2e_reps, "seconds_per_rank": per_rank_s,
                    "copy_gbs_per_rank": [((blk_bytes + (out_bytes if (not do_reduce or r == 0) else 0)) // io_div) * ke / s_ / 1e9
                                          for r, s_ in enumerate(per_rank_s)],
                    "frac_of_value": e2e_value / value,
                    "api": ("ConvolvePE(MixPE(SuperSawPE...), ir).render(start, 64) -> host Snippet (device-resident sources)"
                            if vb is not None else "ConvolveBank.submit/wait (pgx_bank_submit / pgx_bank_wait): pinned host buffers, "
                            f"{E2E_DEPTH} pulls in flight, every pull's H2D and D2H inside the timed region"
                            + ("; PGX_PULL_REDUCE: the mix is summed over the ranks on the device, D2H on rank 0 only" if do_reduce else "")),
                    "host_affinity_bound": numa_bound, "x_write_combined": bool(args.wc),
                    "checksum_mean_abs_y": checksum,
                    "multi_block": (dict(multi, value=(world * N * c_out * pull / sr) * M * multi["submits"] / multi["seconds"],
                                         frac_of_value=(world * N * c_out * pull / sr) * M * multi["submits"] / multi["seconds"] / value)
                                    if multi and "seconds" in multi else multi)},
            "gpu_launches": launches, "host_enqueue_ms_per_step": float(np.median(host_reps)),
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": kbytes, "mean_launch_ms": klaunch_ms,
                         "time_tile": tile,
                         "note": (f"time-tiled pass: one launch serves {tile} block steps and moves every delay-line row once, so "
                                  f"its algorithmic bytes per STEP are 1/{tile} of the per-block schedule's (per_block_pass_bytes) "
                                  "that SURVEY 8d's ceiling assumes; frac is this launch's own bytes over its own time. The "
                                  "per-block kernel (PGX_TILE=1) sits at frac ~1.0 and ~2.9x lower throughput "
                                  "(profiles/r02c_bench_c2_untiled.json)") if tile > 1 and not mix else None,
                         "per_block_pass_bytes": per_block_mac_bytes if tile > 1 else None,
                         "fused_kernel_min_bytes": ((N * info.c_x * Bw * 12 + N * c_out * Bw * 12) if not mix else
                                                    (N * Bw * 8 + N * c_out * Bw * 8)) if (mac_per_step == 0 and prof.ms_conv1 > 0) else None,
                         "timing": "UN-instrumented timed loop (CUDA events on the launching stream around K steps, median "
                                   "repetition): the kernel is launched once per step and its launches run back to back, so "
                                   "time per launch = ms_per_step; K1/K2 overlap it on other streams, whatever they do not "
                                   "hide is charged to the kernel",
                         "launch_plan": {"grid": info.mac_grid, "term_splits": info.mac_split,
                                         "streams_per_cta": info.mac_stream_tile, "ctas_per_sm": info.mac_occupancy},
                         "instrumented": {
                             "note": f"second pass of the same loop, {prof.steps} steps, CUDA events around EVERY kernel on its "
                                     "launching stream (this perturbs the overlap: not used for frac)",
                             "k_fdl_mac_busy_ms_per_launch": mac_ms_union, "k_fdl_mac_start_to_end_ms": mac_ms_each,
                             "k_fdl_mac_gbs_busy": (mac_bytes / (mac_ms_union * 1e-3) / 1e9) if mac_ms_union > 0 else None,
                             "launches_timed": int(prof.n_mac), "share_of_step": prof.ms_mac / max(ksum, 1e-12),
                             "k_fdl_mac_busy_fraction_of_step": mac_ms_union * mac_per_step / (ms_med / K),
                             "kernel_ms": {"k_r2c_ingest": prof.ms_r2c / nprof, "k_fdl_mac": mac_ms_union,
                                           "k_c2r_emit": prof.ms_c2r / nprof, "k_reduce_partials": prof.ms_fold / nprof,
                                           "k_fdl_mac_present_slot": prof.ms_now / nprof,
                                           "k_conv1_or_mix1_fused": prof.ms_conv1 / nprof, "sum": ksum / nprof}},
                         "step": {"algorithmic_bytes": step_bytes, "ms": ms_med / K,
                                  "achieved": step_bytes / (ms_med / K * 1e-3) / 1e9,
                                  "frac": step_bytes / (ms_med / K * 1e-3) / 1e9 / peak,
                                  "note": ("SURVEY 8d's per-step bytes assume one delay-line read per block; the tiled pass "
                                           f"reads it once per {tile} blocks, so this fraction can exceed 1") if tile > 1 and not mix else None}},
        }
        if reduce_info is not None:
            rec["reduce"] = reduce_info
    xp.free()
    for a_ in yps:
        a_.free()
    bank.close()
    if comm is not None:
        barrier()
        comm.close()
    return rec, spec


def run_gpu(args):
    import torch.distributed as dist

    world, rank, local, numa_bound = _setup_dist()
    rec, spec = measure(args, args.workload, world, rank, local, numa_bound)
    sub = None
    if world > 1 and args.workload == "c2" and not args.no_c4:
        # the sharded-mix configuration (BASELINE.json configs[3]) beside the headline: the one workload with an
        # exchange step, so that the driver's scaling runs see it
        sub, _ = measure(args, "c4", world, rank, local, numa_bound)
    ok = True
    if rank == 0:
        line = rec
        if sub is not None:
            line["c4"] = {k: sub[k] for k in ("value", "unit", "ms_per_step", "reps", "config", "parity", "e2e", "reduce",
                                              "gpu_launches", "roofline") if k in sub}
            line["c4"]["note"] = ("BASELINE configs[3]: 4096 streams x 2 s IRs = 512 per GPU at 8 GPUs (here 512 per GPU "
                                  f"x {world}), fused per-GPU mix + cross-GPU sum every 512-sample pull")
        if not args.no_cpu and world == 1 and args.workload == "c2":
            line["cpu_baseline"] = cpu_baseline_single(args.cpu_seconds)
        elif not args.no_cpu and world == 1 and args.workload == "c5v":
            line["cpu_baseline"] = cpu_baseline_c5v(spec["config"]["voices"], args.cpu_seconds)
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
        for r in (rec, sub):
            if r is not None and r.get("parity") is not None and not r["parity"]["ok"]:
                ok = False
                print(f"PARITY FAILED: {r['parity']}", file=sys.stderr, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="shared", choices=["shared", "distinct"],
                    help="one IR shared by all streams (the named reverb) or one IR per stream")
    ap.add_argument("--streams", type=int, default=0, help="streams per GPU (default: the workload's named size)")
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5", "c5v"],
                    help="c2 = BASELINE.json configs[1] (headline); the others are the remaining configs, for profiling")
    ap.add_argument("--tail-block", type=int, default=0,
                    help="c2 / c5: two-level partitioning (not the named uniform configuration): the first TAIL_BLOCK "
                         "taps at the pull-sized block, the rest at TAIL_BLOCK")
    ap.add_argument("--reverb", action="store_true", help="c2: add ReverbPE's fused wet/dry output stage")
    ap.add_argument("--pcm16", action="store_true",
                    help="e2e leg with int16 PCM host buffers converted on the device (WAV staging, half the PCIe bytes)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--wc", action="store_true", help="e2e leg: input staging buffers in write-combined pinned memory")
    ap.add_argument("--reps", type=int, default=0, help="repetitions of the K-step timed loop (default: ~4 s worth, 5..60)")
    ap.add_argument("--e2e-blocks", type=int, default=1,
                    help="c2: also time the host API with this many 512-sample blocks per submit (fewer, larger copies); "
                         "reported as e2e.multi_block beside the named one-block-per-pull e2e")
    ap.add_argument("--no-c4", action="store_true", help="--gpus N>1: skip the sharded-mix (c4) sub-record")
    ap.add_argument("--reduce", default="p2p", choices=["p2p", "nccl"],
                    help="c4 under torchrun: the cross-GPU sum through pgx_mix_reduce (NVLink peer memory, default) or "
                         "NCCL dist.reduce (the library baseline it is measured against)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
