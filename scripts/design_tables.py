#!/usr/bin/env python
"""Regenerate the measured tables of DESIGN.md (between the <!-- BEGIN x --> / <!-- END x --> markers) from the bench
lines committed under profiles/: python scripts/design_tables.py"""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def line(name):
    path = os.path.join(P, name)
    if not os.path.exists(path):
        return None
    txt = [l for l in open(path).read().strip().splitlines() if l.startswith("{")]
    return json.loads(txt[-1]) if txt else None


def fmt_k(v):
    return f"{v:,.0f}".replace(",", " ")


def bench_table():
    rows = [("C2 shared IR, driver-shaped run (`--steps 20`, burst) — headline (time tile 4)", "r02c_bench_c2_k20.json"),
            ("C2 shared IR, sustained (`--steps 2000`, software power cap)", "r02c_bench_c2.json"),
            ("C2 shared IR, per-block pass (`PGX_TILE=1`: the round-1 schedule), same box, sustained", "r02c_bench_c2_untiled.json"),
            ("C2 shared IR, time tile 2 (`PGX_TILE=2`), same box", "r02c_bench_c2_tile2.json"),
            ("C2 distinct IRs (one per stream), time tile 4", "r02c_bench_c2_distinct.json"),
            ("C2 distinct IRs, per-block pass (`PGX_TILE=1`), same box", "r02c_bench_c2_distinct_untiled.json"),
            ("C1 4096 streams × 4096-tap FIR (P=1, `k_conv1_r16`)", "r02c_bench_c1.json"),
            ("C3 256 moving HRTF sources × 2 ears → stereo mix (one `k_mix1` launch per step)", "r02c_bench_c3.json"),
            ("C4 512 mono streams × 88 200 taps, fused mix (one GPU's share)", "r02c_bench_c4.json"),
            ("C5 single stream, 441 000 taps at 64-sample blocks (latency-bound)", "r02c_bench_c5.json")]
    out = ["| workload | ms/step (min…max over reps) | value (audio-s·ch/s) | e2e (of value) | dominant kernel: GB/s (frac of measured peak) | step-level frac | parity (max rel err, pulls) | SM MHz, reasons |",
           "|---|---|---|---|---|---|---|---|"]
    for label, f in rows:
        d = line(f)
        if d is None:
            continue
        r, e, pa, c = d["roofline"], d["e2e"], d.get("parity"), d["clocks"]
        out.append(f"| {label} | {d['ms_per_step']:.4f} ({d['reps']['ms_per_step_min']:.4f}…{d['reps']['ms_per_step_max']:.4f}, {d['reps']['n']} reps) | "
                   f"{fmt_k(d['value'])} | {fmt_k(e['value'])} ({e['frac_of_value']:.2f}) | {r['kernel'].split(' ')[0]}: {r['achieved']:.0f} ({r['frac']:.3f}) | "
                   f"{r['step']['frac']:.3f} | {pa['max_rel_err']:.1e}, {pa['pulls']} | {c['sm_mhz']}, {c['reasons'] or '—'} |")
    return "\n".join(out)


def multi_table():
    out = ["| GPUs | C2 ms/step | C2 value | C2 e2e (of value; copy GB/s per rank min…max) | C4 ms/step | C4 value | C4 e2e (of value) | cross-GPU sum: exposed µs per pull | parity C2 / C4 |",
           "|---|---|---|---|---|---|---|---|---|"]
    for n in (1, 2, 4, 8):
        d = line(f"r02c_scale_n{n}_k20.json") or line(f"r02b_scale_n{n}_k20.json")
        if d is None:
            continue
        e = d["e2e"]
        g = e["copy_gbs_per_rank"]
        c = d.get("c4")
        row = (f"| {n} | {d['ms_per_step']:.4f} | {fmt_k(d['value'])} | {fmt_k(e['value'])} ({e['frac_of_value']:.2f}; {min(g):.1f}…{max(g):.1f}) | ")
        if c:
            row += (f"{c['ms_per_step']:.4f} | {fmt_k(c['value'])} | {fmt_k(c['e2e']['value'])} ({c['e2e']['frac_of_value']:.2f}) | "
                    f"{c['reduce']['exposed_us_per_pull']:.2f} (`pgx_mix_reduce`) | {d['parity']['max_rel_err']:.1e} / {c['parity']['max_rel_err']:.1e} |")
        else:
            row += f"— | — | — | — | {d['parity']['max_rel_err']:.1e} / — |"
        out.append(row)
    d = line("r02b_scale_n8_k20_nccl.json")
    if d and "c4" in d:
        c = d["c4"]
        out.append(f"| 8, NCCL `dist.reduce` baseline (earlier visit, per-block pass) | {d['ms_per_step']:.4f} | {fmt_k(d['value'])} | — | {c['ms_per_step']:.4f} | {fmt_k(c['value'])} | "
                   f"{fmt_k(c['e2e']['value'])} ({c['e2e']['frac_of_value']:.2f}) | {c['reduce']['exposed_us_per_pull']:.2f} (NCCL) | — |")
    extra = []
    for tag, what in (("plain", "float32 staging"), ("wc", "write-combined input buffers"), ("pcm16", "int16 PCM staging (§7 rank 4)")):
        d = line(f"r02b_scale_n8_e2e_{tag}.json")
        if d:
            g = d["e2e"]["copy_gbs_per_rank"]
            extra.append(f"{what}: e2e {fmt_k(d['e2e']['value'])} = {d['e2e']['frac_of_value']:.2f} of value, {min(g):.1f}…{max(g):.1f} GB/s per rank")
    return "\n".join(out) + ("\n\n8-GPU e2e staging experiments (second box): " + "; ".join(extra) + "." if extra else "")


R1 = {"C1": 52, "C2 stereo": 54, "C2 as": 62, "C3 256 HRTF": 67, "C3 256 MOVING HRTF sources -> MixPE,": None,
      "C3 256 MOVING": 375, "C5 1024": 58, "C5 same": None}


def named_table():
    path = os.path.join(P, "r02d_named_configs.jsonl")
    if not os.path.exists(path):
        return ""
    out = ["| named configuration | pull | µs per pull (round 1) | × real time |", "|---|---|---|---|"]
    for l in open(path):
        d = json.loads(l)
        r1 = None
        for k, v in R1.items():
            if d["config"].startswith(k):
                r1 = v
                break
        out.append(f"| {d['config']} | {d['pull']} | {d['ms_per_pull'] * 1e3:.1f} ({r1 if r1 else '—'}) | {d['x_realtime']:.0f} |")
    return "\n".join(out)


def main():
    path = os.path.join(ROOT, "DESIGN.md")
    s = open(path).read()
    for key, fn in (("BENCH_TABLE", bench_table), ("MULTIGPU_NUMBERS", multi_table), ("NAMED_TABLE", named_table)):
        body = fn()
        block = f"<!-- BEGIN {key} -->\n{body}\n<!-- END {key} -->"
        if f"@@{key}@@" in s:
            s = s.replace(f"@@{key}@@", block)
        else:
            s = re.sub(rf"<!-- BEGIN {key} -->.*?<!-- END {key} -->", lambda m: block, s, flags=re.S)
    open(path, "w").write(s)


if __name__ == "__main__":
    main()
