#!/usr/bin/env python
"""What the host link gives: pinned-memory copy bandwidth H2D, D2H and both at once, 1 MiB (the e2e pull size at C2)
and 64 MiB transfers, CUDA-event timed.  Prints one JSON line (evidence for what bounds bench.py's e2e leg)."""
import json
import torch

dev = torch.device("cuda:0")
out = {}
for mb in (1, 64):
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    reps = 400 if mb == 1 else 20

    def run(h2d, d2h):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3

    for _ in range(2):
        run(True, True)
    t = run(True, False); out[f"h2d_{mb}MiB_GBs"] = n * reps / t / 1e9
    t = run(False, True); out[f"d2h_{mb}MiB_GBs"] = n * reps / t / 1e9
    t = run(True, True); out[f"both_{mb}MiB_GBs_per_direction"] = n * reps / t / 1e9
print(json.dumps(out))
