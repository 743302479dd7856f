#!/usr/bin/env python
"""Raw PCIe ceiling of the box: pinned H2D alone, D2H alone, and both at once on two streams (what the e2e
pipeline of mamimo_estimate(MEM_HOST) competes against).  Prints one JSON line."""
import json
import torch

n = 512 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
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
    return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


run(True, True, 1)
print(json.dumps({"h2d_alone_gbs": run(True, False), "d2h_alone_gbs": run(False, True),
                  "both_each_direction_gbs": run(True, True)}))
