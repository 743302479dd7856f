#!/usr/bin/env python
"""GPU diagnostic: accuracy (vs FP64 oracle) and FC-kernel speed as a function of precision scheme and
kb_per_chunk (k-blocks accumulated inside the tensor core between FP32 register drains).
Writes one JSON line per case to stdout.  Test infrastructure (imports oracle)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import mamimo_b200 as mm
from oracle import tables
from _util import oracle_full, rel_l2

nt, nr, nsc, hidden = 32, 4, 1024, (1024, 1024)
x = mm.synth.make_pilots(nsc)
nets = mm.synth.make_nets(nsc, hidden, nsc)
Y, _ = mm.synth.make_packets(2, 3, nt, nr, nsc, snr_db=10.0, x_tones=x)
_, ref_r, ref_i = oracle_full(Y, tables.sylvester_hadamard(nt), x, 1, nets)
ref = ref_r + 1j * ref_i
Yb, _ = mm.synth.make_packets(1, 100, nt, nr, nsc, snr_db=10.0, x_tones=x)
Yd = torch.from_numpy(np.concatenate([Yb] * 5)).cuda()
npkt = Yd.shape[0]
Hr = torch.empty((npkt * nt * nr, nsc), dtype=torch.float32, device="cuda")
Hi = torch.empty_like(Hr)
cases = [("fp32_simt", 0)] + [(p, k) for p in ("tf32x3", "fp16x3", "bf16x1") for k in (1, 2, 4, 8, 1000)]
for prec, kbc in cases:
    with mm.Engine(nt, nr, nsc, hidden=hidden, precision=prec, kb_per_chunk=kbc) as eng:
        eng.set_pilots(x, None)
        eng.load_weights(nets)
        hr, hi = eng.estimate(Y)
        got = hr.astype(np.float64) + 1j * hi
        err = rel_l2(ref, got)
        shrink = float(np.mean((np.abs(got) - np.abs(ref)) / np.abs(ref)))
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            eng.estimate_raw(Yd.data_ptr(), 0, npkt, 0, Hr.data_ptr(), Hi.data_ptr(), 1, st)
        torch.cuda.synchronize()
        eng.profile_begin()
        t0 = time.perf_counter()
        for _ in range(5):
            eng.estimate_raw(Yd.data_ptr(), 0, npkt, 0, Hr.data_ptr(), Hi.data_ptr(), 1, st)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        prof = eng.profile_end()
    fc_ms = prof["fc_ms"] / 5
    print(json.dumps({"precision": prec, "kb_per_chunk": kbc, "rel_l2": err, "mean_rel_shrink": shrink,
                      "step_ms": dt * 1e3, "fc_ms": fc_ms, "ls_ms": prof["ls_ms"] / 5,
                      "fc_tflops": 1.6106e9 * npkt / (fc_ms * 1e-3) / 1e12}), flush=True)
