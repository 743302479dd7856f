#!/usr/bin/env python
"""Device-resident throughput of the BASELINE.json configs that are not the bench line (configs[2..3]):
  C3: Nt32 Nr4 1024 sc, 3000-packet batch     C4: Nt64 Nr8 2048 sc, 3000-packet batch (FC 2048-1024-1024-2048)
A few distinct synthetic packets are tiled on the device up to the batch size (generating 23 GB on the host would
dominate the run).  One JSON line per config."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mamimo_b200 as mm

for name, nt, nr, nsc, npkt in (("C3 32x4x1024 x3000", 32, 4, 1024, 3000), ("C4 64x8x2048 x3000", 64, 8, 2048, 3000)):
    x = mm.synth.make_pilots(nsc)
    nets = mm.synth.make_nets(nsc, (1024, 1024), nsc)
    Yg, _ = mm.synth.make_packets(4, 4, nt, nr, nsc, snr_db=10.0, x_tones=x)
    Yd = torch.from_numpy(Yg).cuda().repeat((npkt + 3) // 4, 1, 1, 1)[:npkt].contiguous()
    rows = npkt * nt * nr
    Hr = torch.empty((rows, nsc), dtype=torch.float32, device="cuda")
    Hi = torch.empty_like(Hr)
    with mm.Engine(nt, nr, nsc, hidden=(1024, 1024), precision="fp16x3") as eng:
        eng.set_pilots(x, None)
        eng.load_weights(nets)
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            eng.estimate_raw(Yd.data_ptr(), 0, npkt, 0, Hr.data_ptr(), Hi.data_ptr(), 1, st)
        torch.cuda.synchronize()
        eng.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng.estimate_raw(Yd.data_ptr(), 0, npkt, 0, Hr.data_ptr(), Hi.data_ptr(), 1, st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        prof = eng.profile_end()
    flop = nt * nr * 2 * 2 * (nsc * 1024 + 1024 * 1024 + 1024 * nsc) * npkt
    print(json.dumps({"config": name, "ms_per_batch": ms, "packets_per_s": npkt / (ms * 1e-3),
                      "fc_tflops": flop / (prof["fc_ms"] / 3 * 1e-3) / 1e12,
                      "ls_gbs": nr * nt * nsc * 16 * npkt / (prof["ls_ms"] / 3 * 1e-3) / 1e9,
                      "tiled_rows_equal": bool(torch.equal(Hr[: 4 * nt * nr], Hr[4 * nt * nr: 8 * nt * nr]))}), flush=True)
    del Yd, Hr, Hi
    torch.cuda.empty_cache()
