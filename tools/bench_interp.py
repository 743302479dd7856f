#!/usr/bin/env python
"""LS + comb-pilot linear interpolation (the north_star's interpolation stage; the reference itself always uses
Nps = 1) at the bench shape: 32x4, 1024 tones, 500 packets, pilot spacing Nps = 1, 2, 4, 8.  Reports the LS kernel time
and GB/s on algorithmic bytes (the whole Y grid is in HBM and every 32-byte sector holds a pilot for Nps <= 4, so the
bytes are Y in + operand planes out as for Nps = 1) plus parity against the oracle."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import mamimo_b200 as mm
from _util import oracle_ls, rel_l2
from oracle import tables

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
nt, nr, nsc, npkt = 32, 4, 1024, 500
nets = mm.synth.make_nets(nsc, (1024, 1024), nsc)
for nps in (1, 2, 4, 8):
    n_pil = (nsc + nps - 1) // nps
    x = mm.synth.make_pilots(n_pil)
    Yg, _ = mm.synth.make_packets(3, 4, nt, nr, nsc, snr_db=10.0, x_tones=np.repeat(x, nps)[:nsc])
    Yd = torch.from_numpy(np.concatenate([Yg] * (npkt // 4))).cuda()
    Hls = torch.empty((npkt, nr, nt, nsc), dtype=torch.complex64, device="cuda")
    rows = npkt * nt * nr
    Hr = torch.empty((rows, nsc), dtype=torch.float32, device="cuda")
    Hi = torch.empty_like(Hr)
    with mm.Engine(nt, nr, nsc, n_ps=nps, hidden=(1024, 1024), precision="fp16x3") as eng:
        eng.set_pilots(x, None)
        eng.load_weights(nets)
        for _ in range(2):
            eng.estimate_raw(Yd.data_ptr(), 0, npkt, Hls.data_ptr(), Hr.data_ptr(), Hi.data_ptr(), 1, 0)
        torch.cuda.synchronize()
        eng.profile_begin()
        for _ in range(5):
            eng.estimate_raw(Yd.data_ptr(), 0, npkt, 0, Hr.data_ptr(), Hi.data_ptr(), 1, 0)
        prof = eng.profile_end()
    ms = prof["ls_ms"] / 5
    ref = oracle_ls(Yg, tables.sylvester_hadamard(nt), x, nps)
    err = rel_l2(ref, Hls[:4].cpu().numpy())
    gbs = nr * nt * nsc * 16 * npkt / (ms * 1e-3) / 1e9
    print(json.dumps({"n_ps": nps, "ls_ms": ms, "ls_gbs": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"], "rel_l2_vs_oracle": err}), flush=True)
