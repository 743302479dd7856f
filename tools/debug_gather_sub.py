#!/usr/bin/env python
"""Diagnostics: pipelined fused all-gather with two virtual ranks on one GPU at the bench shape; reports which rows differ
from the unsharded result for a given MAMIMO_GATHER_SUB / MAMIMO_GATHER_SMS."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mamimo_b200 as mm

nt, nr, nsc, hidden = 32, 4, 1024, (1024, 1024)
npk = int(sys.argv[1]) if len(sys.argv) > 1 else 500
x = mm.synth.make_pilots(nsc)
nets = mm.synth.make_nets(nsc, hidden, nsc)
Yg, _ = mm.synth.make_packets(1, 25, nt, nr, nsc, 10.0, x_tones=x)
Y = torch.from_numpy(np.concatenate([Yg] * (npk // 25 + 1))[:npk].copy()).cuda()
rows = nt * nr
for scale in (6, 0):
    engs = []
    for r in range(2):
        e = mm.Engine(nt, nr, nsc, hidden=hidden, precision="fp16x3", act_scale_log2=scale)
        e.set_pilots(x, None); e.load_weights(nets); engs.append(e)
    ptrs = [e.gather_create(2, r, npk) for r, e in enumerate(engs)]
    for e in engs:
        e.gather_connect([p[0] for p in ptrs], [p[1] for p in ptrs])
    st = torch.cuda.Stream()
    Hr = torch.zeros((npk * rows, nsc), device="cuda"); Hi = torch.zeros_like(Hr)
    ALL = 1 | 2 | 4 | 8
    for it in range(3):
        for e in engs:
            e.estimate_stages_raw(ALL, Y.data_ptr(), 0, npk, 0, Hr.data_ptr(), Hi.data_ptr(), st.cuda_stream)
        st.synchronize()
        gr, gi = engs[0].gather_planes()
        for r in range(2):
            a = gr[r * npk * rows:(r + 1) * npk * rows]
            bad = (a != Hr).any(dim=1).nonzero().flatten()
            badi = (gi[r * npk * rows:(r + 1) * npk * rows] != Hi).any(dim=1).nonzero().flatten()
            print("scale", scale, "iter", it, "rank-slot", r, "rows differing real/imag:", bad.numel(), badi.numel(),
                  (bad[:4].tolist(), bad[-4:].tolist()) if bad.numel() else "")
    ref_r, ref_i = engs[0].estimate(Y)
    print("   vs one-shot estimate: max abs diff", float((ref_r - Hr).abs().max()), "rel", float((ref_r - Hr).norm() / ref_r.norm()))
    for e in engs:
        e.close()
