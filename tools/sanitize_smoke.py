#!/usr/bin/env python
"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): LS (Hadamard + dense + interp),
all FC kernels (pair, 1-CTA, SIMT), OFDM front-end and the fused gather on one GPU.
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mamimo_b200 as mm

nt, nr, nsc, hidden = 8, 2, 128, (128, 64)
x = mm.synth.make_pilots(nsc)
nets = mm.synth.make_nets(nsc, hidden, nsc)
Y, _ = mm.synth.make_packets(0, 3, nt, nr, nsc, snr_db=10.0, x_tones=x)
for prec, single in (("fp16x3", False), ("tf32x3", True), ("fp32_simt", False), ("bf16x1", False)):
    with mm.Engine(nt, nr, nsc, hidden=hidden, precision=prec, fc_single_cta=single) as eng:
        eng.set_pilots(x, None)
        eng.load_weights(nets)
        hr, hi, hls = eng.estimate(Y, want_ls=True)
        assert np.isfinite(hr).all() and np.isfinite(hi).all()
        if prec == "fp16x3":
            import torch
            eng.gather_create(1, 0, 3)
            eng.gather_connect([eng._gather[3]], [eng._gather[4]])
            Yd = torch.from_numpy(Y).cuda()
            eng.estimate_stages_raw(15, Yd.data_ptr(), 0, 3, 0, 0, 0, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            gr, gi = eng.gather_planes()
            assert np.array_equal(gr.cpu().numpy(), hr) and np.array_equal(gi.cpu().numpy(), hi)
    print("ok", prec, "single_cta" if single else "pair")
with mm.Engine(6, 2, 100, n_ps=3, mlp=False) as eng:           # dense P (DFT), interpolation, odd sizes
    eng.set_pilots(mm.synth.make_pilots(100, 3), np.fft.fft(np.eye(6)))
    Yq = (np.random.default_rng(0).standard_normal((2, 2, 6, 100)) + 0j).astype(np.complex64)
    assert np.isfinite(eng.ls_estimate(Yq)).all()
    print("ok dense/interp LS")
with mm.Engine(4, 2, 48, mlp=False) as eng:
    eng.set_ofdm(64, 16, 5, np.arange(9, 57))
    xs = (np.random.default_rng(1).standard_normal((2, 2, 4 * 80)) + 0j).astype(np.complex64)
    assert np.isfinite(eng.ofdm_demod(xs)).all()
    print("ok ofdm")
print("sanitize smoke done")
