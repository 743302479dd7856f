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
# kernels added later in round 1: TMA-fed LS (plain + comb pilots), register-FFT OFDM kernels (bulk-copy fed and
# plain), LMMSE smoother (Toeplitz/Schur route and dense-Cholesky route), CUDA-graph replay
import torch
for nt2, nps in ((32, 1), (32, 4), (64, 2)):
    n2 = 128
    xp = mm.synth.make_pilots(n2, nps)
    xf = np.ones(n2)
    xf[::nps] = xp
    Y2, _ = mm.synth.make_packets(2, 2, nt2, 2, n2, snr_db=10.0, x_tones=xf)
    with mm.Engine(nt2, 2, n2, n_ps=nps, mlp=False) as eng:
        eng.set_pilots(xp, None)
        assert np.isfinite(eng.ls_estimate(Y2)).all()
    print("ok TMA LS", nt2, nps)
for fft, cp, off, ctype in ((256, 64, 64, np.complex64), (1024, 256, 100, np.complex64), (512, 32, 32, np.complex128), (4096, 64, 64, np.complex64)):
    car = np.arange(3, fft - 2)
    with mm.Engine(2, 1, car.size, mlp=False) as eng:
        eng.set_ofdm(fft, cp, off, car)
        xs = (np.random.default_rng(2).standard_normal((3, 1, 2 * (fft + cp))) + 0j).astype(ctype)
        assert np.isfinite(eng.ofdm_demod(xs)).all()
    print("ok ofdm", fft)
for schur in ("1", "0"):
    os.environ["MAMIMO_LMMSE_SCHUR"] = schur
    for nps in (1, 2):
        with mm.Engine(5, 2, 70, n_ps=nps, mlp=False) as eng:
            Hq = (np.random.default_rng(3).standard_normal((3, 2, 5, 70)) + 0j).astype(np.complex128)
            assert np.isfinite(eng.lmmse(Hq, 2.0, np.array([[3.0, 9.0]] * 3))).all()
    print("ok lmmse schur=" + schur)
nets2 = mm.synth.make_nets(128, (128, 64), 128)
Y3, _ = mm.synth.make_packets(3, 3, 32, 2, 128, snr_db=10.0, x_tones=mm.synth.make_pilots(128))
with mm.Engine(32, 2, 128, hidden=(128, 64), precision="fp16x3") as eng:
    eng.set_pilots(mm.synth.make_pilots(128), None)
    eng.load_weights(nets2)
    Yd = torch.from_numpy(Y3).cuda()
    Hr = torch.empty((3 * 64, 128), dtype=torch.float32, device="cuda")
    Hi = torch.empty_like(Hr)
    st = torch.cuda.Stream()
    for _ in range(2):
        eng.estimate_raw(Yd.data_ptr(), 0, 3, 0, Hr.data_ptr(), Hi.data_ptr(), 1, st.cuda_stream)
    st.synchronize()
    assert eng.stats()["graph_launches"] == 2 and torch.isfinite(Hr).all()
    print("ok graph replay")
print("sanitize smoke done")
