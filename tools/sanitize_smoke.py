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
# round 2: range bookkeeping (amax pre-pass, sampled + exact; provisional / verify / repair passes of the LS kernel),
# pipelined fused all-gather (sub-batches with row offsets, both variants), FP64 LS, per-subcarrier SVD (registers and
# shared-memory kernels), modes A / B with the automatic scale
Yb, _ = mm.synth.make_packets(4, 8, 32, 4, 1024, snr_db=10.0, x_tones=mm.synth.make_pilots(1024))
Yb = Yb.copy()
Yb[2, 1, :, 50] *= 3e4                                          # unsampled spike: forces the repair pass
netsb = mm.synth.make_nets(1024, (128,), 1024)
for mode in ("fused", "ce", "push"):
    os.environ["MAMIMO_GATHER_MODE"], os.environ["MAMIMO_GATHER_SUB"], os.environ["MAMIMO_GATHER_SMS"] = mode, "3", "36"
    with mm.Engine(32, 4, 1024, hidden=(128,), precision="fp16x3") as eng:
        eng.set_pilots(mm.synth.make_pilots(1024), None)
        eng.load_weights(netsb)
        hr, hi = eng.estimate(Yb)
        assert np.isfinite(hr).all()
        eng.gather_create(1, 0, 8)
        eng.gather_connect([eng._gather[3]], [eng._gather[4]])
        Yd = torch.from_numpy(Yb).cuda()
        eng.estimate_stages_raw(15, Yd.data_ptr(), 0, 8, 0, 0, 0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        gr, gi = eng.gather_planes()
        assert np.allclose(gr.cpu().numpy(), hr, rtol=0, atol=1e-4 * np.abs(hr).max())
    print("ok pipelined gather", mode)
for k in ("MAMIMO_GATHER_MODE", "MAMIMO_GATHER_SUB", "MAMIMO_GATHER_SMS"):
    os.environ.pop(k)
with mm.Engine(6, 2, 100, n_ps=3, mlp=False) as eng:
    eng.set_pilots(mm.synth.make_pilots(100, 3), np.fft.fft(np.eye(6)))
    Yq = (np.random.default_rng(0).standard_normal((2, 2, 6, 100)) + 0j).astype(np.complex128)
    assert np.isfinite(eng.ls_estimate(Yq)).all()
    print("ok FP64 LS")
for nt3, nr3 in ((8, 2), (32, 4), (16, 8)):
    with mm.Engine(nt3, nr3, 70, mlp=False) as eng:
        Hq = (np.random.default_rng(5).standard_normal((3, nr3, nt3, 70)) + 1j).astype(np.complex64)
        sg, v1 = eng.svd(Hq)
        assert np.isfinite(sg).all() and np.isfinite(v1).all()
    print("ok svd", nt3, nr3)
rng = np.random.default_rng(6)
with mm.Engine(1, 1, 1, n_ltf=1, hidden=(64,), d_in=96, d_out=40, input_mode="planes", precision="fp16x3") as eng:
    eng.load_weights(mm.synth.make_nets(96, (64,), 40))
    a, b = eng.predict_planes(rng.standard_normal((70, 96)).astype(np.float32) * 1e-3, rng.standard_normal((70, 96)).astype(np.float32))
    assert np.isfinite(a).all() and np.isfinite(b).all()
with mm.Engine(8, 2, 8, hidden=(64, 32), d_in=160 + 8, d_out=40, input_mode="time_p", len_ltf=160, precision="fp16x3") as eng:
    eng.set_pilots(None, mm.synth.sylvester(8))
    eng.load_weights(mm.synth.make_nets(168, (64, 32), 40))
    a, b = eng.predict_time(rng.standard_normal((3, 2, 160)), rng.standard_normal((3, 2, 160)))
    assert np.isfinite(a).all()
print("ok modes A/B auto scale")
# OMP hybrid precoder: both correlation kernels (Ns = 1 tall block, general), resident and reloaded residual tiles,
# every refit instantiation, ragged tone count
for nt4, ns4, nrf4, nsc4 in ((32, 1, 3, 70), (8, 2, 2, 100), (64, 4, 4, 65), (64, 6, 8, 33)):
    with mm.Engine(nt4, 8, nsc4, mlp=False) as eng:
        eng.set_steering_dictionary(np.exp(2j * np.pi * rng.random((nt4, 150))))
        Hq = (rng.standard_normal((2, 8, nt4, nsc4)) + 1j * rng.standard_normal((2, 8, nt4, nsc4))).astype(np.complex64)
        ix, er, fb = eng.omp_precoder(Hq, ns4, nrf4)
        assert (ix >= 0).all() and np.isfinite(fb).all() and np.isfinite(er).all()
    print("ok omp", nt4, ns4, nrf4)
print("sanitize smoke done")
