#!/usr/bin/env python
"""Throughput of the OMP hybrid-precoder row (svd + omp) on device-resident H-hat: packets/s, and the
FP64 rate of the correlation kernel (8 * n_rays * n_tx * Ns flops per tone and greedy round -- the contraction
At' * Wres of pg/ompdecomp.m:107) next to a cuBLAS ZGEMM of the same shape (torch.matmul, complex128) as the board's
own FP64 yardstick.  The oracle (numpy, one tone at a time like the reference) is timed on a few tones beside it."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mamimo_b200 as mm
from oracle import omp as oomp

for nt, nr, nsc, npkt, ns, nrf, nrays in ((32, 4, 1024, 500, 1, 1, 500), (32, 4, 234, 500, 1, 1, 500),
                                            (32, 4, 1024, 100, 2, 4, 500), (64, 8, 2048, 40, 4, 4, 500)):
    rng = np.random.default_rng(74)
    _, Hg = mm.synth.make_packets(74, 5, nt, nr, nsc, snr_db=10.0)
    At = np.exp(2j * np.pi * rng.random((nt, nrays)))
    H = torch.from_numpy(Hg).cuda().repeat(npkt // 5, 1, 1, 1).contiguous()
    with mm.Engine(nt, nr, nsc, mlp=False, max_pkts=npkt) as eng:
        eng.set_steering_dictionary(At)
        H128 = H.to(torch.complex128)
        _, V = eng.svd(H128, check_flags=False)
        for _ in range(2):
            eng.omp(V, ns, nrf, check_flags=False)
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        reps = 5
        e0.record()
        for _ in range(reps):
            eng.svd(H128, check_flags=False)
        e1.record()
        for _ in range(reps):
            eng.omp(V, ns, nrf, check_flags=False)
        e2.record()
        torch.cuda.synchronize()
        ms_svd, ms_omp = e0.elapsed_time(e1) / reps, e1.elapsed_time(e2) / reps
    flops = 8.0 * nrays * nt * ns * nrf * npkt * nsc
    # cuBLAS ZGEMM of the same contraction: [n_rays x n_tx] * [n_tx x (tones * ns)]
    A = torch.from_numpy(At.conj().T.copy()).cuda()
    W = torch.randn((nt, min(npkt * nsc * ns, 1 << 18)), dtype=torch.complex128, device="cuda")
    for _ in range(2):
        A @ W
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        A @ W
    e1.record()
    torch.cuda.synchronize()
    zgemm_tf = 8.0 * nrays * nt * W.shape[1] / (e0.elapsed_time(e1) / reps) / 1e9
    t0 = time.perf_counter()
    Fo = V[:1, :ns, :, :8].cpu().numpy()
    for k in range(8):
        oomp.precoder_for_subcarrier(Fo[0, :, :, k].T, At, nrf)
    cpu_ms_per_pkt = (time.perf_counter() - t0) / 8 * nsc * 1e3
    print(json.dumps({"shape": "%dx%dx%d" % (nt, nr, nsc), "packets": npkt, "Ns": ns, "NtRF": nrf, "rays": nrays,
                      "svd_ms": ms_svd, "omp_ms": ms_omp, "packets_per_s": npkt / (ms_svd + ms_omp) * 1e3,
                      "omp_fp64_tflops": flops / ms_omp / 1e9, "cublas_zgemm_tflops": zgemm_tf,
                      "oracle_numpy_ms_per_packet": cpu_ms_per_pkt}), flush=True)
