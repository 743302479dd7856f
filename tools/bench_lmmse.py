#!/usr/bin/env python
"""Measurement of the LMMSE smoother at the reference numerology (32x4, 234 tones) and at the
bench shape (32x4, 1024 tones): packets/s, FP64 FLOP rate (algorithmic FLOPs of the route in use, see below), and the numpy restatement of LMMSE_ce.m on the host beside it.
Test infrastructure (imports oracle)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mamimo_b200 as mm
from oracle import lmmse



def fp64_gemm_peak():
    """measured FP64 denominator, the way MEASURED_PEAKS.json measures bf16: cuBLAS DGEMM 4096^3, best of 5"""
    a = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
    b = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
    best = 1e9
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2 * 4096 ** 3 / (best * 1e-3) / 1e12


PEAK64 = fp64_gemm_peak()

for name, nt, nr, nsc, npkt, cpu_pairs in (("ref-numerology 32x4x234", 32, 4, 234, 500, 8), ("config-2 shape 32x4x1024", 32, 4, 1024, 32, 1)):
    rng = np.random.default_rng(2)
    H = (rng.standard_normal((npkt, nr, nt, nsc)) + 1j * rng.standard_normal((npkt, nr, nt, nsc))).astype(np.complex64)
    t_rms = rng.uniform(1.0, 5.0, npkt)
    snr = rng.uniform(0.0, 20.0, (npkt, nr))
    Hd = torch.from_numpy(H).cuda()
    with mm.Engine(nt, nr, nsc, mlp=False) as eng:
        for _ in range(2):
            out = eng.lmmse(Hd, t_rms, snr)
        torch.cuda.synchronize()
        eng.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            out = eng.lmmse(Hd, t_rms, snr)
        e1.record()
        torch.cuda.synchronize()
        prof = eng.profile_end()
        ms = e0.elapsed_time(e1) / reps
    ref = lmmse.lmmse_batched(H[:1], t_rms[:1], snr[:1])
    err = float(np.linalg.norm(out[:1].cpu().numpy() - ref) / np.linalg.norm(ref))
    n = nsc
    # per (pkt, rx) slab: Toeplitz (Schur) factorisation ~8 n^2 flops + two triangular solves with Nt right-hand
    # sides (Nt n^2 complex MACs, 8 flops each).  The dense-Cholesky route (MAMIMO_LMMSE_SCHUR=0) needs 8 n^3/3 more.
    schur = os.environ.get("MAMIMO_LMMSE_SCHUR", "1") != "0"
    flops = (8.0 * n * n * (nt + 1) if schur else 8.0 * (n ** 3 / 3.0 + nt * n * n)) * npkt * nr
    # CPU: the reference rebuilds + inverts per PAIR (LMMSE_ce.m is called inside the tx loop): time a few pairs
    t0 = time.perf_counter()
    for j in range(cpu_pairs):
        lmmse.lmmse_ce(H[0, 0, j % nt].astype(np.complex128), n, n, 1, np.array([1.0, 0, 0, 1]), 10.0)
    cpu_pair_s = (time.perf_counter() - t0) / cpu_pairs
    print(json.dumps({"case": name, "pkts": npkt, "ms": ms, "packets_per_s": npkt / (ms * 1e-3), "rel_l2_vs_oracle": err,
                      "roofline": {"bound": "fp64", "achieved": flops / (ms * 1e-3) / 1e12, "peak": PEAK64, "unit": "TFLOP/s",
                                   "frac": flops / (ms * 1e-3) / 1e12 / PEAK64, "peak_source": "measured cuBLAS DGEMM 4096^3 (this run)",
                                   "flops": "8 n^2 (Nt + 1) per (packet, rx) [Schur + 2 triangular solves]" if schur else "8 (n^3/3 + Nt n^2) per (packet, rx)"},
                      "launches": prof["lmmse_launches"] // reps,
                      "cpu_numpy_s_per_packet_literal": cpu_pair_s * nt * nr, "cpu_cores": os.cpu_count(),
                      "reference_published_s_per_packet": 1.139 if nsc == 234 else None}), flush=True)
