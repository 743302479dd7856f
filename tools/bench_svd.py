#!/usr/bin/env python
"""Throughput of the per-subcarrier SVD row on device-resident H-hat: packets/s and achieved HBM GB/s
(algorithmic bytes: H read once + V1 written once + sigma; the second read of H is meant to hit L1/L2)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mamimo_b200 as mm

for nt, nr, nsc, npkt in ((32, 4, 1024, 500), (32, 4, 234, 500), (64, 8, 2048, 100)):
    _, Hg = mm.synth.make_packets(73, 5, nt, nr, nsc, snr_db=10.0)
    H = torch.from_numpy(Hg).cuda().repeat(npkt // 5, 1, 1, 1).contiguous()
    with mm.Engine(nt, nr, nsc, mlp=False, max_pkts=npkt) as eng:
        for _ in range(3):
            eng.svd(H, check_flags=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.svd(H, check_flags=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
    nbytes = npkt * nr * nt * nsc * 8 * 2 + npkt * nr * nsc * 4
    t0 = time.perf_counter()
    np.linalg.svd(np.transpose(Hg[:1], (0, 3, 1, 2)), full_matrices=False)
    cpu_ms_per_pkt = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"shape": "%dx%dx%d" % (nt, nr, nsc), "packets": npkt, "ms": ms, "packets_per_s": npkt / ms * 1e3,
                      "GB_per_s": nbytes / ms / 1e6, "numpy_lapack_ms_per_packet": cpu_ms_per_pkt}), flush=True)
