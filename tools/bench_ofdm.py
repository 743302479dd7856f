#!/usr/bin/env python
"""Measurement of the OFDM demod front-end at the bench workload's shape:
500 packets x 4 rx x 32 symbols, FFT 1024 + CP 256 -> 1024... (config-2 analogue: all 1024 bins kept) and the
reference numerology (FFT 256 / CP 64 / 234 carriers).  Prints one JSON line per case: achieved GB/s against the
measured HBM peak, and the numpy CPU restatement beside it.  Test infrastructure (imports oracle)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mamimo_b200 as mm
from oracle import ofdm, tables

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
for name, fft_len, cp, nt, nr, npkt, car in (
        ("ref-numerology 256/64/234", 256, 64, 32, 4, 500, tables.carriers_locations()),
        ("config-2 analogue 1024/256/1024", 1024, 256, 32, 4, 500, np.arange(1, 1025, dtype=np.int32))):
    rng = np.random.default_rng(1)
    n_sc = len(car)
    x1 = (rng.standard_normal((4, nr, nt * (fft_len + cp))) + 1j * rng.standard_normal((4, nr, nt * (fft_len + cp)))).astype(np.complex64)
    xd = torch.from_numpy(np.concatenate([x1] * (npkt // 4))).cuda()
    Yd = torch.empty((npkt, nr, nt, n_sc), dtype=torch.complex64, device="cuda")
    with mm.Engine(nt, nr, n_sc, mlp=False, max_pkts=npkt) as eng:
        eng.set_ofdm(fft_len, cp, cp, car)
        lib = sys.modules["_mamimo_b200_pkg"]._capi.lib
        import ctypes as C
        st = torch.cuda.current_stream().cuda_stream
        call = lambda: lib.mamimo_ofdm_demod(eng._h, C.c_void_p(xd.data_ptr()), 0, npkt, C.c_void_p(Yd.data_ptr()), 1, C.c_void_p(st))
        for _ in range(3):
            assert call() == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
    err = float(np.linalg.norm(Yd[:4].cpu().numpy() - ofdm.ofdm_demod(x1, fft_len, cp, cp, car)) /
                np.linalg.norm(ofdm.ofdm_demod(x1, fft_len, cp, cp, car)))
    # algorithmic bytes: the FFT window (fft_len samples; the cyclic prefix is skipped, not read) in + kept carriers out
    bytes_alg = npkt * nr * nt * (fft_len * 8 + n_sc * 8)
    t0 = time.perf_counter()
    ofdm.ofdm_demod(np.concatenate([x1] * 8), fft_len, cp, cp, car)
    cpu_s = (time.perf_counter() - t0) / 32 * npkt
    print(json.dumps({"case": name, "ms": ms, "symbols_per_s": npkt * nr * nt / (ms * 1e-3), "rel_l2_vs_oracle": err,
                      "roofline": {"bound": "hbm", "achieved": bytes_alg / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                                   "unit": "GB/s", "frac": bytes_alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
                      "cpu_numpy_ms_same_batch": cpu_s * 1e3}), flush=True)
