#!/usr/bin/env python
"""Single-call latency of the full path at small batches (BASELINE configs[0] is one packet): 32x4x1024, n packets,
device-resident buffers on a non-default stream (one CUDA-graph launch per call), host-synchronised per call, and
the same through pinned host buffers.  One JSON line per batch size."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mamimo_b200 as mm

nt, nr, nsc = 32, 4, 1024
x = mm.synth.make_pilots(nsc)
nets = mm.synth.make_nets(nsc, (1024, 1024), nsc)
with mm.Engine(nt, nr, nsc, hidden=(1024, 1024), precision="fp16x3", max_pkts=64) as eng:
    eng.set_pilots(x, None)
    eng.load_weights(nets)
    st = torch.cuda.Stream()
    for npkt in (1, 2, 4, 8, 16, 64):
        Y, _ = mm.synth.make_packets(5, npkt, nt, nr, nsc, snr_db=10.0, x_tones=x)
        Yd = torch.from_numpy(Y).cuda()
        Yh = mm.pinned_empty(Y.shape, np.complex64)
        Yh[:] = Y
        rows = npkt * nt * nr
        Hr = torch.empty((rows, nsc), dtype=torch.float32, device="cuda")
        Hi = torch.empty_like(Hr)
        Hr_h, Hi_h = mm.pinned_empty((rows, nsc), np.float32), mm.pinned_empty((rows, nsc), np.float32)
        lat = []
        for it in range(60):
            t0 = time.perf_counter()
            eng.estimate_raw(Yd.data_ptr(), 0, npkt, 0, Hr.data_ptr(), Hi.data_ptr(), 1, st.cuda_stream)
            st.synchronize()
            lat.append(time.perf_counter() - t0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(50):
            eng.estimate_raw(Yd.data_ptr(), 0, npkt, 0, Hr.data_ptr(), Hi.data_ptr(), 1, st.cuda_stream)
        e1.record(st)
        st.synchronize()
        lat_h = []
        for it in range(30):
            t0 = time.perf_counter()
            eng.estimate_raw(Yh.ctypes.data, 0, npkt, 0, Hr_h.ctypes.data, Hi_h.ctypes.data, 0)
            lat_h.append(time.perf_counter() - t0)
        print(json.dumps({"packets": npkt, "device_call_us_median": float(np.median(lat[10:]) * 1e6),
                          "device_back_to_back_us": e0.elapsed_time(e1) * 1e3 / 50,
                          "host_buffers_call_us_median": float(np.median(lat_h[5:]) * 1e6)}), flush=True)
