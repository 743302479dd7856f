#!/usr/bin/env python
"""The author's own timing chart (timing_cpu_vs_gpu_barplot.eps, BASELINE.md section 1: seconds per packet at Nt = 32,
Nr = 4, 234 tones: LS 1.80e-3 incl. ofdmdemod, LMMSE 1.139, DNN 5.99e-4) re-measured like for like on one B200 with the
reference numerology (FFT 256 / CP 64 / 234 carriers, mode-A network 10272 -> 1024 -> 1024 -> 234), batch of 500 packets,
device-resident buffers, CUDA events.  One JSON line per bar."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mamimo_b200 as mm

lib = sys.modules["_mamimo_b200_pkg"]._capi.lib
nt, nr, npkt, fft, cp = 32, 4, 500, 256, 64
car = mm.carriers_locations()
nsc = car.size
len_ltf = nt * (fft + cp)
rng = np.random.default_rng(0)
x1 = (rng.standard_normal((4, nr, len_ltf)) + 1j * rng.standard_normal((4, nr, len_ltf))).astype(np.complex64) * 0.05
xd = torch.from_numpy(np.concatenate([x1] * (npkt // 4))).cuda()
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# bar 1: LS = ofdmdemod + helperMIMOChannelEstimate  (generate_maMIMO_LTF.m:367-383)
Hls = torch.empty((npkt, nr, nt, nsc), dtype=torch.complex64, device="cuda")
with mm.Engine(nt, nr, nsc, mlp=False, max_pkts=npkt) as eng:
    eng.set_pilots(mm.vht_ltf256()[car - 1].astype(np.float64), None)
    eng.set_ofdm(fft, cp, cp, car)
    ms_ls = timed(lambda: lib.mamimo_estimate_time(eng._h, C.c_void_p(xd.data_ptr()), 0, npkt, C.c_void_p(Hls.data_ptr()),
                                                   None, None, 1, C.c_void_p(st)))
    # bar 2: LMMSE on top of the LS estimate, per-(packet, rx) SNR  (helperMIMOChannelEstimate.m:37-39)
    t_rms = rng.uniform(1.0, 5.0, npkt)
    snr = rng.uniform(0.0, 20.0, (npkt, nr))
    ms_mmse = timed(lambda: eng.lmmse(Hls, t_rms, snr), reps=5)
print(json.dumps({"bar": "LS (ofdmdemod + LS)", "author_s_per_packet": 1.80e-3, "b200_us_per_packet": ms_ls * 1e3 / npkt,
                  "b200_packets_per_s": npkt / (ms_ls * 1e-3), "speedup_vs_author_chart": 1.80e-3 / (ms_ls * 1e-3 / npkt)}), flush=True)
print(json.dumps({"bar": "LMMSE", "author_s_per_packet": 1.139, "b200_us_per_packet": ms_mmse * 1e3 / npkt,
                  "b200_packets_per_s": npkt / (ms_mmse * 1e-3), "speedup_vs_author_chart": 1.139 / (ms_mmse * 1e-3 / npkt)}), flush=True)

# bar 3: DNN = predict of nTX*nRX rows through the pipeline's own network (mode A), both nets
hidden, d_in = (1024, 1024), len_ltf + nt
nets = mm.synth.make_nets(d_in, hidden, nsc)
sr = torch.from_numpy(np.ascontiguousarray(np.concatenate([x1.real] * (npkt // 4)), dtype=np.float32)).cuda()
si = torch.from_numpy(np.ascontiguousarray(np.concatenate([x1.imag] * (npkt // 4)), dtype=np.float32)).cuda()
yr = torch.empty((npkt * nt * nr, nsc), dtype=torch.float32, device="cuda")
yi = torch.empty_like(yr)
with mm.Engine(nt, nr, 8, hidden=hidden, d_in=d_in, d_out=nsc, input_mode="time_p", len_ltf=len_ltf, precision="fp16x3",
               max_pkts=npkt) as eng:
    eng.load_weights(nets)
    ms_dnn = timed(lambda: lib.mamimo_predict_time(eng._h, C.c_void_p(sr.data_ptr()), C.c_void_p(si.data_ptr()), npkt,
                                                   C.c_void_p(yr.data_ptr()), C.c_void_p(yi.data_ptr()), 1, C.c_void_p(st)))
print(json.dumps({"bar": "DNN (mode A, both nets)", "author_s_per_packet": 5.99e-4, "b200_us_per_packet": ms_dnn * 1e3 / npkt,
                  "b200_packets_per_s": npkt / (ms_dnn * 1e-3), "speedup_vs_author_chart": 5.99e-4 / (ms_dnn * 1e-3 / npkt)}), flush=True)
