#!/usr/bin/env python
"""Throughput of the reference-literal network (mode A: time-domain LTF (10240) || P row (32) -> 1024 -> 1024 -> 234,
full_pipeline_maMIMO_DNNEst.sh:40,47) with and without the de-duplicated first layer.  One JSON line per case."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mamimo_b200 as mm

nt, nr, len_ltf, d_out, hidden, npkt = 32, 4, 10240, 234, (1024, 1024), 128
d_in = len_ltf + nt
nets = mm.synth.make_nets(d_in, hidden, d_out)
rng = np.random.default_rng(0)
sr = mm.pinned_empty((npkt, nr, len_ltf), np.float32)
si = mm.pinned_empty((npkt, nr, len_ltf), np.float32)
sr[:] = rng.standard_normal(sr.shape) * 0.05
si[:] = rng.standard_normal(si.shape) * 0.05
out = {}
for dedup in (True, False):
    if dedup:
        os.environ.pop("MAMIMO_NO_DEDUP", None)
    else:
        os.environ["MAMIMO_NO_DEDUP"] = "1"
    with mm.Engine(nt, nr, 8, hidden=hidden, d_in=d_in, d_out=d_out, input_mode="time_p", len_ltf=len_ltf,
                   precision="fp16x3", max_pkts=npkt) as eng:
        eng.load_weights(nets)
        yr, yi = eng.predict_time(sr, si)
        t0 = time.perf_counter()
        for _ in range(5):
            yr, yi = eng.predict_time(sr, si)
        dt = (time.perf_counter() - t0) / 5
    out[dedup] = yr
    print(json.dumps({"mode_a_dedup": dedup, "packets": npkt, "ms_per_call_host_buffers": dt * 1e3,
                      "packets_per_s": npkt / dt}), flush=True)
print(json.dumps({"dedup_vs_dense_rel_l2": float(np.linalg.norm(out[True] - out[False]) / np.linalg.norm(out[False]))}))
