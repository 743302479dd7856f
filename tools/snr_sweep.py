#!/usr/bin/env python
"""BASELINE.json configs[2]: tolerance against the reference (FP64 oracle) per SNR in {-25, ..., 10} dB at the real
shape (Nt 32, Nr 4, 1024 tones, 1024-1024 hidden, both nets), for both tensor-core schemes.  4 packets per SNR are
compared in full (the oracle's FP64 forward is the slow part).  One JSON line per (scheme, SNR): global rel-L2 and
NMSE_subk (BER_test_maMIMO_LTF.m:675-686) of H_LS and H_DNN."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mamimo_b200 as mm
from oracle import tables, postproc
from _util import oracle_full, rel_l2, nmse_per_packet

nt, nr, nsc, hidden, npkt = 32, 4, 1024, (1024, 1024), 4
x = mm.synth.make_pilots(nsc)
nets = mm.synth.make_nets(nsc, hidden, nsc)
P = tables.sylvester_hadamard(nt)
for prec in ("fp16x3", "tf32x3"):
    with mm.Engine(nt, nr, nsc, hidden=hidden, precision=prec) as eng:
        eng.set_pilots(x, None)
        eng.load_weights(nets)
        for si, snr in enumerate(range(-25, 11, 5)):
            Y, _ = mm.synth.make_packets(3, npkt, nt, nr, nsc, snr_db=float(snr), x_tones=x, first_pkt=npkt * si)
            Hr, Hi, Hls = eng.estimate(Y, want_ls=True)
            H, ref_r, ref_i = oracle_full(Y, P, x, 1, nets)
            e_ls = rel_l2(H, Hls)
            e_dnn = rel_l2(ref_r + 1j * ref_i, Hr.astype(np.float64) + 1j * Hi)
            n_ls = float(np.mean([postproc.nmse_subk(np.transpose(H[p], (2, 1, 0)), np.transpose(Hls[p], (2, 1, 0))) for p in range(npkt)]))
            n_dnn = nmse_per_packet(ref_r, ref_i, Hr, Hi, npkt, nr, nt)
            print(json.dumps({"precision": prec, "snr_db": snr, "rel_l2_H_ls": e_ls, "rel_l2_H_dnn": e_dnn,
                              "nmse_subk_H_ls": n_ls, "nmse_subk_H_dnn": n_dnn, "within_1e-5": bool(e_ls <= 1e-5 and e_dnn <= 1e-5)}), flush=True)
