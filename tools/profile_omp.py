#!/usr/bin/env python
"""One OMP call at the bench shape (32x4x1024, 64 packets, 500 rays) for `ncu -k regex:omp_` captures:
argv[1] = Ns (default 1), argv[2] = NtRF (default 1)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mamimo_b200 as mm

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nrf = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(1)
nt, nr, nsc, npkt = 32, 4, 1024, 64
V = torch.from_numpy(rng.standard_normal((npkt, nr, nt, nsc)) + 1j * rng.standard_normal((npkt, nr, nt, nsc))).cuda()
with mm.Engine(nt, nr, nsc, mlp=False, max_pkts=npkt) as eng:
    eng.set_steering_dictionary(np.exp(2j * np.pi * rng.random((nt, 500))))
    for _ in range(2):
        eng.omp(V, ns, nrf)
    torch.cuda.synchronize()
print("done")
