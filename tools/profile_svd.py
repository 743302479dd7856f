import os, sys
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import mamimo_b200 as mm
_, Hg = mm.synth.make_packets(73, 5, 32, 4, 1024, snr_db=10.0)
H = torch.from_numpy(Hg).cuda().repeat(20, 1, 1, 1).contiguous()
with mm.Engine(32, 4, 1024, mlp=False, max_pkts=100) as eng:
    for _ in range(2):
        eng.svd(H, check_flags=False)
    torch.cuda.synchronize()
