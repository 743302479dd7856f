"""N>1 path on CPU: world_size-2 gloo run of the packet-sharded estimator + chunked all-gather.
The per-rank compute is stood in for by the oracle (this tests the host-side sharding logic only)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_PKT, NT, NR, NSC, CHUNK = 7, 4, 2, 32, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    import mamimo_b200 as mm
    from oracle import tables
    from _util import oracle_full
    x = mm.synth.make_pilots(NSC)
    nets = mm.synth.make_nets(NSC, (16,), NSC)
    Y, _ = mm.synth.make_packets(21, N_PKT, NT, NR, NSC, snr_db=10.0, x_tones=x)
    return Y, x, nets, tables.sylvester_hadamard(NT), oracle_full


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mamimo_b200 as mm
    sharding = sys.modules["_mamimo_b200_pkg"].sharding
    Y, x, nets, P, oracle_full = _problem()
    rows = NT * NR

    def estimate_fn(lo, hi):          # stand-in for engine.estimate on this rank's GPU
        _, r, i = oracle_full(Y[lo:hi], P, x, 1, nets)
        return torch.from_numpy(r.astype(np.float32)), torch.from_numpy(i.astype(np.float32))

    se = sharding.ShardedEstimator(N_PKT, rows, NSC, CHUNK)
    out_r = torch.empty((se.gathered_rows(), NSC))
    out_i = torch.empty_like(out_r)
    scr_r = [torch.empty((se.chunk * rows, NSC)) for _ in range(se.n_chunks)]
    scr_i = [torch.empty((se.chunk * rows, NSC)) for _ in range(se.n_chunks)]
    se.run(estimate_fn, out_r, out_i, scr_r, scr_i)
    if rank == 1:                      # any rank holds the full gathered tensor
        np.savez(out_path, r=out_r.numpy(), i=out_i.numpy(), chunk=se.chunk, n_chunks=se.n_chunks)
    dist.barrier()
    dist.destroy_process_group()


def test_packet_range_partition():
    import mamimo_b200  # noqa: F401
    sharding = sys.modules["_mamimo_b200_pkg"].sharding
    for n in (0, 1, 7, 500, 24000):
        for w in (1, 2, 3, 8):
            ranges = [sharding.packet_range(n, r, w) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1 and sizes == sharding.shard_sizes(n, w)
    assert sharding.chunk_ranges(5, 2) == [(0, 2), (2, 4), (4, 5)]


@pytest.mark.timeout(120)
def test_sharded_estimate_equals_unsharded_world2(tmp_path):
    out_path = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    z = np.load(out_path)
    import mamimo_b200  # noqa: F401
    sharding = sys.modules["_mamimo_b200_pkg"].sharding
    Y, x, nets, P, oracle_full = _problem()
    _, ref_r, ref_i = oracle_full(Y, P, x, 1, nets)
    rows = NT * NR
    for p in range(N_PKT):
        g = sharding.gathered_row_index(p, N_PKT, 2, rows, int(z["chunk"]), None)
        assert np.array_equal(z["r"][g:g + rows], ref_r[p * rows:(p + 1) * rows].astype(np.float32)), p
        assert np.array_equal(z["i"][g:g + rows], ref_i[p * rows:(p + 1) * rows].astype(np.float32)), p
