"""Packet sharding across ranks (one process per GPU) and the one collective of the path.

Every (packet, rx, tx) pair is independent in LS and in both FC nets, so packets are
split into contiguous per-rank ranges -- each shard keeps whole [rx][sym][k] slabs and whole row tiles --
and the only exchange is an all-gather of the output planes.  The gather is issued per chunk so a
chunk's transfer over NVLink overlaps the next chunk's compute.

Pure host logic + torch.distributed; works on the gloo backend (CPU tensors) for tests and on NCCL
(CUDA tensors) on the B200 box.  No estimator math in here.
"""
import math


def packet_range(n_pkt, rank, world):
    """Contiguous, balanced split: the first (n_pkt % world) ranks get one extra packet."""
    base, extra = divmod(n_pkt, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_pkt, world):
    return [packet_range(n_pkt, r, world)[1] - packet_range(n_pkt, r, world)[0] for r in range(world)]


def chunk_ranges(n_local, chunk):
    """[(lo, hi)] covering [0, n_local) in steps of `chunk` packets."""
    chunk = max(1, int(chunk))
    return [(lo, min(n_local, lo + chunk)) for lo in range(0, n_local, chunk)]


def gathered_row_index(global_pkt, n_pkt, world, rows_per_pkt, chunk, padded_local):
    """Row offset of a global packet inside the chunk-major gathered buffer written by
    ShardedEstimator (layout [chunk][rank][pkt_in_chunk][rows_per_pkt]); used by tests and consumers."""
    for r in range(world):
        lo, hi = packet_range(n_pkt, r, world)
        if lo <= global_pkt < hi:
            local = global_pkt - lo
            c, within = divmod(local, chunk)
            chunk_rows_before = c * world * chunk
            return (chunk_rows_before + r * chunk + within) * rows_per_pkt
    raise IndexError(global_pkt)


class ShardedEstimator:
    """Runs `estimate_fn(local packets lo..hi) -> (H_real, H_imag)` chunk by chunk on this rank's shard and
    all-gathers every chunk's planes.  estimate_fn is the engine call on a GPU box (or any stand-in in
    CPU tests); tensors must be torch tensors on the backend's device."""

    def __init__(self, n_pkt, rows_per_pkt, d_out, chunk_pkts, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_pkt, self.rows_per_pkt, self.d_out = n_pkt, rows_per_pkt, d_out
        self.lo, self.hi = packet_range(n_pkt, self.rank, self.world)
        self.max_local = math.ceil(n_pkt / self.world)
        self.chunk = max(1, min(int(chunk_pkts), self.max_local))
        self.n_chunks = math.ceil(self.max_local / self.chunk)

    def gathered_rows(self):
        return self.n_chunks * self.world * self.chunk * self.rows_per_pkt

    def run(self, estimate_fn, out_real, out_imag, scratch_real, scratch_imag, async_op=True):
        """out_* : [gathered_rows(), d_out] destination (chunk-major, rank-minor), scratch_* :
        [chunk*rows_per_pkt, d_out] x n_chunks staging the local planes (padded with zeros for ranks whose
        shard is shorter).  Returns when every gather has completed."""
        dist = self.dist
        rpc = self.chunk * self.rows_per_pkt
        works = []
        for c in range(self.n_chunks):
            lo = self.lo + c * self.chunk
            hi = min(self.hi, lo + self.chunk)
            sr, si = scratch_real[c], scratch_imag[c]
            n = max(0, hi - lo)
            if n > 0:
                hr, hi_ = estimate_fn(lo, hi)
                sr[: n * self.rows_per_pkt].copy_(hr)
                si[: n * self.rows_per_pkt].copy_(hi_)
            if n * self.rows_per_pkt < rpc:
                sr[n * self.rows_per_pkt:].zero_()
                si[n * self.rows_per_pkt:].zero_()
            dst_r = out_real[c * self.world * rpc:(c + 1) * self.world * rpc]
            dst_i = out_imag[c * self.world * rpc:(c + 1) * self.world * rpc]
            works.append(dist.all_gather_into_tensor(dst_r, sr, group=self.group, async_op=async_op))
            works.append(dist.all_gather_into_tensor(dst_i, si, group=self.group, async_op=async_op))
        for w in works:
            if w is not None:
                w.wait()


def connect_fused_gather(engine, pkts_per_rank, group=None):
    """One process per GPU: allocate this rank's gathered planes in the engine, exchange CUDA IPC handles over
    torch.distributed, map every peer's planes and hand them to the engine.  Afterwards
    engine.estimate_stages_raw(stages | engine.STAGE_GATHER, ...) all-gathers H-hat from inside the final FC
    kernels.  Returns this rank's gathered planes as torch tensors."""
    import torch.distributed as dist
    from .engine import ipc_export, ipc_open
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    pr, pi = engine.gather_create(world, rank, pkts_per_rank)
    mine = (ipc_export(pr), ipc_export(pi))
    allh = [None] * world
    dist.all_gather_object(allh, mine, group=group)
    reals, imags = [], []
    for r in range(world):
        if r == rank:
            reals.append(pr); imags.append(pi)
        else:
            reals.append(ipc_open(allh[r][0])); imags.append(ipc_open(allh[r][1]))
    engine.gather_connect(reals, imags)
    dist.barrier(group)
    return engine.gather_planes()


def connect_symmetric_gather(engine, pkts_per_rank, group=None, require_multicast=False):
    """Same contract as connect_fused_gather, but the planes come from torch's symmetric-memory allocator
    (cuMemCreate + peer mappings + an NVSwitch multicast binding when the fabric has one) and are handed to the engine
    with mamimo_gather_attach.  With a multicast address every row leaves this GPU once (multimem.st) and the switch
    replicates it.  Returns (real, imag, has_multicast); the tensors stay alive with the returned objects."""
    import torch
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", engine.cfg.device)
    n = world * pkts_per_rank * engine.rows_per_pkt * engine.cfg.d_out
    pg = group if group is not None else dist.group.WORLD
    planes, handles = [], []
    for _ in range(2):
        t = symm_mem.empty(n, dtype=torch.float32, device=dev)
        t.zero_()
        h = symm_mem.rendezvous(t, pg.group_name)
        planes.append(t)
        handles.append(h)
    ptrs = []
    for t, h in zip(planes, handles):
        off = t.data_ptr() - int(h.buffer_ptrs[rank])
        if off < 0 or off % 128:
            raise RuntimeError("symmetric-memory tensor is not at a 128-byte offset of its buffer")
        ptrs.append(([int(p) + off for p in h.buffer_ptrs], (int(h.multicast_ptr) + off) if h.multicast_ptr else 0))
    has_mc = bool(ptrs[0][1] and ptrs[1][1])
    if require_multicast and not has_mc:
        raise RuntimeError("no NVSwitch multicast address for the gathered planes")
    engine.gather_attach(world, rank, pkts_per_rank, ptrs[0][0], ptrs[1][0], ptrs[0][1] if has_mc else 0,
                         ptrs[1][1] if has_mc else 0)
    engine._symm = (planes, handles)               # keep the allocation and its mappings alive with the engine
    torch.cuda.synchronize()
    dist.barrier(group)
    shape = (world * pkts_per_rank * engine.rows_per_pkt, engine.cfg.d_out)
    return planes[0].view(shape), planes[1].view(shape), has_mc
