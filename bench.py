#!/usr/bin/env python
"""Benchmark of the hot path: LS + (interp) + real/imag FC denoiser, BASELINE.json metric
"channel-estimates/sec (32x4, 1024-sc pkts)", one channel estimate = one packet's full H-hat.

    python bench.py --gpus 1 --steps 20 --warmup 3            # our arm (prints ONE JSON line)
    python bench.py --impl reference --steps 3 --warmup 1     # CPU restatement of the reference, same metric
    torchrun ... bench.py --gpus N ...                        # N ranks, packets sharded, all-gather of H-hat

A step = one pass of the whole path over one 500-packet batch per GPU (configs[1] of BASELINE.json).
`value`  : packets/s with Y resident in HBM (device pointers through the C ABI), CUDA events, max over ranks.
`e2e`    : same metric through the same C-ABI call with pinned HOST buffers; H2D of Y and D2H of both planes
           inside the timed region (the library overlaps them with compute in chunks).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NT, NR, NSC, HIDDEN, SNR_DB = 32, 4, 1024, (1024, 1024), 10.0
WORKLOAD = "configs[1]: Nt=32 Nr=4, 1024 sc, 500-packet batch, SNR=10 dB, FC 1024-1024-1024-1024 x2 nets"
METRIC = "channel-estimates/sec (32x4, 1024-sc pkts)"
MLP_FLOP_PER_PKT = NT * NR * 2 * 2 * (NSC * HIDDEN[0] + HIDDEN[0] * HIDDEN[1] + HIDDEN[1] * NSC)
LS_BYTES_PER_PKT = NR * NT * NSC * 8 * 2                                                            # Y in + H planes out
SNR_LEVELS = None          # c3: per-packet SNR levels cycled through the batch

# BASELINE.json configs.  c2 = configs[1] is the bench line (the config the metric is quoted on); the others are the
# larger parity configs, timed with the same code so the driver can run them too.
CONFIGS = {
    "c2": dict(nt=32, nr=4, nsc=1024, npkt=500, e2e_pkts=None,
               workload="configs[1]: Nt=32 Nr=4, 1024 sc, 500-packet batch, SNR=10 dB, FC 1024-1024-1024-1024 x2 nets",
               metric="channel-estimates/sec (32x4, 1024-sc pkts)"),
    "c3": dict(nt=32, nr=4, nsc=1024, npkt=3000, e2e_pkts=1000, snr_levels=[-25, -20, -15, -10, -5, 0, 5, 10],
               workload="configs[2]: Nt=32 Nr=4, 1024 sc, 3000-packet batch, SNR sweep -25..10 dB (8 levels interleaved), "
                        "FC 1024-1024-1024-1024 x2 nets",
               metric="channel-estimates/sec (32x4, 1024-sc pkts)"),
    "c4": dict(nt=64, nr=8, nsc=2048, npkt=3000, e2e_pkts=250,
               workload="configs[3]: Nt=64 Nr=8, 2048 sc, 3000-packet batch, SNR=10 dB, FC 2048-1024-1024-2048 x2 nets",
               metric="channel-estimates/sec (64x8, 2048-sc pkts)"),
    "c5": dict(nt=64, nr=8, nsc=2048, npkt=125, e2e_pkts=None,
               workload="configs[4]: Nt=64 Nr=8, 2048 sc, 3000 packets per GPU as chunks of 125 per step, FC 2048-1024-1024-2048 "
                        "x2 nets, all-gather of H-hat per chunk",
               metric="channel-estimates/sec (64x8, 2048-sc pkts)"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index=0, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------- CPU restatement (reference arm)
def cpu_reference_run(n_pkt_sample, steps, warmup):
    """The reference's own path restated on the host cores (MATLAB/TF cannot run here: SURVEY 8c):
    LS in complex128 (helperMIMOChannelEstimate.m:33-36, one batched matmul per packet) and the two Keras
    nets as torch CPU FP32 with batch = Nt*Nr per predict call (..._DNN.py:339), all host threads."""
    import torch
    from oracle import ls as o_ls, mlp as o_mlp
    import mamimo_b200_synth as synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x = synth.make_pilots(NSC)
    nets = synth.make_nets(NSC, HIDDEN, NSC)
    P = synth.sylvester(NT)
    snr = SNR_DB if SNR_LEVELS is None else np.resize(np.asarray(SNR_LEVELS, dtype=np.float64), n_pkt_sample)
    Y, _ = synth.make_packets(1, n_pkt_sample, NT, NR, NSC, snr, x_tones=x)
    Yt = torch.from_numpy(Y)
    tnets = {}
    for name in ("real", "imag"):
        tnets[name] = [(torch.from_numpy(np.asarray(L["W"], np.float32)), torch.from_numpy(np.asarray(L["b"], np.float32)),
                        None if L["bn"] is None else [torch.from_numpy(np.asarray(t, np.float32)) for t in L["bn"]])
                       for L in nets[name]]

    def predict(layers, xin):
        h = xin
        for i, (W, b, bn) in enumerate(layers):
            h = torch.addmm(b, h, W)
            if i < len(layers) - 1:
                h = torch.relu(h)
                if bn is not None:
                    g, be, mu, var = bn
                    h = g * (h - mu) / torch.sqrt(var + o_mlp.BN_EPS) + be
        return h

    def one_step():
        for p in range(n_pkt_sample):
            with torch.no_grad():
                H = o_ls.ls_estimate_torch(Yt[p:p + 1], P, x)      # complex128, as MATLAB computes
                X = H.reshape(-1, NSC)
                predict(tnets["real"], X.real.to(torch.float32))
                predict(tnets["imag"], X.imag.to(torch.float32))

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / steps
    return n_pkt_sample / dt, dt, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = args.cpu_sample
    val, dt, cores = cpu_reference_run(n_sample, max(1, args.steps), max(0, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "packets/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 LS + f32 FC (CPU)",
        "data": "synthetic", "config": {"workload": WORKLOAD},
        "sample_pkts_per_step": n_sample,
        "cpu_baseline": {"value": val, "unit": "packets/s", "cores": cores, "kind": "port",
                         "sample": "%d packets per step (torch-CPU c128 LS + fp32 FC, batch=Nt*Nr per predict)" % n_sample},
        "e2e": {"value": val, "unit": "packets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bind_to_gpu_numa_node(index):
    """Pin this rank's threads to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned host buffers are
    allocated (first touch places them on that node): the end-to-end path is PCIe/host-memory bound and with several
    ranks the default placement sends DMA traffic across sockets."""
    try:
        import torch
        prop = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read().strip())
        cpulist = open(base + "/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if node < 0 or not cpus or cpus == allowed:
            return "node %d (no change)" % node
        os.sched_setaffinity(0, cpus)
        return "node %d, %d cpus" % (node, len(cpus))
    except Exception as ex:                                  # containers often hide the topology: not an error
        return "unavailable (%s)" % type(ex).__name__


# --------------------------------------------------------------------------- our arm
def copy_ceiling_probe(torch, dev, world, dist, mib=256, reps=4):
    """Pinned-copy ceiling of this box measured in the run: plain cudaMemcpyAsync H2D alone, D2H alone and both at once
    (the e2e path is duplex), per rank and summed over ranks running at the same time.  GB/s."""
    n = mib << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def timed(h2d, d2h):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(dev)
        return reps * n / (time.perf_counter() - t0) / 1e9

    timed(True, True)
    res = {"h2d_alone": timed(True, False), "d2h_alone": timed(False, True), "duplex_each_way": timed(True, True)}
    if world > 1:
        t = torch.tensor([res["h2d_alone"], res["d2h_alone"], res["duplex_each_way"]], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        res.update({"aggregate_h2d_alone": float(t[0]), "aggregate_d2h_alone": float(t[1]), "aggregate_duplex_each_way": float(t[2])})
    return res


def traffic_from_profiles(kernel):
    """DRAM bytes per launch of `kernel` from the ncu --set full capture committed under profiles/ (a run under ncu is
    never a bench run, so this cannot be measured live); returns (bytes | None, source)."""
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(tp):
        return None, None
    d = json.load(open(tp)).get(kernel, {})
    return d.get("bytes_per_launch"), d.get("source")


def run_ours(args):
    import torch
    import torch.distributed as dist
    import mamimo_b200 as mm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a B200: no CUDA device visible, and there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_note = bind_to_gpu_numa_node(local_rank) if not args.no_numa_bind else "off"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    npkt = args.npkt
    cfg_id = 1
    x = mm.synth.make_pilots(NSC)
    nets = mm.synth.make_nets(NSC, HIDDEN, NSC)
    # per-rank packet shard: rank r owns global packets [r*npkt, (r+1)*npkt)  (weak scaling); gen_pkts distinct synthetic
    # packets are generated on the host and tiled ON THE DEVICE to npkt (c4: 23 GiB of Y never exists on the host)
    gen_pkts = min(npkt, args.gen_pkts)
    snr = SNR_DB if SNR_LEVELS is None else np.resize(np.asarray(SNR_LEVELS, dtype=np.float64), gen_pkts)
    Yg, _ = mm.synth.make_packets(cfg_id, gen_pkts, NT, NR, NSC, snr, x_tones=x, first_pkt=rank * npkt)
    reps = (npkt + gen_pkts - 1) // gen_pkts
    Yd = torch.from_numpy(Yg).to(dev).repeat(reps, 1, 1, 1)[:npkt].contiguous()
    rows = npkt * NT * NR
    Hr = torch.empty((rows, NSC), dtype=torch.float32, device=dev)
    Hi = torch.empty_like(Hr)
    e2e_n = 0 if args.no_e2e else min(npkt, args.e2e_pkts or npkt)
    e2e_rows = max(64, e2e_n * NT * NR)
    Yh = torch.from_numpy(np.concatenate([Yg] * ((max(e2e_n, 1) + gen_pkts - 1) // gen_pkts))[:max(e2e_n, 1)].copy()).pin_memory()
    Hr_h = torch.empty((e2e_rows, NSC), dtype=torch.float32).pin_memory()
    Hi_h = torch.empty((e2e_rows, NSC), dtype=torch.float32).pin_memory()
    gathered = None
    if world > 1:
        gathered = [torch.empty((world * rows, NSC), dtype=torch.float32, device=dev) for _ in range(2)]

    eng = mm.Engine(NT, NR, NSC, hidden=HIDDEN, precision=args.precision, max_pkts=args.max_pkts, device=local_rank,
                    fc_sm_reserve=(args.sm_reserve if (world > 1 and args.gather == "nccl") else 0))
    eng.set_pilots(x, None)
    eng.load_weights(nets)
    stream = torch.cuda.Stream(dev)          # non-default stream: the library replays the step as one CUDA-graph launch
    torch.cuda.set_stream(stream)

    # N > 1: the one collective of the path is the all-gather of the H-hat planes.
    #  --gather fused (default): the final FC layer of each net TMA-stores every output tile straight into every
    #      rank's gathered plane (peer memory over NVLink) -- compute and collective are one kernel; a 4-byte
    #      all-reduce per step is the cross-rank completion signal a consumer would need.
    #  --gather nccl: NCCL all-gather of each plane; the real plane's gather overlaps the imaginary net and the
    #      FC kernels leave --sm-reserve SMs free so the NCCL kernel never blocks a persistent CTA.
    #  --gather mc: the planes live in symmetric memory with an NVSwitch multicast binding; a multimem.st stream on a
    #      side stream sends every finished sub-batch ONCE and the switch replicates it into every rank's plane.
    fused = world > 1 and args.gather in ("fused", "mc")
    gather_note = args.gather
    if fused and args.gather == "mc":
        try:
            g_real, g_imag, _ = mm.sharding.connect_symmetric_gather(eng, npkt, require_multicast=True)
            okf = torch.ones(1, device=dev)
        except Exception as ex:
            okf = torch.zeros(1, device=dev)
            gather_note = "fused (multicast unavailable: %s)" % str(ex)[:80]
        dist.all_reduce(okf, op=dist.ReduceOp.MIN)
        if okf.item() == 0:
            args.gather = "fused"
            if gather_note == "mc":
                gather_note = "fused (multicast unavailable on a peer)"
    if fused and args.gather == "fused":
        try:
            g_real, g_imag = mm.sharding.connect_fused_gather(eng, npkt)
            okf = torch.ones(1, device=dev)
        except Exception as ex:                       # e.g. CUDA IPC / peer access unavailable on this box
            okf = torch.zeros(1, device=dev)
            gather_note = "nccl (fused unavailable: %s)" % str(ex)[:80]
        dist.all_reduce(okf, op=dist.ReduceOp.MIN)     # every rank must take the same path
        if okf.item() == 0:
            fused = False
            if not gather_note.startswith("nccl"):
                gather_note = "nccl (fused unavailable on a peer)"
    if fused:
        flag = torch.zeros(1, device=dev)
    ALL = eng.STAGE_LS | eng.STAGE_NET_REAL | eng.STAGE_NET_IMAG

    def step_device():
        if world == 1:
            eng.estimate_raw(Yd.data_ptr(), 0, npkt, 0, Hr.data_ptr(), Hi.data_ptr(), 1, stream.cuda_stream)
        elif fused:
            eng.estimate_stages_raw(ALL | eng.STAGE_GATHER, Yd.data_ptr(), 0, npkt, 0, 0, 0, stream.cuda_stream)
            dist.all_reduce(flag)
        else:
            eng.estimate_stages_raw(eng.STAGE_LS | eng.STAGE_NET_REAL, Yd.data_ptr(), 0, npkt, 0, Hr.data_ptr(), 0,
                                    stream.cuda_stream)
            w0 = dist.all_gather_into_tensor(gathered[0], Hr, async_op=True)
            eng.estimate_stages_raw(eng.STAGE_NET_IMAG, 0, 0, npkt, 0, 0, Hi.data_ptr(), stream.cuda_stream)
            w1 = dist.all_gather_into_tensor(gathered[1], Hi, async_op=True)
            w0.wait()
            w1.wait()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(3, args.warmup)):
        step_device()
    barrier()
    eng.poll_flags(stream.cuda_stream)           # a latched range / timeout condition must fail the run, not time garbage

    # ---- timed region 1: device-resident (value)
    sampler = ClockSampler(local_rank)
    sampler.start()
    st0 = eng.stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()                      # clocks / throttle reasons DURING the value region only
    st1 = eng.stats()
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    graph_launches = st1["graph_launches"] - st0["graph_launches"]
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * npkt / (ms_step * 1e-3)
    eng.poll_flags(stream.cuda_stream)

    # ---- the same step back to back for >= args.sustain seconds: the board settles under its power cap (value above is
    # what K steps after idle deliver; this is what a long-running job gets)
    value_sustained = None
    if args.sustain > 0:
        barrier()
        n_sus = 0
        t_end = time.perf_counter() + args.sustain
        ev0.record(stream)
        while time.perf_counter() < t_end:
            for _ in range(10):
                step_device()
            n_sus += 10
            stream.synchronize()
        ev1.record(stream)
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1) / n_sus], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        value_sustained = {"value": world * npkt / (float(t.item()) * 1e-3), "unit": "packets/s", "seconds": args.sustain,
                           "steps": n_sus, "ms_per_step": float(t.item())}

    # ---- the same K steps again with every kernel bracketed by CUDA events on its stream (live per-kernel-class
    # profile for the rooflines; plain launches on ONE stream, because events inside a replayed graph cannot be read
    # back and brackets on two interleaved streams would overlap).  One second of idle first: back to back, the second
    # region would be measured deeper into the board's power-capped regime than the first.
    time.sleep(1.0)
    for _ in range(3):
        step_device()
    if args.timeline and rank == 0:
        os.environ["MAMIMO_TIMELINE"] = args.timeline      # rank 0's per-launch start/end times of the profiled region
    eng.profile_begin()
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    barrier()
    ms_prof_step = ev0.elapsed_time(ev1) / args.steps
    prof = eng.profile_end()
    os.environ.pop("MAMIMO_TIMELINE", None)

    # compute-only (no all-gather) for N > 1, reported beside value
    value_compute_only = None
    if world > 1:
        barrier()
        ev0.record(stream)
        for _ in range(args.steps):
            eng.estimate_raw(Yd.data_ptr(), 0, npkt, 0, Hr.data_ptr(), Hi.data_ptr(), 1, stream.cuda_stream)
        ev1.record(stream)
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        value_compute_only = world * npkt / (float(t.item()) / args.steps * 1e-3)

    # ---- checks outside the timed regions
    # (1) fused gather: the planes written by the kernels of all ranks == an NCCL all-gather of the same data, bitwise
    # (2) parity: one gathered packet PER RANK against the FP64 oracle on rank 0 (N = 1: one packet of the batch)
    gather_check, parity = None, None
    if world > 1:
        if fused:       # same kernels, same schedule, local planes written too: NCCL-gather those and compare bit for bit
            Hr.zero_(); Hi.zero_(); g_real.zero_(); g_imag.zero_()      # a skipped or late write must show, not hide behind
            barrier()                                                   # an equal value left by an earlier step
            eng.estimate_stages_raw(ALL | eng.STAGE_GATHER, Yd.data_ptr(), 0, npkt, 0, Hr.data_ptr(), Hi.data_ptr(), stream.cuda_stream)
        else:
            eng.estimate_raw(Yd.data_ptr(), 0, npkt, 0, Hr.data_ptr(), Hi.data_ptr(), 1, stream.cuda_stream)
        dist.all_gather_into_tensor(gathered[0], Hr)
        dist.all_gather_into_tensor(gathered[1], Hi)
        barrier()
        if fused:
            same = torch.equal(gathered[0], g_real) and torch.equal(gathered[1], g_imag)
            if not same:        # say where: rows of which rank's slot differ (diagnostics on stderr, every rank)
                for name, a, b in (("real", gathered[0], g_real), ("imag", gathered[1], g_imag)):
                    bad = (a != b).any(dim=1).nonzero().flatten()
                    if bad.numel():
                        slots = torch.unique(bad // rows).tolist()
                        sys.stderr.write("[rank %d] fused != nccl on the %s plane: %d rows, slots %s, first %d last %d, max |diff| %.3e; "
                                         "all-zero rows among them: nccl %d, fused %d\n"
                                         % (rank, name, bad.numel(), slots, int(bad[0]), int(bad[-1]), float((a - b).abs().max()),
                                            int((a[bad] == 0).all(dim=1).sum()), int((b[bad] == 0).all(dim=1).sum())))
            ok = torch.tensor([int(same)], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            gather_check = bool(ok.item())
    if rank == 0 and not args.no_parity_check:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle import tables as o_tables, postproc as o_post
        from _util import oracle_full
        planes = (g_real, g_imag) if fused else ((gathered[0], gathered[1]) if world > 1 else (Hr, Hi))
        if world == 1:
            eng.estimate_raw(Yd.data_ptr(), 0, npkt, 0, Hr.data_ptr(), Hi.data_ptr(), 1, stream.cuda_stream)
            torch.cuda.synchronize(dev)
        errs = []
        rpp = NT * NR
        for r in range(world):
            snr_r = SNR_DB if SNR_LEVELS is None else float(SNR_LEVELS[0])
            Yr_, _ = mm.synth.make_packets(cfg_id, 1, NT, NR, NSC, snr_r, x_tones=x, first_pkt=r * npkt)
            _, ref_r, ref_i = oracle_full(Yr_, o_tables.sylvester_hadamard(NT), x, 1, nets)
            got_r = planes[0][r * rows:r * rows + rpp].cpu().numpy()
            got_i = planes[1][r * rows:r * rows + rpp].cpu().numpy()
            errs.append(float(o_post.rel_l2(ref_r + 1j * ref_i, got_r.astype(np.float64) + 1j * got_i)))
        parity = {"rel_l2_vs_oracle_per_rank": errs, "max": max(errs), "bound": 1e-5, "ok": bool(max(errs) <= 1e-5),
                  "what": "first packet of every rank's shard, read from rank 0's %s planes" % ("gathered" if world > 1 else "output")}

    # ---- timed region 2: end to end through the C ABI with pinned HOST buffers
    def step_host():
        eng.estimate_raw(Yh.data_ptr(), 0, e2e_n, 0, Hr_h.data_ptr(), Hi_h.data_ptr(), 0)

    e2e_steps = 0 if e2e_n == 0 else args.steps
    for _ in range(2 if e2e_steps else 0):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()                      # synchronous: returns when H-hat is on the host
    torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / max(1, e2e_steps) if e2e_steps else float("inf")
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * e2e_n / float(t.item()) if e2e_steps else None
    checksum = float(Hr_h[:: max(1, e2e_rows // 64)].double().sum().item()) if e2e_steps else None    # device->host result actually read
    ceiling = copy_ceiling_probe(torch, dev, world, dist) if e2e_steps else None

    # ---- roofline of the dominant kernel (FC layers on the tensor pipe), from the live profile
    peaks, peak_src = measured_peaks()
    fc_ms_per_step = prof["fc_ms"] / args.steps
    ls_ms_per_step = prof["ls_ms"] / args.steps
    fc_tflops = MLP_FLOP_PER_PKT * npkt / (fc_ms_per_step * 1e-3) / 1e12 if fc_ms_per_step > 0 else 0.0
    fc_name = "fc_tc2_kernel" if args.precision != "fp32_simt" else "fc_simt_kernel"
    is_bench_shape = args.precision == "fp16x3" and args.config == "c2" and npkt == 500
    fc_traffic, fc_src = traffic_from_profiles(fc_name) if is_bench_shape else (None, None)
    ls_traffic, ls_src = traffic_from_profiles("ls_kernel") if is_bench_shape else (None, None)
    roofline = {"bound": "tensor", "kernel": fc_name,
                "achieved": fc_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": fc_tflops / peaks["bf16_tflops"], "peak_source": peak_src + " bf16 burst (cuBLAS)",
                "traffic": fc_traffic, "traffic_source": fc_src,
                "traffic_unit": "bytes/launch (ncu dram read+write, committed capture)",
                "flop_per_launch": MLP_FLOP_PER_PKT * npkt / max(1, prof["fc_launches"] // args.steps),
                "avg_launch_ms": prof["fc_ms"] / max(1, prof["fc_launches"]),
                "share_of_step": fc_ms_per_step / ms_prof_step if ms_prof_step > 0 else None,
                "mma_passes": {"tf32x3": 3, "fp16x3": 3, "bf16x1": 1, "fp32_simt": 0}[args.precision],
                "issue_rate_frac_of_peak": fc_tflops * {"tf32x3": 3, "fp16x3": 3, "bf16x1": 1, "fp32_simt": 0}[args.precision] / peaks["bf16_tflops"],
                "measured_on": "the profiled region (plain launches, nets back to back on one stream); the value region "
                               "replays the same kernels as one graph with the two nets on two branches"}
    ls_gbs = LS_BYTES_PER_PKT * npkt / (ls_ms_per_step * 1e-3) / 1e9 if ls_ms_per_step > 0 else 0.0
    roofline_ls = {"bound": "hbm", "kernel": "ls_kernel", "achieved": ls_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                   "frac": ls_gbs / peaks["hbm_gbs"], "avg_launch_ms": prof["ls_ms"] / max(1, prof["ls_launches"]),
                   "traffic": ls_traffic, "traffic_source": ls_src,
                   "bytes_per_launch": LS_BYTES_PER_PKT * npkt / max(1, prof["ls_launches"] // args.steps),
                   "note": "algorithmic bytes = Y in + one operand-plane set out"}
    stage_ms_per_step = prof["stage_ms"] / args.steps

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, dt, cores = cpu_reference_run(args.cpu_sample, 1, 1)
        cpu_baseline = {"value": val, "unit": "packets/s", "cores": cores, "kind": "port",
                        "sample": "%d packets (torch-CPU c128 LS + fp32 FC, batch=Nt*Nr per predict), %.1f s" % (args.cpu_sample, dt)}

    if rank == 0:
        io_gib = npkt * NT * NR * NSC * 8 / 2 ** 30
        if world == 1:
            launch_note = ("value: every step is ONE CUDA-graph launch per internal chunk (%d graph launches, %d kernels in the "
                           "timed region)" % (graph_launches, launches))
        else:
            launch_note = ("value: %d kernels in the timed region (two streams: the gathering layers run on a side stream), issued as "
                           "%d CUDA-graph launches, + one 4-byte all-reduce per step" % (launches, graph_launches))
        line = {
            "metric": METRIC, "value": value, "unit": "packets/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"tf32x3": "tf32x3 split (fp32-grade), fp32 accumulate", "fp16x3": "fp16x3 split (fp32-grade), fp32 accumulate",
                      "bf16x1": "bf16", "fp32_simt": "f32"}[args.precision],
            "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "detail": {"pkts_per_gpu": npkt, "precision": args.precision,
                       "fp16_operand_scale": "auto (per-call amax pre-pass + per-level device-side scales)" if args.precision == "fp16x3" else None,
                       "l2": "inputs (%.2f GiB Y) and outputs (%.2f GiB) per step exceed the 126 MB L2" % (io_gib, io_gib),
                       "launch": launch_note + "; rooflines: the same K steps re-run after 1 s idle as plain launches with "
                                               "per-kernel CUDA events (ms_per_step_profiled)",
                       "numa_bind": numa_note,
                       "parallelism": "packets sharded over %d GPU(s)%s" % (
                           world, (", all-gather of H planes in step (%s)" % gather_note) if world > 1 else "")},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "packets/s", "h2d_bytes_per_step": int(e2e_n * NR * NT * NSC * 8),
                    "d2h_bytes_per_step": int(2 * e2e_n * NT * NR * NSC * 4), "checksum": checksum, "pkts_per_step": e2e_n,
                    "copy_ceiling_gbs": ceiling,
                    "achieved_gbs_each_way": (e2e_val / world * NR * NT * NSC * 8 / 1e9) if e2e_val else None},
            "gpu_launches": int(launches), "graph_launches": int(graph_launches),
            "roofline": roofline, "roofline_ls": roofline_ls,
            "amax_prepass_ms_per_step": stage_ms_per_step,
            "pair_estimates_per_s": value * NT * NR, "us_per_packet": 1e6 / value,
            "ms_per_step_profiled": ms_prof_step,
        }
        if value_sustained is not None:
            line["value_sustained"] = value_sustained
        if value_compute_only is not None:
            line["value_compute_only"] = value_compute_only
        if gather_check is not None:
            line["fused_gather_equals_nccl"] = gather_check
        if parity is not None:
            line["parity"] = parity
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("MAMIMO_BENCH_PRECISION", "fp16x3"),
                    choices=["tf32x3", "fp16x3", "bf16x1", "fp32_simt"])
    ap.add_argument("--npkt", type=int, default=500)
    ap.add_argument("--gen-pkts", type=int, default=125, help="distinct synthetic packets generated (tiled to npkt)")
    ap.add_argument("--max-pkts", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=500, help="packets per CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to its GPU's NUMA node")
    ap.add_argument("--gather", default="fused", choices=["fused", "mc", "nccl"], help="N>1: how H-hat is all-gathered")
    ap.add_argument("--sm-reserve", type=int, default=16, help="N>1: SMs left free for the concurrent NCCL kernels")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS),
                    help="c2 = BASELINE configs[1] (the bench line); c3 = configs[2] (3000 packets, 8 SNR levels); c4 = configs[3] "
                         "(Nt64 Nr8 2048 sc, 3000 packets); c5 = configs[4]: 3000 packets per GPU processed as 24 steps of 125 with "
                         "the fused all-gather (gathered chunk consumed/overwritten per step)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer region")
    ap.add_argument("--e2e-pkts", type=int, default=0, help="packets per end-to-end step (default: the batch, c4: 250)")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the oracle comparison after the timed regions")
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of back-to-back steps for value_sustained (0 = skip)")
    ap.add_argument("--timeline", default="", help="write rank 0's per-launch timeline of the profiled region to this CSV")
    args = ap.parse_args()
    global NT, NR, NSC, WORKLOAD, METRIC, MLP_FLOP_PER_PKT, LS_BYTES_PER_PKT, SNR_LEVELS
    c = CONFIGS[args.config]
    NT, NR, NSC, WORKLOAD, METRIC, SNR_LEVELS = c["nt"], c["nr"], c["nsc"], c["workload"], c["metric"], c.get("snr_levels")
    MLP_FLOP_PER_PKT = NT * NR * 2 * 2 * (NSC * HIDDEN[0] + HIDDEN[0] * HIDDEN[1] + HIDDEN[1] * NSC)
    LS_BYTES_PER_PKT = NR * NT * NSC * 8 * 2
    if args.npkt == 500:
        args.npkt = c["npkt"]
    if not args.e2e_pkts:
        args.e2e_pkts = c["e2e_pkts"] or 0
    if args.config in ("c4", "c5"):
        args.gen_pkts = min(args.gen_pkts, 5)
        args.cpu_sample = min(args.cpu_sample, 16)
    if args.config == "c3":
        args.gen_pkts = max(8, args.gen_pkts // 8 * 8)          # whole cycles of the 8 SNR levels
    # the synth module is pure numpy: load it standalone so the reference arm never touches the CUDA library
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "mamimo_b200_synth", os.path.join(ROOT, "dl-channel-estimation-mamimo_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["mamimo_b200_synth"] = mod
    spec.loader.exec_module(mod)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
