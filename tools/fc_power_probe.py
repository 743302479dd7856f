#!/usr/bin/env python
"""Is the FC stage limited by its pipeline or by the board's power/clock governor?  

For each variant the SAME FC-only work (64 000 rows through 1024-1024-1024-1024, both nets = 6 layer launches per
step, device-resident, mode B) runs back to back for >= `--seconds` while
  * pynvml samples SM clock, power draw and throttle reasons every ~10 ms, and
  * the kernel itself reports the clock it ran at: the MMA-issuing thread of every cluster reads %clock64 and
    %globaltimer at entry and exit (library built with -DMAMIMO_FC_DEBUG_COUNTERS, MAMIMO_FC_DEBUG=1), so
    cycles / ns = the SM clock DURING the kernel, and the role counters say what share of those cycles the issuing
    thread spent waiting for a drained TMEM buffer / for operands.
Variants: fp16x3 with accumulation chains 4 (default), 8, 1000 (no register drains: fewer cycles per tile), tf32x3,
bf16x1 (one MMA pass instead of three), and cuBLAS bf16 8192^3 (torch.matmul) as the board's own reference point.
One JSON line per variant.  If the kernel with fewer cycles per tile does NOT get proportionally faster because its
in-kernel clock falls, the stage is at the power wall and pipeline bubbles are not the lever.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Sampler(threading.Thread):
    def __init__(self, index=0, period=0.01):
        super().__init__(daemon=True)
        import pynvml
        pynvml.nvmlInit()
        self.nv, self.h, self.period = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index), period
        self.clk, self.pw, self.reasons = [], [], {}
        self._stop_evt = threading.Event()

    def run(self):
        nv = self.nv
        names = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake": 0x80}
        while not self._stop_evt.is_set():
            try:
                self.clk.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.pw.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons[k] = self.reasons.get(k, 0) + 1
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        import numpy as np
        self._stop_evt.set()
        self.join(timeout=2)
        c, p = np.asarray(self.clk, float), np.asarray(self.pw, float)
        half = len(c) // 2                                  # second half of the run = the settled regime
        return {"samples": int(len(c)), "sm_mhz_first_100ms": float(np.median(c[:10])) if len(c) >= 10 else None,
                "sm_mhz_median": float(np.median(c)) if len(c) else None,
                "sm_mhz_settled": float(np.median(c[half:])) if half else None,
                "sm_mhz_min": float(c.min()) if len(c) else None,
                "power_w_median": float(np.median(p)) if len(p) else None,
                "power_w_settled": float(np.median(p[half:])) if half else None,
                "power_w_max": float(p.max()) if len(p) else None,
                "reason_sample_counts": self.reasons}


def run_variant(precision, kbc, seconds, rows):
    import numpy as np
    import torch
    import mamimo_b200 as mm
    d = 1024
    nets = mm.synth.make_nets(d, (d, d), d)
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    Xr = torch.randn((rows, d), device=dev, generator=g)
    Xi = torch.randn((rows, d), device=dev, generator=g)
    flop_per_step = 2 * 2.0 * rows * 3 * d * d
    out = {"variant": "%s kb_per_chunk=%d" % (precision, kbc), "rows": rows}
    with mm.Engine(1, 1, 1, n_ltf=1, hidden=(d, d), d_in=d, d_out=d, input_mode="planes", precision=precision,
                   kb_per_chunk=kbc, max_pkts=rows) as eng:
        eng.load_weights(nets)
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            eng.predict_planes(Xr, Xi, check_flags=False)
        torch.cuda.synchronize()
        try:
            eng.debug_counters(reset=True)
            have_dbg = True
        except mm.MamimoError:
            have_dbg = False
        time.sleep(1.5)                                     # start every variant from an idle, cool-ish board
        # burst: the first 20 steps after idle (what a short bench run sees)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.predict_planes(Xr, Xi, check_flags=False)
        e1.record()
        torch.cuda.synchronize()
        burst_ms = e0.elapsed_time(e1) / 20
        dbg_burst = eng.debug_counters(reset=True) if have_dbg else None
        time.sleep(1.5)
        s = Sampler()
        s.start()
        t0 = time.perf_counter()
        steps, marks = 0, []
        e0.record()
        while time.perf_counter() - t0 < seconds:
            for _ in range(10):
                eng.predict_planes(Xr, Xi, check_flags=False)
            steps += 10
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append(ev)
            if steps % 50 == 0:
                torch.cuda.synchronize()                    # keep the launch queue short: the clock samples stay in phase
        e1.record()
        torch.cuda.synchronize()
        clocks = s.stop()
        total_ms = e0.elapsed_time(e1)
        last = marks[-6].elapsed_time(marks[-1]) / 50 if len(marks) >= 6 else total_ms / steps
        dbg = eng.debug_counters(reset=True) if have_dbg else None
        out.update({"burst_ms_per_step": burst_ms, "burst_tflops": flop_per_step / burst_ms / 1e9,
                    "sustained_ms_per_step": total_ms / steps, "sustained_tflops": flop_per_step / (total_ms / steps) / 1e9,
                    "last_50_steps_ms_per_step": last, "last_50_steps_tflops": flop_per_step / last / 1e9,
                    "steps": steps, "seconds": total_ms / 1e3, "clocks": clocks})
        for tag, c in (("burst", dbg_burst), ("sustained", dbg)):
            if c and c[5] and c[6]:
                out["in_kernel_" + tag] = {
                    "sm_ghz": c[4] / c[6], "cycles_per_cluster_launch": c[4] / c[5],
                    "mma_wait_tmem_frac": c[2] / c[4], "mma_wait_operands_frac": c[3] / c[4],
                    "producer_wait_stage_frac": c[0] / max(1, c[1]), "cluster_launches": c[5]}
    return out


def run_cublas(seconds):
    import torch
    n = 8192
    a = torch.randn((n, n), device="cuda", dtype=torch.bfloat16)
    b = torch.randn((n, n), device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    time.sleep(1.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        a @ b
    e1.record()
    torch.cuda.synchronize()
    burst_ms = e0.elapsed_time(e1) / 20
    time.sleep(1.5)
    s = Sampler()
    s.start()
    t0 = time.perf_counter()
    steps = 0
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(10):
            a @ b
        steps += 10
        if steps % 50 == 0:
            torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    clocks = s.stop()
    ms = e0.elapsed_time(e1) / steps
    f = 2.0 * n ** 3
    return {"variant": "cuBLAS bf16 8192^3 (torch.matmul)", "burst_ms_per_step": burst_ms, "burst_tflops": f / burst_ms / 1e9,
            "sustained_ms_per_step": ms, "sustained_tflops": f / ms / 1e9, "steps": steps, "clocks": clocks}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=2.5)
    ap.add_argument("--rows", type=int, default=64000)
    ap.add_argument("--variant", default=None, help="internal: run one variant in this process")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    if args.variant:
        if args.variant == "cublas":
            print(json.dumps(run_cublas(args.seconds)), flush=True)
        else:
            prec, kbc = args.variant.split(":")
            print(json.dumps(run_variant(prec, int(kbc), args.seconds, args.rows)), flush=True)
        return
    # one subprocess per variant: the debug library is selected by environment before the package is imported
    import importlib.util
    spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "dl-channel-estimation-mamimo_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    env = dict(os.environ, MAMIMO_LIB=b.DEBUG_LIB_PATH if os.path.exists(b.DEBUG_LIB_PATH) else b.LIB_PATH, MAMIMO_FC_DEBUG="1")
    lines = []
    for v in ("fp16x3:4", "fp16x3:8", "fp16x3:1000", "tf32x3:4", "bf16x1:4", "cublas"):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--variant", v, "--seconds", str(args.seconds),
                            "--rows", str(args.rows)], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if line:
            lines.append(line[-1])
            print(line[-1], flush=True)
        else:
            print(json.dumps({"variant": v, "error": (r.stderr or r.stdout)[-400:]}), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
