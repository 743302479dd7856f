"""bench.py contract checks that need no GPU: the reference arm (CPU restatement of the reference's path) prints ONE
JSON line with the agreed keys, under torchrun-style env only rank 0 prints, and our arm refuses to run without a
CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                          timeout=300)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "4"])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "packets/s"
    assert d["metric"].startswith("channel-estimates/sec") and d["value"] > 0 and d["vs_baseline"] is None
    # `config` carries the workload only, so both arms print the SAME config dict (the driver compares them)
    assert d["config"] == {"workload": d["config"]["workload"]} and d["config"]["workload"].startswith("configs[1]")
    assert d["sample_pkts_per_step"] == 4
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "4 packets" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "packets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_configs_keep_metric_and_workload_in_step():
    r = _run(["--impl", "reference", "--config", "c4", "--steps", "1", "--warmup", "0", "--cpu-sample", "1"])
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["metric"] == "channel-estimates/sec (64x8, 2048-sc pkts)" and d["config"]["workload"].startswith("configs[3]")


def test_reference_arm_only_rank0_prints():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-sample", "2"],
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(["--steps", "1"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
