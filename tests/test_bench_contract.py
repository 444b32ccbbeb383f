"""The bench.py JSON contract, checked on the CPU through the reference arm (the C/OpenMP port of the reference
loop) on a tiny volume: one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--shape", "16,16,8",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "voxels/sec (GQI recon+peaks)" and d["unit"] == "voxels/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert set(d["config"]) >= {"workload"} and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0


import pytest


@pytest.mark.gpu
def test_our_arm_prints_one_contract_line_on_the_gpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--shape", "64,64,32", "--steps", "3", "--warmup", "3",
                          "--e2e-steps", "2", "--no-cpu"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert "impl" not in d and d["metric"] == "voxels/sec (GQI recon+peaks)" and d["unit"] == "voxels/s"
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["value"] > 0
    assert d["kernel"] == "tc" and d["gpu_launches"] >= 3 * 5          # the native kernels ran (no fallback exists)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1 and abs(r["achieved"] / r["peak"] - r["frac"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 64 * 64 * 32 * 288 * 4 + 64 * 64 * 32
    assert e["d2h_bytes_per_step"] == 64 * 64 * 32 * 4 * (321 + 9 + 3)
    assert 0 < e["frac_of_ceiling"] <= 1.2 and e["pcie_ceiling_GBps_each_way"] > 1
    c = d["clocks"]
    assert c["sm_max_mhz"] > 0 and isinstance(c["reasons"], list)
    # extra legs of the N = 1 line: pageable caller arrays (bounce ring) and odf = NULL
    assert d["e2e_pageable"]["value"] > 0 and d["e2e_pageable"]["d2h_bytes_per_step"] == e["d2h_bytes_per_step"]
    assert d["e2e_no_odf"]["value"] > 0 and d["e2e_no_odf"]["d2h_bytes_per_step"] == 64 * 64 * 32 * 4 * 12
    assert d["roofline"]["traffic_source"].startswith("committed ncu")


def test_reference_arm_under_torchrun_prints_once():
    """N > 1: the driver launches the reference arm through torchrun as well; rank 0 alone runs and prints."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--shape", "16,16,8", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
