"""world_size-2 CPU tests (gloo) of the N > 1 host logic: subject/slab sharding, max-over-ranks
timing and whole-job throughput, and the reference arm's "rank 0 only" rule."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, os.environ["FIBERS_ROOT"])
import torch, torch.distributed as dist
from fibers_jl_b200 import batch
rank, world, local = batch.rank_info()
dist.init_process_group("gloo", rank=rank, world_size=world)
subs = batch.assign_subjects(5, world, rank)
secs = 1.0 + rank                      # rank 1 is the slow one
units = 100.0 * len(subs)
worst = batch.reduce_max(secs, dist)
total = batch.reduce_sum(units, dist)
thr = batch.whole_job_throughput(units, secs, dist)
dist.barrier()
if rank == 0:
    print(json.dumps({"subs0": subs, "worst": worst, "total": total, "thr": thr, "world": world}))
dist.destroy_process_group()
'''


def test_gloo_world2_reductions(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, FIBERS_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2 and out["subs0"] == [0, 1, 2]
    assert out["worst"] == 2.0 and out["total"] == 500.0 and out["thr"] == 250.0


def test_assign_subjects_and_slabs():
    sys.path.insert(0, ROOT)
    from fibers_jl_b200 import batch
    for nsub in (0, 1, 7, 16):
        for world in (1, 2, 4, 8):
            got = [batch.assign_subjects(nsub, world, r) for r in range(world)]
            assert sorted(sum(got, [])) == list(range(nsub))
            sizes = [len(g) for g in got]
            assert max(sizes) - min(sizes) <= 1
    for nz, world in ((145, 8), (3, 8), (40, 2)):
        r = batch.slab_ranges(nz, 100, world)
        assert r[0][0] == 0 and r[-1][1] == nz * 100
        assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(b > a for a, b in r)
    with pytest.raises(ValueError):
        batch.assign_subjects(4, 2, 2)


def test_reference_arm_runs_only_on_rank0():
    """bench.py --impl reference under a 2-rank launch: rank 1 exits 0 without work or output."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1", "--shape", "24,20,16"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["n_gpus"] == 2
