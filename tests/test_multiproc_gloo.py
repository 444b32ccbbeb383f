"""world_size-2 CPU tests (gloo) of the N > 1 host logic: max-over-ranks timing and whole-job throughput as
bench.py computes them, the library's z-slab partitioner, and the reference arm's "rank 0 only" rule."""
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
subs = [0, 1, 2] if rank == 0 else [3, 4]
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


def test_library_partitioner_balances_by_mask_count():
    """fibers_host_partition_slabs (what fibers_*(..., ngpu) and the batch queue use): contiguous, non-empty z-slabs that
    cover the volume, balanced by masked-voxel count rather than by z."""
    sys.path.insert(0, ROOT)
    from fibers_jl_b200 import _lib, phantom
    L = _lib.lib()
    for shape, ngpu in (((20, 16, 29), 8), ((9, 7, 3), 8), ((30, 30, 40), 2), ((12, 12, 24), 4)):
        mask = phantom.ellipsoid_mask(shape, 0.4)
        nxny, nz = shape[0] * shape[1], shape[2]
        n = min(ngpu, nz)
        out = np.zeros(2 * n, np.int64)
        assert L.fibers_host_partition_slabs(_lib.ptr(mask), nxny, nz, ngpu, _lib.ptr(out)) == 0
        r = out.reshape(n, 2)
        assert r[0, 0] == 0 and r[-1, 1] == nxny * nz
        assert np.all(r[1:, 0] == r[:-1, 1]) and np.all(r[:, 1] > r[:, 0]) and np.all(r % nxny == 0)
        if nz >= 4 * n:                         # balanced: no slab carries more than twice the mean masked count (+ one slice)
            flat = mask.reshape(-1, order="F")
            cnt = np.array([flat[a:b].sum() for a, b in r])
            per_slice = flat.reshape(nz, nxny).sum(axis=1).max()
            assert cnt.max() <= 2 * cnt.mean() + per_slice


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
