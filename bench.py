#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Fibers.jl reconstruction path.

Metric (BASELINE.json): voxels/s of GQI ODF reconstruction + peak extraction on the synthetic
HCP-shaped volume 145 x 174 x 145 x 288 (18 b0 + 90 x b=1000/2000/3000), mask == 1, sphere_642.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shape nx,ny,nz]

One process per GPU (torchrun for N > 1).  The path shards by voxel/z-slab with NO collective, so
N ranks each reconstruct one HCP-shaped subject (weak scaling, cfg4-style batch); torch.distributed
is used only for the barrier and the max-over-ranks of the timed region.

  value     : device-resident throughput (inputs already in HBM), K steps, CUDA events, max over ranks
  e2e       : same metric through the host-pointer C-ABI call fibers_gqi_rec (what Julia ccalls):
              pinned HOST buffers in, HOST buffers out, H2D + D2H inside the timed region; reported beside the
              box's measured PCIe ceiling (concurrent H2D + D2H of pinned buffers) as `frac_of_ceiling`
  e2e_pageable : the same call on ordinary pageable numpy arrays (what a Julia Array is): the library's pinned
              bounce ring + host copy threads (N = 1 only)
  e2e_no_odf: the same call with odf = NULL (peaks + QA only, what stream() consumes)
  roofline  : dominant kernel (fused GQI contraction + peak epilogue) against the measured HBM peak
  cpu_baseline / --impl reference : the C/OpenMP port of the reference voxel loop (oracle/), all host
              cores, on a bounded z-sub-slab of the same workload
  zslab / batch (N > 1, rank 0 after the weak-scaling legs, other ranks idle): the IN-LIBRARY multi-GPU paths,
              fibers_gqi_rec(..., ngpu = N) on ONE cfg2 subject (strong scaling, host gather, host odfmax reduce) and
              fibers_dti_gqi_fit_batch on 16 cfg2-shaped subjects (BASELINE cfg4) over the N GPUs
  --mode zslab|batch [--config cfg2|cfg5] : only those paths, from a single process (python bench.py --mode zslab --gpus 8)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT]

HCP_SHAPE = (145, 174, 145)
NB0, SHELLS = 18, ((1000.0, 90), (2000.0, 90), (3000.0, 90))
M_VERT = 321


def algorithmic_bytes_per_voxel(nvol, nvert):      # 4N + 4M + 36 + 12 + 1
    return 4 * nvol + 4 * nvert + 36 + 12 + 1


def profiled_traffic(kernel, nvox):
    """DRAM bytes of one launch of the dominant kernel from the committed ncu --set full capture
    (profiles/roofline_traffic.json), or None when no capture exists for this kernel / size."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        d = json.load(open(p)).get(kernel)
        if d and int(d["nvox"]) == int(nvox):
            return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except (OSError, ValueError, KeyError):
        pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# synthetic HCP-shaped data, generated on the device (same signal model as phantom.gqi_phantom)
# ----------------------------------------------------------------------------------------------
def make_tables():
    from fibers_jl_b200 import phantom
    return phantom.shells_table(NB0, list(SHELLS))


def synth_dwi_device(torch, nvox, bval, bvec, seed, device, snr=30.0, chunk=1 << 19, pitch=None):
    """[nvol, pitch >= nvox] float32 on `device` (columns >= nvox are zero padding): two fibres + isotropic
    compartment, Rician noise, ~0.1 % negatives."""
    g = torch.Generator(device=device); g.manual_seed(seed)
    b = torch.tensor(bval, device=device, dtype=torch.float32)[:, None]
    G = torch.tensor(bvec, device=device, dtype=torch.float32)
    nvol = b.shape[0]
    out = torch.zeros((nvol, pitch or nvox), device=device, dtype=torch.float32)
    for c0 in range(0, nvox, chunk):
        n = min(chunk, nvox - c0)
        def dirs():
            v = torch.randn((3, n), generator=g, device=device)
            return v / v.norm(dim=0, keepdim=True)
        e1, e2 = dirs(), dirs()
        f1 = (0.3 + 0.4 * torch.rand((1, n), generator=g, device=device)) * 0.9
        f2 = 0.9 - f1
        S0 = 500 + 1000 * torch.rand((1, n), generator=g, device=device)
        sig = S0 / snr
        s = (f1 * torch.exp(-b * (2.0e-4 + 1.5e-3 * (G @ e1) ** 2)) + f2 * torch.exp(-b * (2.0e-4 + 1.5e-3 * (G @ e2) ** 2))
             + 0.1 * torch.exp(-b * 3.0e-3)) * S0
        s = torch.sqrt((s + sig * torch.randn((nvol, n), generator=g, device=device)) ** 2
                       + (sig * torch.randn((nvol, n), generator=g, device=device)) ** 2)
        neg = torch.rand((nvol, n), generator=g, device=device) < 1e-3
        s = torch.where(neg, -0.1 * s, s)
        out[:, c0:c0 + n] = s
        del s, neg, e1, e2
    return out


class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (NVML, 10 ms period; falls back to
    `nvidia-smi -lms 100` when pynvml is unavailable)."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
               "hw_power_brake_slowdown": 0x80}

    def __init__(self, index):
        self.index, self.samples, self.bits, self.stop_flag, self.thread, self.mx = index, [], 0, False, None, None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[index])
            except (ValueError, IndexError):
                pass
        return index

    def _loop(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.samples.append(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                self.bits |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nvml:
            try:
                self.mx = self.nvml.nvmlDeviceGetMaxClockInfo(self.h, self.nvml.NVML_CLOCK_SM)
            except Exception:
                self.mx = None
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        else:
            try:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                              "--format=csv,noheader,nounits", "-lms", "100"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            except OSError:
                self.proc = None

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.thread.join(timeout=1)
            reasons = sorted(k for k, b in self.REASONS.items() if self.bits & b)
            return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.mx,
                    "reasons": reasons, "samples": len(self.samples), "source": "nvml"}
        if not getattr(self, "proc", None):
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        time.sleep(0.15); self.proc.terminate()
        sm, mx = [], None
        for l in self.proc.stdout:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": [], "samples": len(sm),
                "source": "nvidia-smi"}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: C/OpenMP port of the reference voxel loop on a bounded sub-slab
# ----------------------------------------------------------------------------------------------
def cpu_sample(shape, bval, bvec, nz_sample, seed):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from fibers_jl_b200 import phantom
    rng = np.random.default_rng(seed)
    nx, ny, _ = shape
    nv = nx * ny * nz_sample
    S, *_ = phantom.multifibre_signal(rng, nv, bval, bvec, 30.0, lpar=1.7e-3, lperp=2.0e-4)
    dwi = np.asfortranarray(S.astype(np.float32).reshape((nx, ny, nz_sample, bval.shape[0]), order="F"))
    mask = np.ones((nx, ny, nz_sample), np.uint8, order="F")
    return dwi, mask


def run_cpu_reference(shape, steps, warmup, nz_sample=None, budget_s=12.0):
    """Times oracle/fibers_oracle.c (gqi voxel loop + peaks + odfmax post-pass) with all host threads.
    Returns (voxels_per_s, ms_per_step, cores, sample_description)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle as CO
    import fibers_oracle as O
    bval, bvec = make_tables()
    v, f = O.load_sphere(642)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)   # torchrun pins OMP_NUM_THREADS=1: ask for all cores explicitly
    setup = CO.GqiSetup(bval, bvec, v, f, 1.25)
    nx, ny, nz = shape
    if nz_sample is None:
        # calibrate on one slice, then size the sample for ~budget_s of CPU work in total
        dwi, mask = cpu_sample(shape, bval, bvec, 1, 7)
        outs = CO._recon_outputs(nx, ny, 1, 321)
        CO.gqi_rec(dwi, mask, setup=setup, outputs=outs, nthreads=cores)
        t = time.perf_counter(); CO.gqi_rec(dwi, mask, setup=setup, outputs=outs, nthreads=cores); dt = time.perf_counter() - t
        nz_sample = int(max(cores, min(nz, budget_s / max(dt, 1e-6) / max(1, steps + warmup))))
        nz_sample = max(cores, nz_sample // cores * cores)      # static z partition: keep threads balanced
    dwi, mask = cpu_sample(shape, bval, bvec, nz_sample, 11)
    outs = CO._recon_outputs(nx, ny, nz_sample, 321)
    for _ in range(warmup):
        CO.gqi_rec(dwi, mask, setup=setup, outputs=outs, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        for o in (outs[0], *outs[1], *outs[2]):
            o.fill(0)                                   # the reference allocates zero-filled outputs per call
        CO.gqi_rec(dwi, mask, setup=setup, outputs=outs, nthreads=cores)
    dt = (time.perf_counter() - t0) / steps
    nv = nx * ny * nz_sample
    return nv / dt, dt * 1e3, cores, f"z-sub-slab {nx}x{ny}x{nz_sample} of {nx}x{ny}x{nz} ({nv} voxels/step), C/OpenMP port of the reference loop, {cores} threads"


# ----------------------------------------------------------------------------------------------
# host-side helpers of the end-to-end legs
# ----------------------------------------------------------------------------------------------
def pcie_ceiling(torch, dev_index, nbytes=1 << 29, reps=4):
    """What this box moves between PINNED host memory and one GPU with H2D and D2H running at the same time
    (GB/s each way): the ceiling of every host-pointer call.  tools/gpu/pcie_probe.py is the long form."""
    dev = torch.device("cuda", dev_index)
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev); d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run(n):
        for _ in range(n):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(dev)
    run(1)
    t = time.perf_counter(); run(reps); dt = time.perf_counter() - t
    return nbytes * reps / dt / 1e9


class HostSubject:
    """One subject's host arrays for the C-ABI calls: pinned (torch) or pageable (numpy, Fortran order)."""

    def __init__(self, torch, shape, nvol, pinned, dwi_src=None, want_odf=True, want_dti=False, share_dwi=None):
        nvox = int(np.prod(shape))
        self.keep = []

        def alloc(rows, dtype=np.float32):
            if pinned:
                t = torch.empty((rows, nvox) if rows > 1 else (nvox,), dtype={np.float32: torch.float32, np.uint8: torch.uint8}[dtype], pin_memory=True)
                self.keep.append(t)
                return t.data_ptr(), t
            a = np.zeros((nvox, rows) if rows > 1 else (nvox,), dtype, order="F")     # [nvox, frames] column-major == [frames][nvox]
            self.keep.append(a)
            return a.ctypes.data, a
        if share_dwi is not None:
            self.dwi_ptr, self.dwi = share_dwi.dwi_ptr, share_dwi.dwi
            self.mask_ptr, self.mask = share_dwi.mask_ptr, share_dwi.mask
        else:
            self.dwi_ptr, self.dwi = alloc(nvol)
            self.mask_ptr, self.mask = alloc(1, np.uint8)
            if callable(dwi_src):                 # volumes that do not fit on one GPU: generated slab by slab (pinned only)
                self.mask.fill_(1)
                dwi_src(self.dwi)
            elif pinned:
                self.dwi.copy_(dwi_src[0][:, :nvox]); self.mask.copy_(dwi_src[1])
            else:
                self.dwi.T[...] = dwi_src[0][:, :nvox].cpu().numpy(); self.mask[...] = dwi_src[1].cpu().numpy()
        self.odf_ptr = alloc(M_VERT)[0] if want_odf else None
        self.peak = [alloc(3)[0] for _ in range(3)]
        self.qa = [alloc(1) for _ in range(3)]
        self.dti = [alloc(n)[0] for n in (1, 1, 1, 1, 3, 3, 3, 1, 1, 1)] if want_dti else None
        self.h2d = nvox * (4 * nvol + 1)
        self.d2h = nvox * 4 * ((M_VERT if want_odf else 0) + 9 + 3 + (16 if want_dti else 0))

    def gqi_out(self):
        return [self.odf_ptr] + self.peak + [q[0] for q in self.qa]


def timed_calls(fn, nwarm, nrep):
    for _ in range(nwarm):
        fn()
    ts = []
    for _ in range(nrep):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), float(np.mean(ts))


def library_multi_gpu_legs(torch, F, shape, bval, bvec, ngpu, dwi_src, nsub=16, reps=2, config="cfg2", only=None):
    """The in-library multi-GPU paths from ONE process (what a Julia session with FIBERS_CUDA_NGPU=N gets):
       zslab : fibers_gqi_rec(..., ngpu) on one subject = strong scaling, z-slabs balanced by mask count, host gather,
               host-side odfmax reduce (north_star "slab partitioner")
       batch : fibers_dti_gqi_fit_batch on `nsub` subjects (BASELINE cfg4), subjects queued over the GPUs."""
    import ctypes as C
    L = F._lib.lib()
    nvox, nvol = int(np.prod(shape)), int(bval.shape[0])
    V = np.asfortranarray(F.sphere_642.vertices); Fc = np.asfortranarray(F.sphere_642.faces); bv = np.asfortranarray(bvec)
    F.device.set_devices(list(range(ngpu)))
    out = {}
    subj = HostSubject(torch, shape, nvol, True, dwi_src)

    def zslab(s=subj):
        F._lib.check(L.fibers_gqi_rec(s.dwi_ptr, 0, s.mask_ptr, shape[0], shape[1], shape[2], nvol, F._lib.ptr(bval), F._lib.ptr(bv),
                                      F._lib.ptr(V), V.shape[0], F._lib.ptr(Fc), Fc.shape[0], 1.25, *s.gqi_out(), None, ngpu))
    if only in (None, "zslab"):
        med, mean = timed_calls(zslab, 2, max(3, reps))
        out["zslab"] = {"api": f"fibers_gqi_rec(..., ngpu={ngpu}): one {config} subject, z-slabs over {ngpu} GPUs, host gather + host odfmax reduce, no collective",
                        "scaling": "strong", "value": nvox / med, "unit": "voxels/s", "ms_per_call": med * 1e3, "ms_per_call_mean": mean * 1e3,
                        "h2d_bytes": subj.h2d, "d2h_bytes": subj.d2h, "host_memory": "pinned"}
    if only == "zslab":
        L.fibers_cuda_release_cache()
        return out
    # ---- cfg4: 16 subjects, DTI + GQI each.  The subjects share ONE synthetic DWI buffer (read-only); every subject owns its
    #      outputs.  Full outputs need 16 x 5.1 GB of pinned host memory: attempted only when the box has the RAM to spare.
    def batch_leg(want_odf):
        subs = [HostSubject(torch, shape, nvol, True, want_odf=want_odf, want_dti=True, share_dwi=subj) for _ in range(nsub)]
        parr = lambda xs: (C.c_void_p * len(xs))(*xs)
        dwi_t = parr([s.dwi_ptr for s in subs]); mask_t = parr([s.mask_ptr for s in subs])
        dti_t = parr([p for s in subs for p in s.dti]); gqi_t = parr([p for s in subs for p in s.gqi_out()])

        def call():
            F._lib.check(L.fibers_dti_gqi_fit_batch(nsub, dwi_t, mask_t, shape[0], shape[1], shape[2], nvol, F._lib.ptr(bval), F._lib.ptr(bv),
                                                    dti_t, F._lib.ptr(V), V.shape[0], F._lib.ptr(Fc), Fc.shape[0], 1.25, gqi_t, ngpu))
        med, mean = timed_calls(call, 1, reps)
        return {"api": f"fibers_dti_gqi_fit_batch: {nsub} {config}-shaped subjects, DTI + GQI each, queued over {ngpu} GPUs" + ("" if want_odf else ", odf = NULL"),
                "subjects": nsub, "value": nsub * nvox / med, "unit": "voxels/s", "s_per_batch": med, "ms_per_subject": med / nsub * 1e3,
                "h2d_bytes": nsub * subs[0].h2d, "d2h_bytes": nsub * subs[0].d2h, "host_memory": "pinned; the subjects share one DWI buffer, outputs are per subject"}
    try:
        out["batch_no_odf"] = batch_leg(False)
    except Exception as ex:          # noqa: BLE001
        out["batch_no_odf"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
    try:
        import psutil
        need = nsub * nvox * 4 * (M_VERT + 28) * 1.15
        if psutil.virtual_memory().available > need + 64e9:
            out["batch"] = batch_leg(True)
        else:
            out["batch"] = {"skipped": f"full outputs need {need / 1e9:.0f} GB of pinned host memory; available {psutil.virtual_memory().available / 1e9:.0f} GB"}
    except Exception as ex:          # noqa: BLE001
        out["batch"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
    L.fibers_cuda_release_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default=",".join(map(str, HCP_SHAPE)))
    ap.add_argument("--kernel", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--mask", default="ones", choices=["ones", "ellipsoid"],
                    help="ones: roofline run (every voxel computed); ellipsoid: ~25 %% fill brain-like mask (SURVEY 8d second run)")
    ap.add_argument("--mode", default="default", choices=["default", "zslab", "batch"],
                    help="zslab / batch: only the in-library multi-GPU legs, single process over --gpus devices")
    ap.add_argument("--subjects", type=int, default=16)
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg5"],
                    help="--mode zslab only: cfg5 = 400x400x300, 8 b0 + 120 x b=4000 (BASELINE configs[4]), data generated slab by slab")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-dti", action="store_true", help="skip the dti_fit / adc_fit legs (extra keys of the JSON line)")
    ap.add_argument("--no-multi", action="store_true", help="skip the in-library zslab / batch legs at N > 1")
    args = ap.parse_args()
    shape = tuple(int(x) for x in args.shape.split(","))
    nvox = int(np.prod(shape))
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    steps, warmup = args.steps, max(args.warmup, 0)
    bval, bvec = make_tables()
    nvol = bval.shape[0]
    workload = f"cfg2 GQI recon+peaks {shape[0]}x{shape[1]}x{shape[2]}x{nvol} (18 b0 + 90x b=1000/2000/3000), sphere_642, mask=={1 if args.mask == 'ones' else 'ellipsoid'}"
    config = {"workload": workload, "per_gpu": "one HCP-shaped subject per GPU (weak; cfg4-style batch)",
              "l2_policy": "inputs (4.2 GB/step) larger than L2 (126 MB); no explicit flush", "sigma": 1.25,
              "layout": "frame-major [frame][voxel]; dwi and output frame pitch = nvox rounded up to 64"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        vps, ms, cores, sample = run_cpu_reference(shape, steps, max(warmup, 1))
        line = {"impl": "reference", "metric": "voxels/sec (GQI recon+peaks)", "value": vps, "unit": "voxels/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": vps, "unit": "voxels/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": vps, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "host_cores": cores,
                "note": "Fibers.jl is pure Julia and Julia is not installed: this is the C/OpenMP restatement of its voxel loop, not Fibers.jl itself"}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import fibers_jl_b200 as F
    from fibers_jl_b200 import device as D
    if not torch.cuda.is_available() or F.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: libfibers_cuda has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    D.set_kernel(args.kernel)

    if args.mode != "default":              # in-library multi-GPU legs only (single process, no torch.distributed)
        ngpu = min(args.gpus, F.device_count())
        if args.config == "cfg5":
            from fibers_jl_b200 import phantom
            shape = (400, 400, 300) if args.shape == ",".join(map(str, HCP_SHAPE)) else shape
            nvox = int(np.prod(shape))
            bval, bvec = phantom.shells_table(8, [(4000.0, 120)])
            config["workload"] = f"cfg5 GQI recon+peaks {shape[0]}x{shape[1]}x{shape[2]}x{bval.shape[0]} (8 b0 + 120x b=4000), sphere_642, mask==1"

            def fill(h_dwi, step=1 << 21):
                for off in range(0, nvox, step):
                    n = min(step, nvox - off)
                    h_dwi[:, off:off + n].copy_(synth_dwi_device(torch, n, bval, bvec, 5000 + off // step, dev))
                torch.cuda.synchronize()
            src = fill
        else:
            src = (synth_dwi_device(torch, nvox, bval, bvec, 1000, dev), torch.ones(nvox, dtype=torch.uint8, device=dev))
        legs = library_multi_gpu_legs(torch, F, shape, bval, bvec, ngpu, src, nsub=args.subjects, reps=max(2, args.e2e_steps // 2),
                                      config=args.config, only=args.mode)
        key = "zslab" if args.mode == "zslab" else ("batch" if "value" in legs.get("batch", {}) else "batch_no_odf")
        line = {"metric": "voxels/sec (GQI recon+peaks)" if args.mode == "zslab" else "voxels/sec (DTI fit + GQI recon+peaks, batch of subjects)",
                "value": legs[key].get("value"), "unit": "voxels/s", "n_gpus": ngpu, "mode": args.mode, "scaling": "strong",
                "higher_is_better": True, "dtype": "f32", "data": "synthetic", "config": config, "host_cores": len(os.sched_getaffinity(0)), **legs}
        print(json.dumps(line))
        return 0

    dist = None
    side = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        side = dist.new_group(backend="gloo")          # host-side barrier for the legs in which only rank 0 works (no kernel spinning on the idle GPUs)
    D.set_devices([local_rank])

    # frame pitch of the DWI slab and of the outputs: rows 16-byte aligned (TMA) -- BENCH_PITCH_ALIGN / BENCH_PITCH_EXTRA (voxels) for experiments
    _pa = int(os.environ.get("BENCH_PITCH_ALIGN", "64")); _pe = int(os.environ.get("BENCH_PITCH_EXTRA", "0"))
    pitch = (nvox + _pa - 1) // _pa * _pa + _pe
    dwi = synth_dwi_device(torch, nvox, bval, bvec, 1000 + rank, dev, pitch=pitch)
    if args.mask == "ones":
        mask = torch.ones(nvox, dtype=torch.uint8, device=dev)
    else:
        ax = [torch.linspace(-1, 1, n, device=dev) for n in shape]
        r2 = (0.25 * 8 / (4 / 3 * np.pi)) ** (2 / 3)
        m3 = (ax[0][:, None, None] ** 2 + ax[1][None, :, None] ** 2 + ax[2][None, None, :] ** 2) <= r2
        mask = m3.permute(2, 1, 0).contiguous().reshape(-1).to(torch.uint8)          # x fastest (column-major volume)
        config["mask"] = f"ellipsoid, {mask.float().mean().item():.3f} fill; value counts ALL voxels of the volume"
    odf = torch.empty((M_VERT, pitch), dtype=torch.float32, device=dev)
    peak = [torch.empty((3, pitch), dtype=torch.float32, device=dev) for _ in range(3)]
    qa = [torch.empty(nvox, dtype=torch.float32, device=dev) for _ in range(3)]
    stats = torch.zeros(2, dtype=torch.int32, device=dev)
    plan = D.Plan("gqi", local_rank, bval, bvec, F.sphere_642, 1.25)
    stream = torch.cuda.current_stream().cuda_stream
    pk = [p.data_ptr() for p in peak]; qp = [q.data_ptr() for q in qa]

    def step(ev=None):
        D.stats_init(stats.data_ptr(), stream)
        if ev: ev[0].record()
        plan.recon(dwi.data_ptr(), pitch, mask.data_ptr(), nvox, pitch, odf.data_ptr(), pk, qp, stats.data_ptr(),
                   finalize=False, stream=stream)
        if ev: ev[1].record()
        D.qa_scale(qp, nvox, d_stats=stats.data_ptr(), stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the other half of BASELINE.json's metric ("GQI recon+peaks, DTI fit"): dti_fit and adc_fit on the SAME device-resident
    #      volume (cfg4 runs both on every subject), timed apart from the headline step: extra keys of the line, not part of `value`
    dti_legs = {}

    def run_dti_legs():
        nonlocal dti_legs
        try:
            pd = D.Plan("dti", local_rank, bval, bvec); pa = D.Plan("adc", local_rank, bval)
            douts = [torch.empty((n, pitch), dtype=torch.float32, device=dev) for n in (1, 1, 1, 1, 3, 3, 3, 1, 1, 1)]
            dptr = [o.data_ptr() for o in douts]
            peak_gbs0 = measured_peaks()[0]
            for name, fn, bpv_d in (("dti_fit", lambda: pd.dti_fit(dwi.data_ptr(), pitch, mask.data_ptr(), nvox, pitch, dptr, stream=stream), 4 * nvol + 65),
                                    ("adc_fit", lambda: pa.adc_fit(dwi.data_ptr(), pitch, mask.data_ptr(), nvox, dptr[0], dptr[1], stream=stream), 4 * nvol + 9)):
                for _ in range(3): fn()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(steps): fn()
                b.record(); torch.cuda.synchronize()
                ms = a.elapsed_time(b) / steps
                gbs = nvox * bpv_d / (ms * 1e-3) / 1e9
                dti_legs[name] = {"value": nvox / (ms * 1e-3), "unit": "voxels/s", "ms_per_step": ms, "algorithmic_bytes_per_voxel": bpv_d,
                                  "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak_gbs0, "unit": "GB/s", "frac": gbs / peak_gbs0},
                                  "api": f"fibers_{name}_device on the same {nvol}-volume slab (src/dti.jl:{'221' if name == 'dti_fit' else '164'})"}
            del douts, pd, pa
        except Exception as ex:                     # noqa: BLE001
            dti_legs = {"dti_fit": {"error": f"{type(ex).__name__}: {ex}"[:200]}}

    dti_first = os.environ.get("BENCH_DTI_FIRST", "1") != "0"      # before the tensor-heavy GQI loop (the board's clock governor lags after it)
    if rank == 0 and world == 1 and not args.no_dti and dti_first:
        run_dti_legs()
        torch.cuda.synchronize()
    for _ in range(max(warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank); sampler.start()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = D.launch_count()
    barrier()
    e0.record()
    for i in range(steps):
        step(kev[i])
    e1.record()
    barrier()
    launches = D.launch_count() - l0
    clocks = sampler.stop()
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    dd = dist if world > 1 else None
    # whole-job throughput = voxels of all ranks / max over ranks of the timed region (device clock)
    value = F.batch.whole_job_throughput(steps * nvox, e0.elapsed_time(e1) * 1e-3, dd, dev)
    ms_per_step = world * nvox / value * 1e3

    # One extra, untimed launch with the kernel's own trace switched on: SM cycles (clock64) against wall-clock
    # nanoseconds (globaltimer) over the kernel.  Under tensor load the part runs below the NVML figure sampled above
    # (no throttle reason is raised for it); the JSON line reports both.
    if rank == 0 and plan.kernel == "tc":
        import tempfile
        tf = os.path.join(tempfile.gettempdir(), f"fibers_tc_trace_{os.getpid()}.bin")
        os.environ["FIBERS_TC_TRACE"] = tf
        try:
            step()
            torch.cuda.synchronize()
            raw = np.fromfile(tf, dtype=np.int64)
            cyc = raw[512 + 2 * 127: 512 + 2 * 128]; span = raw[512:514]
            if span[1] > span[0] and cyc[1] > cyc[0]:
                clocks["kernel_effective_sm_mhz"] = round(float(cyc[1] - cyc[0]) / float(span[1] - span[0]) * 1e3, 1)
                clocks["kernel_effective_note"] = "clock64 ticks / globaltimer ns over recon_tc_kernel (one extra untimed launch)"
        except Exception:
            pass
        finally:
            os.environ.pop("FIBERS_TC_TRACE", None)
            if os.path.exists(tf): os.remove(tf)

    if rank == 0 and world == 1 and not args.no_dti and not dti_first:
        run_dti_legs()

    # ---- end to end through the host-pointer C ABI (what the Julia wrapper ccalls) ----------
    e2e = None
    e2e_error = None
    extra = {}
    if not args.no_e2e:
        # Set-up can fail on a box that cannot pin 9 GB of host memory per rank: every rank then agrees (one reduction)
        # to report "e2e": null instead of hanging in a barrier or losing the whole line.
        e2e_step = None
        L = F._lib.lib()
        V = np.asfortranarray(F.sphere_642.vertices); Fc = np.asfortranarray(F.sphere_642.faces)
        bv = np.asfortranarray(bvec)

        def call(s, odf=True):
            F._lib.check(L.fibers_gqi_rec(s.dwi_ptr, 0, s.mask_ptr, shape[0], shape[1], shape[2], nvol, F._lib.ptr(bval), F._lib.ptr(bv),
                                          F._lib.ptr(V), V.shape[0], F._lib.ptr(Fc), Fc.shape[0], 1.25,
                                          s.odf_ptr if odf else None, *s.peak, *[q[0] for q in s.qa], None, 1))
        try:
            barrier()
            ceiling = pcie_ceiling(torch, local_rank)          # N > 1: every rank measures its link at the same time
            subj = HostSubject(torch, shape, nvol, True, (dwi, mask))
            torch.cuda.synchronize()
            dwi_keep = dwi if (world == 1 or rank == 0) else None
            del odf, peak
            if dwi_keep is None:
                del dwi
            torch.cuda.empty_cache()
            e2e_step = lambda: call(subj)
            e2e_step(); e2e_step()                  # warm-up: first-touch of the pinned pages, context cache
            ok = 1.0
        except Exception as ex:                     # noqa: BLE001 - reported in the JSON line
            ok = 0.0
            e2e_error = f"{type(ex).__name__}: {ex}"[:200]
        ok_all = -F.batch.reduce_max(-ok, dd, dev)      # min over ranks
        if ok_all > 0.5:
            barrier()
            per_step = []
            for _ in range(args.e2e_steps):
                t0 = time.perf_counter()
                e2e_step()                          # blocking call: returns when the host buffers hold the results
                per_step.append(time.perf_counter() - t0)
            # host / PCIe side of a shared box is noisy: the median step is reported, the mean is kept beside it
            dt = float(np.median(per_step)); dt_mean = float(np.mean(per_step))
            dt = F.batch.reduce_max(dt, dd, dev)
            ceil_min = -F.batch.reduce_max(-ceiling, dd, dev)          # the slowest rank's link while ALL ranks copy
            floor_ms = max(subj.h2d, subj.d2h) / ceil_min / 1e6
            e2e = {"value": world * nvox / dt, "unit": "voxels/s", "h2d_bytes_per_step": subj.h2d,
                   "d2h_bytes_per_step": subj.d2h, "ms_per_step": dt * 1e3, "ms_per_step_mean": dt_mean * 1e3,
                   "statistic": "median of per-step wall times", "steps": args.e2e_steps,
                   "api": "fibers_gqi_rec (host pointers, pinned buffers)",
                   "pcie_ceiling_GBps_each_way": ceil_min, "ceiling_ms_per_step": floor_ms, "frac_of_ceiling": floor_ms / (dt * 1e3),
                   "ceiling_note": "concurrent H2D + D2H of 512 MiB pinned buffers, measured in this run" +
                                   (f" by all {world} ranks at the same time (the slowest link is quoted; tools/gpu/pcie_probe.py is the long form)" if world > 1 else "")}
            if world == 1:
                try:                                # odf = NULL: peaks + QA only
                    med, _ = timed_calls(lambda: call(subj, odf=False), 1, 3)
                    extra["e2e_no_odf"] = {"value": nvox / med, "unit": "voxels/s", "ms_per_step": med * 1e3, "h2d_bytes_per_step": subj.h2d,
                                           "d2h_bytes_per_step": nvox * 4 * 12, "api": "fibers_gqi_rec(odf = NULL): peaks and QA only (src/stream.jl:76-173 reads nothing else)"}
                except Exception as ex:             # noqa: BLE001
                    extra["e2e_no_odf"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
                try:                                # pageable numpy arrays = what a Julia Array is
                    pg = HostSubject(torch, shape, nvol, False, (dwi_keep, mask))
                    med, mean = timed_calls(lambda: call(pg), 1, 3)
                    extra["e2e_pageable"] = {"value": nvox / med, "unit": "voxels/s", "ms_per_step": med * 1e3, "ms_per_step_mean": mean * 1e3,
                                             "h2d_bytes_per_step": pg.h2d, "d2h_bytes_per_step": pg.d2h,
                                             "api": "fibers_gqi_rec on pageable numpy Fortran arrays (what ccall passes, src/mri.jl:249-255): pinned bounce ring + host copy threads",
                                             "copy_threads": min(int(os.environ.get("FIBERS_CUDA_COPY_THREADS", "16")), len(os.sched_getaffinity(0)))}
                    del pg
                except Exception as ex:             # noqa: BLE001
                    extra["e2e_pageable"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
            # ---- N > 1: the in-library multi-GPU paths, rank 0 alone (the other ranks release their GPUs and wait on the host)
            if world > 1 and not args.no_multi:
                L.fibers_cuda_release_cache()
                del subj
                torch.cuda.empty_cache()
                dist.barrier(group=side)
                if rank == 0:
                    try:
                        extra.update(library_multi_gpu_legs(torch, F, shape, bval, bvec, world, (dwi_keep, mask), nsub=args.subjects))
                    except Exception as ex:         # noqa: BLE001
                        extra["zslab"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
                dist.barrier(group=side)
        elif e2e_error is None:
            e2e_error = "end-to-end set-up failed on another rank"

    if rank == 0:
        peak_gbs, peak_src = measured_peaks()
        # voxels outside the mask read 1 mask byte and have their outputs zero-filled; the others move the full 4N + ...
        fill = float(mask.float().mean().item())
        bpv = algorithmic_bytes_per_voxel(nvol, M_VERT)
        bpv_avg = fill * bpv + (1.0 - fill) * (4 * M_VERT + 36 + 12 + 1)
        abytes = bpv_avg * nvox
        achieved = abytes / (kern_ms * 1e-3) / 1e9
        line = {"metric": "voxels/sec (GQI recon+peaks)", "value": value, "unit": "voxels/s", "n_gpus": world,
                "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "kernel": plan.kernel, "gpu_launches": int(launches), "clocks": clocks, "host_cores": len(os.sched_getaffinity(0)),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                             "frac": achieved / peak_gbs, "traffic": profiled_traffic(plan.kernel, nvox) if fill == 1.0 else None,
                             "traffic_source": "committed ncu --set full capture (profiles/roofline_traffic.json), not measured in this run",
                             "peak_source": peak_src,
                             "kernel_ms": kern_ms, "algorithmic_bytes_per_voxel": bpv_avg, "mask_fill": fill,
                             "algorithmic_tflops": 2.0 * nvol * M_VERT * fill * nvox / (kern_ms * 1e-3) / 1e12}}
        if e2e:
            line["e2e"] = e2e
        elif e2e_error:
            line["e2e"] = None
            line["e2e_error"] = e2e_error
        line.update(dti_legs)
        line.update(extra)
        if not args.no_cpu and world == 1:              # reported baseline: rank 0 at N = 1 only
            vps, ms, cores, sample = run_cpu_reference(shape, 2, 1, budget_s=12.0)
            line["cpu_baseline"] = {"value": vps, "unit": "voxels/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.barrier(group=side)
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
