"""Kernel-only (device-resident) timings of every kernel on the hot path, with roofline fractions.
Runs on the GPU box:  python tools/gpu/bench_kernels.py > gpurun_out/kernels.json
  DTI  cfg1  64x64x40x31         (HBM-bound, 189 B/voxel)
  DTI  cfg4  145x174x145x288     (HBM-bound, 1217 B/voxel)
  ADC  cfg4-shaped
  GQI  cfg2  145x174x145x288     tensor-core and SIMT kernels (2485 B/voxel)
  GQI  cfg5  400x400x38x128      one GPU's z-slab of the 8-GPU ex-vivo config (1845 B/voxel)
  DSI  cfg3  96x96x60x515        SIMT matrix-form kernel (5453 B/voxel; 861 kflop/voxel)
"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
import fibers_jl_b200 as F
from fibers_jl_b200 import device as D, phantom

dev = torch.device("cuda", 0)
HBM = bench.measured_peaks()[0]
out = []


def timeit(fn, steps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def synth(nvox, bval, bvec, seed, pitch=None):
    return bench.synth_dwi_device(torch, nvox, bval, bvec, seed, dev, pitch=pitch)


def report(name, nvox, ms, bytes_per_voxel, flops_per_voxel=0, note=""):
    gbs = nvox * bytes_per_voxel / (ms * 1e-3) / 1e9
    out.append({"kernel": name, "voxels": nvox, "ms": ms, "voxels_per_s": nvox / (ms * 1e-3), "algorithmic_GBps": gbs,
                "frac_of_measured_hbm": gbs / HBM, "algorithmic_TFLOPs": nvox * flops_per_voxel / (ms * 1e-3) / 1e12, "note": note})
    print(json.dumps(out[-1]), flush=True)


def run_dti(shape, bval, bvec, tag):
    nvox = int(np.prod(shape)); N = bval.shape[0]
    dwi = synth(nvox, bval, bvec, 3)
    mask = torch.ones(nvox, dtype=torch.uint8, device=dev)
    outs = [torch.empty((n, nvox), dtype=torch.float32, device=dev) for n in (1, 1, 1, 1, 3, 3, 3, 1, 1, 1)]
    plan = D.Plan("dti", 0, bval, bvec)
    ms = timeit(lambda: plan.dti_fit(dwi.data_ptr(), nvox, mask.data_ptr(), nvox, nvox, [o.data_ptr() for o in outs]))
    report(f"dti_fit {tag}", nvox, ms, 4 * N + 65, 15 * N)
    a = torch.empty(nvox, dtype=torch.float32, device=dev); s0 = torch.empty_like(a)
    plan2 = D.Plan("adc", 0, bval)
    ms = timeit(lambda: plan2.adc_fit(dwi.data_ptr(), nvox, mask.data_ptr(), nvox, a.data_ptr(), s0.data_ptr()))
    report(f"adc_fit {tag}", nvox, ms, 4 * N + 9, 5 * N)


def run_recon(kind, shape, bval, bvec, kernel, tag, odf_dirs=F.sphere_642, aligned=True, dpitch=None, opitch=None, env=None):
    nvox = int(np.prod(shape)); N = bval.shape[0]; M = odf_dirs.nvert
    pitch = opitch or (nvox + 63) // 64 * 64
    dpitch = dpitch or (pitch if aligned else nvox)          # frame pitch of the DWI slab (16-byte aligned rows or not)
    os.environ.update(env or {})
    dwi = synth(nvox, bval, bvec, 5, dpitch)
    mask = torch.ones(nvox, dtype=torch.uint8, device=dev)
    odf = torch.empty((M, pitch), dtype=torch.float32, device=dev)
    pdf = torch.empty((N, pitch), dtype=torch.float32, device=dev) if kind == "dsi" else None
    peak = [torch.empty((3, pitch), dtype=torch.float32, device=dev) for _ in range(3)]
    qa = [torch.empty(pitch, dtype=torch.float32, device=dev) for _ in range(3)]
    stats = torch.zeros(2, dtype=torch.int32, device=dev)
    D.set_kernel(kernel)
    plan = D.Plan(kind, 0, bval, bvec, odf_dirs)
    D.set_kernel("auto")
    fn = lambda: plan.recon(dwi.data_ptr(), dpitch, mask.data_ptr(), nvox, pitch, odf.data_ptr(), [p.data_ptr() for p in peak],
                            [q.data_ptr() for q in qa], stats.data_ptr(), d_pdf=pdf.data_ptr() if pdf is not None else 0, finalize=True)
    ms = timeit(fn, steps=5 if kind == "dsi" else 10)
    for k in (env or {}):
        del os.environ[k]
    if kind == "gqi":
        report(f"gqi_rec {tag} [{plan.kernel}]", nvox, ms, 4 * N + 4 * M + 49, 2 * N * M)
    else:
        report(f"dsi_rec {tag} [{plan.kernel}]", nvox, ms, 8 * N + 4 * M + 49, 2 * N * (M + N), "tensor-bound in matrix form (3 kernel passes: odf rows, 2 x pdf rows)")


ONLY = os.environ.get("BENCH_KERNELS_ONLY", "")          # "dti": the DTI / ADC rows only (A/B of kernel variants); "pitch": the frame-pitch rows
if ONLY == "pitch":
    b2, g2 = bench.make_tables()
    nv = 145 * 174 * 145
    for rep in range(2):
        run_recon("gqi", (145, 174, 145), b2, g2, "tc", "cfg2, rows 16-byte aligned (1 map)")
        run_recon("gqi", (145, 174, 145), b2, g2, "tc", "cfg2, DWI pitch = nvox: rows 8-byte aligned (2 maps)", dpitch=nv)
        run_recon("gqi", (145, 174, 145), b2, g2, "tc", "cfg2, DWI pitch = nvox + 1: rows 4-byte aligned (4 maps)", dpitch=nv + 1)
        run_recon("gqi", (145, 174, 145), b2, g2, "tc", "cfg2, DWI and output pitch = nvox (the reference's own array layout)", dpitch=nv, opitch=nv)
        run_recon("gqi", (145, 174, 145), b2, g2, "tc", "cfg2, DWI pitch = nvox, cp.async staging (FIBERS_TC_NO_SPLIT_TMA)", dpitch=nv, env={"FIBERS_TC_NO_SPLIT_TMA": "1"})
    sys.exit(0)
if ONLY == "dsi":
    b3, g3 = phantom.dsi_grid_table()
    run_recon("dsi", (96, 96, 60), b3, g3, "tc", "cfg3 96x96x60x515")
    sys.exit(0)
b1, g1 = phantom.shells_table(1, [(1000.0, 30)])
run_dti((64, 64, 40), b1, g1, "cfg1 64x64x40x31")
b2, g2 = bench.make_tables()
run_dti((145, 174, 145), b2, g2, "cfg4 145x174x145x288")
if ONLY == "dti":
    sys.exit(0)
run_recon("gqi", (145, 174, 145), b2, g2, "tc", "cfg2 145x174x145x288")
run_recon("gqi", (145, 174, 145), b2, g2, "tc", "cfg2 145x174x145x288, DWI pitch = nvox (rows 8-byte aligned: 2 TMA maps)", aligned=False)
run_recon("gqi", (145, 174, 145), b2, g2, "tc", "cfg2 145x174x145x288, DWI and output pitch = nvox (the reference's own array layout)", dpitch=145 * 174 * 145, opitch=145 * 174 * 145)
run_recon("gqi", (145, 174, 145), b2, g2, "simt", "cfg2 145x174x145x288")
b5, g5 = phantom.shells_table(8, [(4000.0, 120)])
run_recon("gqi", (400, 400, 38), b5, g5, "tc", "cfg5 slab 400x400x38x128 (1/8 of 400x400x300)")
run_recon("gqi", (145, 174, 73), b2, g2, "tc", "cfg2 half volume, sphere_362", F.sphere_362)
run_recon("gqi", (145, 174, 73), b2, g2, "tc", "cfg2 half volume, sphere_724", F.sphere_724)
b3, g3 = phantom.dsi_grid_table()
run_recon("dsi", (96, 96, 60), b3, g3, "tc", "cfg3 96x96x60x515")
run_recon("dsi", (96, 96, 60), b3, g3, "simt", "cfg3 96x96x60x515")
