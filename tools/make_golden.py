#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ from the float64 oracle.

The reference ships no golden vectors (test/runtests.jl is empty) and cannot be run here (no
Julia), so these fixtures pin the ORACLE (oracle/fibers_oracle.py, float64 "truth" variant) on
small seeded phantoms; tests/test_oracle.py re-derives them and tests/test_gpu_golden.py checks
the CUDA path against them.  Re-run only when the oracle's semantics are deliberately changed:
    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import fibers_oracle as O                      # noqa: E402
from fibers_jl_b200 import phantom             # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    v, f = O.load_sphere(642)
    # DTI (cfg1-shaped, scaled down): includes partial-sample and zero voxels
    ph = phantom.dti_phantom((10, 9, 6), seed=101)
    r = O.dti_fit(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], np.float64)
    a, s = O.adc_fit(ph["dwi"], ph["mask"], ph["bval"], np.float64)
    np.savez_compressed(os.path.join(OUT, "dti_small.npz"), dwi=ph["dwi"], mask=ph["mask"], bval=ph["bval"],
                        bvec=ph["bvec"], adc=a, adc_s0=s,
                        **{k: r[k] for k in ("s0", "eigval1", "eigval2", "eigval3", "eigvec1", "rd", "md", "fa", "valid", "kind")})
    # GQI (cfg2 protocol: 18 b0 + 3 x 90)
    ph = phantom.gqi_phantom((6, 5, 4), seed=102, mask_fill=0.8)
    r = O.gqi_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, f, 1.25, np.float64)
    np.savez_compressed(os.path.join(OUT, "gqi_small.npz"), dwi=ph["dwi"], mask=ph["mask"], bval=ph["bval"],
                        bvec=ph["bvec"], odf=r["odf"].astype(np.float32), peak_idx=r["peak_idx"],
                        qa=np.stack(r["qa"]).astype(np.float32), odfmax=np.float64(r["odfmax"]), computed=r["computed"])
    # DSI (cfg3 protocol: 515-point grid)
    ph = phantom.dsi_phantom((5, 4, 3), seed=103)
    r = O.dsi_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, f, 32, np.float64)
    np.savez_compressed(os.path.join(OUT, "dsi_small.npz"), dwi=ph["dwi"], mask=ph["mask"], bval=ph["bval"],
                        bvec=ph["bvec"], odf=r["odf"].astype(np.float32), pdf=r["pdf"].astype(np.float32),
                        peak_idx=r["peak_idx"], qa=np.stack(r["qa"]).astype(np.float32), computed=r["computed"])
    # find_peaks! known-answer vectors (hand-built edge cases, evaluated with the LITERAL face-based rule)
    M = 321
    ff = O.fold_faces(f, M)
    nbr = O.neighbour_table(ff, M)
    rng = np.random.default_rng(7)
    cases = []
    o = np.full(M, 1.0, np.float32); cases.append(o)                                   # global plateau: no peaks
    o = np.zeros(M, np.float32); o[10] = 5; cases.append(o)                            # single peak
    o = np.zeros(M, np.float32); o[10] = 5; o[nbr[10, 0]] = 5; cases.append(o)         # adjacent tie kills both
    o = -np.ones(M, np.float32); o[7] = -0.5; cases.append(o)                          # negative local max never counts
    o = np.zeros(M, np.float32); o[[3, 200, 310]] = 2.0; cases.append(o)               # equal non-adjacent peaks: index order
    o = np.zeros(M, np.float32); o[[5, 100, 150, 250]] = [1, 4, 2, 3]; cases.append(o)  # > 3 peaks: top 3 by value
    for _ in range(26):
        cases.append((np.round(rng.normal(size=M) * 3) / 2).astype(np.float32))        # tie-rich, partly negative
    cases = np.stack(cases)
    exp = []
    for o in cases:
        isort, nvalid = O.find_peaks_literal(o, ff)
        exp.append([isort[k] if k < min(nvalid, 3) else -1 for k in range(3)] + [nvalid])
    np.savez_compressed(os.path.join(OUT, "peaks_kat.npz"), odf=cases, expected=np.asarray(exp, np.int32))
    make_stream()
    make_structens()
    make_rumba()
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


def make_stream():
    """stream (src/stream.jl:730): two smooth orientation fields with holes, amplitudes, a mask; nsub = 0 (no random sub-voxel
    offsets: the call is deterministic in the reference as well), otherwise the reference's defaults."""
    import stream_oracle as SO
    g = np.random.default_rng(104)
    shape = (14, 12, 6)
    xs, ys, zs = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
    vols, fs = [], []
    for i in range(2):
        ph = g.uniform(0, 2 * np.pi, 6)
        th = 0.6 * np.sin(xs / 5.0 + ph[0]) + 0.5 * np.cos(ys / 4.0 + ph[1]) + 0.3 * np.sin(zs / 3.0 + ph[2]) + i * 1.1
        el = 0.5 * np.sin(xs / 6.0 + ph[3]) * np.cos(ys / 7.0 + ph[4]) + 0.2 * np.sin(zs / 2.5 + ph[5])
        v = np.stack([np.cos(th) * np.cos(el), np.sin(th) * np.cos(el), np.sin(el)], axis=-1)
        v /= np.linalg.norm(v, axis=-1, keepdims=True)
        v[g.random(shape) < 0.05 + 0.1 * i] = 0
        vols.append(v.astype(np.float32))
        fs.append(g.uniform(0.0, 0.3, shape).astype(np.float32))
    mask = (g.random(shape) < 0.9).astype(np.uint8)
    lines = SO.stream(vols, [np.zeros(3, np.float32)], f=fs, f_thresh=0.05, mask=mask)
    np.savez_compressed(os.path.join(OUT, "stream_small.npz"), ovec=np.stack(vols), f=np.stack(fs), mask=mask, f_thresh=np.float32(0.05),
                        npts=np.array([s.shape[1] for s in lines], np.int32), xyz=np.concatenate(lines, axis=1).astype(np.float32))
    print("stream_small:", len(lines), "streamlines,", sum(s.shape[1] for s in lines), "points")


def make_structens():
    """st_recon (src/structens.jl:40-88): two crossing families of planes + noise, sigma = 1, rho = 2."""
    import structens_oracle as S
    g = np.random.default_rng(105)
    shape = (14, 12, 10)
    x, y, z = np.meshgrid(*[np.arange(n, dtype=np.float64) for n in shape], indexing="ij")

    def planes(normal, period):
        n = np.asarray(normal, np.float64); n /= np.linalg.norm(n)
        return np.sin(2 * np.pi * (x * n[0] + y * n[1] + z * n[2]) / period)
    vol = (planes((1.0, 0.3, 0.2), 6.0) + 0.5 * planes((-0.2, 1.0, 0.5), 9.0) + 0.05 * g.normal(size=shape)).astype(np.float32)
    evec, evals = S.st_recon(vol, 1.0, 2.0)
    np.savez_compressed(os.path.join(OUT, "structens_small.npz"), vol=vol, sigma=np.float64(1.0), rho=np.float64(2.0),
                        eigvec=evec.astype(np.float32), eigval=evals.astype(np.float64))
    print("structens_small:", vol.shape)


def make_rumba():
    """rumba_rec (src/rusd.jl:419-636) on sphere_362 with 40 iterations (positional arguments 3 and 4), otherwise the defaults."""
    import rumba_oracle as R
    ph = phantom.gqi_phantom((6, 5, 4), nb0=3, shells=((2000.0, 45),), seed=106, mask_fill=0.8)
    v, _ = O.load_sphere(362)
    r = R.rumba_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, niter=40, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "rumba_small.npz"), dwi=ph["dwi"], mask=ph["mask"], bval=ph["bval"], bvec=ph["bvec"],
                        niter=np.int32(40), fodf=r["fodf"].astype(np.float32), fgm=r["fgm"], fcsf=r["fcsf"], gfa=r["gfa"], var=r["var"],
                        snr_mean=np.float64(r["snr_mean"]), snr_std=np.float64(r["snr_std"]), peak_idx=r["peak_idx"],
                        peak=np.stack(r["peak"]).astype(np.float32))
    print("rumba_small:", ph["dwi"].shape, "snr", r["snr_mean"], r["snr_std"])


if __name__ == "__main__":
    main()
