#!/usr/bin/env python
"""Pinning kit, step 1 of 3: write the inputs of the committed golden fixtures (tests/golden/*.npz) as files the REAL
reference reads with its own `mri_read` / `readdlm`:

    python tools/pin/export_inputs.py pin_inputs
    JULIA_NUM_THREADS=auto julia --project=<Fibers.jl checkout> julia/pin_golden.jl pin_inputs pin_outputs
    python tools/pin/check_outputs.py pin_outputs

Parity of this repository is UNPINNED because Julia is not available where it was built: the
oracle restates the reference, and the golden fixtures are the oracle's own output.  Anyone with Julia and Fibers.jl
closes that gap with the three commands above: step 2 runs the unmodified reference on the fixtures' inputs, step 3
compares its volumes with the oracle's (the same tolerances the GPU parity tests use).

Per reconstruction fixture: <name>_dwi.nii.gz (float32 [nx,ny,nz,nvol]), <name>_mask.nii.gz (float32 0 / 1), <name>_bval.txt (one value
per line), <name>_bvec.txt (three columns).  The tables are written with the shortest decimal that round-trips float32
and under names `mri_read` does NOT pick up by itself (it would re-normalise the gradient vectors, src/mri.jl:705-712):
the Julia script assigns them to `dwi.bval` / `dwi.bvec` as they are.  The tractography fixture is stream_small_ovec<i>.nii.gz,
stream_small_f<i>.nii.gz and stream_small_mask.nii.gz.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import fibers_jl_b200 as Fb                    # noqa: E402  (volume I/O is host code: no GPU needed)

FIXTURES = ("dti_small", "gqi_small", "dsi_small", "rumba_small")


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    M = np.diag([2.0, 2.0, 2.0, 1.0]).astype(np.float32)
    for name in FIXTURES:
        d = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        hdr = dict(vox2ras0=M, volres=[2.0, 2.0, 2.0])
        Fb.mri_write(Fb.MRI(np.asfortranarray(d["dwi"], dtype=np.float32), header=dict(hdr)), os.path.join(outdir, name + "_dwi.nii.gz"))
        Fb.mri_write(Fb.MRI(np.asfortranarray(d["mask"].astype(np.float32)), header=dict(hdr)), os.path.join(outdir, name + "_mask.nii.gz"))
        with open(os.path.join(outdir, name + "_bval.txt"), "w") as f:
            f.write("".join(f"{np.float32(b)!s}\n" for b in d["bval"]))
        with open(os.path.join(outdir, name + "_bvec.txt"), "w") as f:
            f.write("".join(" ".join(str(np.float32(x)) for x in g) + "\n" for g in d["bvec"]))
        print(f"{name}: dwi {d['dwi'].shape}, {int(d['mask'].sum())} mask voxels, {d['bval'].shape[0]} volumes")
    # stream (src/stream.jl:730): two orientation-vector volumes, their amplitudes, a mask; run with nsub = 0 (deterministic)
    d = np.load(os.path.join(ROOT, "tests", "golden", "stream_small.npz"))
    hdr = dict(vox2ras0=M, volres=[2.0, 2.0, 2.0])
    for i in range(d["ovec"].shape[0]):
        Fb.mri_write(Fb.MRI(np.asfortranarray(d["ovec"][i]), header=dict(hdr)), os.path.join(outdir, f"stream_small_ovec{i + 1}.nii.gz"))
        Fb.mri_write(Fb.MRI(np.asfortranarray(d["f"][i]), header=dict(hdr)), os.path.join(outdir, f"stream_small_f{i + 1}.nii.gz"))
    Fb.mri_write(Fb.MRI(np.asfortranarray(d["mask"].astype(np.float32)), header=dict(hdr)), os.path.join(outdir, "stream_small_mask.nii.gz"))
    print(f"stream_small: {d['ovec'].shape[0]} orientation volumes {d['ovec'].shape[1:4]}, {int(d['mask'].sum())} mask voxels")
    # structure tensor (src/structens.jl:40-88): one scalar volume; sigma = 1, rho = 2 in the Julia step
    d = np.load(os.path.join(ROOT, "tests", "golden", "structens_small.npz"))
    Fb.mri_write(Fb.MRI(np.asfortranarray(d["vol"]), header=dict(hdr)), os.path.join(outdir, "structens_small_vol.nii.gz"))
    print(f"structens_small: volume {d['vol'].shape}")
    print("inputs written to", outdir)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "pin_inputs")
