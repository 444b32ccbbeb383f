"""GPU parity tests: the CUDA path, called through the C ABI (libfibers_cuda.so via the host
mirror of the reference API), against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

import fibers_oracle as O
import parity as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import fibers_jl_b200 as F
    assert F.device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"
    return F


def _mri(F, ph):
    return F.MRI(ph["dwi"], ph["bval"], ph["bvec"]), F.MRI(ph["mask"])


# ---------------------------------------------------------------- DTI / ADC
@pytest.mark.parametrize("shape,ndir,b", [((24, 24, 10), 30, 1000.0), ((20, 16, 8), 64, 2500.0)])
def test_dti_parity(F, shape, ndir, b):
    from fibers_jl_b200 import phantom
    ph = phantom.dti_phantom(shape, ndir=ndir, b=b, seed=11)
    dwi, mask = _mri(F, ph)
    got = F.dti_fit(dwi, mask)
    r32 = O.dti_fit(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], np.float32)
    r64 = O.dti_fit(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], np.float64)
    # valid-voxel set: bit exact (pure integer logic on s > 0)
    assert np.array_equal(got.valid, r32["valid"])
    assert set(np.unique(r32["kind"])) == {0, 1, 2}, "phantom must exercise full, partial and zero branches"
    v = P.flat(r32["valid"])
    full = P.flat(r32["kind"]) == 1
    # untouched voxels exactly zero
    for name in ("s0", "eigval1", "rd", "md", "fa"):
        assert np.all(P.flat(getattr(got, name).vol)[~v] == 0)
    for name in ("s0", "eigval1", "md", "fa"):
        assert P.rel_err(getattr(got, name).vol, r64[name], full) < P.SCALAR_TOL, name
    # lambda2/3, rd: relative to lambda1 scale where nearly degenerate (conditioning-aware)
    l1 = np.abs(P.flat(r64["eigval1"]))[full]
    for name in ("eigval2", "eigval3", "rd"):
        d = np.abs(P.flat(getattr(got, name).vol)[full].astype(np.float64) - P.flat(r64[name])[full]) / l1
        assert d.max() < P.SCALAR_TOL, name
    # V1: |dot| >= 0.9999 where lambda1 is separated from lambda2
    gap = (P.flat(r64["eigval1"]) - P.flat(r64["eigval2"]))[full] / l1
    dots = np.abs((P.flat(got.eigvec1.vol, 3)[full].astype(np.float64) * P.flat(r64["eigvec1"], 3)[full]).sum(axis=1))
    sep = gap > 2e-2
    assert sep.mean() > 0.9
    assert dots[sep].min() >= P.V1_DOT
    # eigenvectors orthonormal
    V = np.stack([P.flat(getattr(got, f"eigvec{k}").vol, 3)[full] for k in (1, 2, 3)], axis=2).astype(np.float64)
    G = np.einsum("nik,nil->nkl", V, V)
    assert np.abs(G - np.eye(3)).max() < 1e-4
    # partial-sample branch (per-voxel pinv): looser conditioning, compare with the float64 oracle
    part = P.flat(r32["kind"]) == 2
    assert part.sum() > 0
    assert P.rel_err(got.md.vol, r64["md"], part) < 1e-3
    assert P.rel_err(got.fa.vol, r64["fa"], part) < 1e-3


def test_adc_parity(F):
    from fibers_jl_b200 import phantom
    ph = phantom.dti_phantom((20, 20, 8), ndir=30, seed=5)
    dwi, mask = _mri(F, ph)
    adc, s0 = F.adc_fit(dwi, mask)
    a64, s64 = O.adc_fit(ph["dwi"], ph["mask"], ph["bval"], np.float64)
    nz = P.flat(a64) != 0
    assert np.array_equal(P.flat(adc.vol) != 0, nz)
    assert P.rel_err(adc.vol, a64, nz) < P.SCALAR_TOL
    assert P.rel_err(s0.vol, s64, nz) < P.SCALAR_TOL


def test_dti_noise_free_kat(F):
    """Known answer: noise-free single tensor -> exact eigenvalues / V1."""
    from fibers_jl_b200 import phantom
    ph = phantom.dti_phantom((16, 16, 8), snr=0, inject=False, seed=3)
    ph["mask"][:] = 1
    dwi, mask = _mri(F, ph)
    got = F.dti_fit(dwi, mask)
    assert np.abs(P.flat(got.eigval1.vol) - ph["l1"]).max() / 1e-3 < 2e-4
    dots = np.abs((P.flat(got.eigvec1.vol, 3) * ph["e1"]).sum(axis=1))
    assert dots.min() > 0.9999
    l = np.stack([ph["l1"], ph["l2"], ph["l3"]], 1); md = l.mean(1)
    fa = np.sqrt(1.5 * ((l - md[:, None]) ** 2).sum(1) / (l ** 2).sum(1))
    assert np.abs(P.flat(got.fa.vol) - fa).max() < 2e-4


# ---------------------------------------------------------------- GQI
def _check_recon(got, r32, r64, vertices, faces, M, what):
    nbr = O.neighbour_table(O.fold_faces(faces, M), M)
    assert P.odf_rel_err(got.odf.vol, r64["odf"]) < P.ODF_TOL, what
    nbad, nun = P.peak_mismatch_report(got.peak_idx, r64, nbr)
    nvox = r64["computed"].size
    assert nun == 0, f"{what}: {nun} unexplained peak mismatches ({nbad} total of {nvox})"
    assert nbad <= max(2, 2e-3 * nvox), f"{what}: too many tie-explained mismatches: {nbad}"
    # peak vectors are verbatim vertex rows of the reported indices
    idx = P.flat(got.peak_idx, 3)
    for k in range(3):
        pk = P.flat(got.peak[k].vol, 3)
        ok = idx[:, k] >= 0
        assert np.array_equal(pk[ok], vertices[idx[ok, k]])
        assert np.all(pk[~ok] == 0)
    # QA: where indices agree, within 1e-4 of the voxel's max ODF / odfmax
    same = (idx == P.flat(r64["peak_idx"], 3)).all(axis=1)
    scale = np.abs(P.flat(r64["odf"], M)).max(axis=1) / float(r64["odfmax"])
    for k in range(3):
        d = np.abs(P.flat(got.qa[k].vol).astype(np.float64) - P.flat(r64["qa"][k]))[same]
        assert np.all(d <= 2 * P.ODF_TOL * scale[same] + 1e-12), what
    return nbad


@pytest.fixture(params=["tc", "simt"])
def kernel(request, F):
    """Runs the test once per reconstruction kernel (both are CUDA paths behind the same ABI)."""
    F.device.set_kernel(request.param)
    yield request.param
    F.device.set_kernel("auto")


def test_auto_selects_tensor_core_kernel(F):
    from fibers_jl_b200 import phantom
    bval, bvec = phantom.shells_table(18, [(1000.0, 90), (2000.0, 90), (3000.0, 90)])
    F.device.set_kernel("auto")
    assert F.device.Plan("gqi", 0, bval, bvec).kernel == "tc"
    assert F.device.Plan("gqi", 0, bval, bvec, F.sphere_362).kernel == "tc"
    assert F.device.Plan("gqi", 0, bval, bvec, F.sphere_724).kernel == "tc"       # 362 half-sphere vertices: 2-stage B ring
    bq, gq = phantom.dsi_grid_table()
    assert F.device.Plan("dsi", 0, bq, gq).kernel == "tc"


@pytest.mark.parametrize("nsphere", [642, 362, 724])
def test_gqi_parity(F, nsphere, kernel):
    from fibers_jl_b200 import phantom
    v, f = O.load_sphere(nsphere)
    odf_dirs = F.ODF(v, f)
    ph = phantom.gqi_phantom((24, 20, 12) if nsphere == 642 else (12, 10, 6), seed=2, mask_fill=0.6)
    dwi, mask = _mri(F, ph)
    got = F.gqi_rec(dwi, mask, odf_dirs)
    r32 = O.gqi_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, f, 1.25, np.float32)
    r64 = O.gqi_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, f, 1.25, np.float64)
    _check_recon(got, r32, r64, v, f, v.shape[0] // 2, f"gqi sphere_{nsphere}")
    assert (~r64["computed"]).sum() > 0, "phantom must contain skipped voxels"


def test_gqi_int16_input_and_sigma(F, sphere642, kernel):
    from fibers_jl_b200 import phantom
    v, f = sphere642
    ph = phantom.gqi_phantom((10, 8, 6), seed=9)
    ph["dwi"] = np.asfortranarray(np.round(ph["dwi"]).astype(np.int16))
    dwi, mask = _mri(F, ph)
    got = F.gqi_rec(dwi, mask, F.sphere_642, 1.0)
    r64 = O.gqi_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, f, 1.0, np.float64)
    _check_recon(got, None, r64, v, f, 321, "gqi int16")


def test_gqi_tc_overflow_fixup(F, sphere642):
    """A few voxels 1000x brighter than the sampled maximum overflow the scaled fp16 operand; the
    tensor-core kernel must detect them and the SIMT fix-up must restore fp32-accurate results."""
    from fibers_jl_b200 import phantom
    v, f = sphere642
    ph = phantom.gqi_phantom((40, 33, 9), seed=31)          # 11880 voxels; the sample sees runs of 32 every 2048
    dwi = ph["dwi"]
    flat = dwi.reshape(-1, dwi.shape[3], order="F")
    hot = [700, 5000, 9001]                                  # none of them inside a sampled run
    for h in hot:
        assert h % 2048 >= 32
        flat[h] *= 3000.0
    ph["dwi"] = np.asfortranarray(flat.reshape(dwi.shape, order="F"))
    F.device.set_kernel("tc")
    try:
        got = F.gqi_rec(*_mri(F, ph))
    finally:
        F.device.set_kernel("auto")
    r64 = O.gqi_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, f, 1.25, np.float64)
    assert np.isfinite(got.odf.vol).all()
    _check_recon(got, None, r64, v, f, 321, "gqi tc overflow fix-up")


def test_gqi_multichunk_pipeline(F, sphere642, kernel):
    """> 2^18 voxels: exercises the slab pipeline (several chunks, stream ring) and odfmax reduce."""
    from fibers_jl_b200 import phantom
    v, f = sphere642
    ph = phantom.gqi_phantom((80, 80, 48), nb0=1, shells=((1500.0, 20),), seed=4, mask_fill=0.5)
    dwi, mask = _mri(F, ph)
    got = F.gqi_rec(dwi, mask, F.sphere_642)
    r64 = O.gqi_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, f, 1.25, np.float64)
    _check_recon(got, None, r64, v, f, 321, "gqi multichunk")


def test_multi_gpu_zslab_sharding(F, sphere642):
    """ngpu = 2 inside one process (what the Julia ccall uses): z-slabs on two devices, host gather,
    host-side odfmax reduce.  Results must be identical to the single-GPU call."""
    if F.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    from fibers_jl_b200 import phantom
    v, f = sphere642
    ph = phantom.gqi_phantom((20, 18, 11), seed=17, mask_fill=0.6)
    F.device.set_devices([0, 1])
    try:
        g1 = F.gqi_rec(*_mri(F, ph), ngpu=1)
        g2 = F.gqi_rec(*_mri(F, ph), ngpu=2)
        assert np.array_equal(g1.odf.vol, g2.odf.vol) and np.array_equal(g1.peak_idx, g2.peak_idx)
        for k in range(3):
            assert np.array_equal(g1.qa[k].vol, g2.qa[k].vol, equal_nan=True)
        r64 = O.gqi_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, f, 1.25, np.float64)
        _check_recon(g2, None, r64, v, f, 321, "gqi 2 gpus")
        pt = phantom.dti_phantom((16, 14, 9), seed=18)
        d1 = F.dti_fit(*_mri(F, pt), ngpu=1); d2 = F.dti_fit(*_mri(F, pt), ngpu=2)
        assert np.array_equal(d1.fa.vol, d2.fa.vol, equal_nan=True) and np.array_equal(d1.valid, d2.valid)
        # the fused DTI + GQI call shards the same way
        fd, fg = F.dti_gqi_fit(*_mri(F, ph), ngpu=2)
        assert np.array_equal(fg.odf.vol, g1.odf.vol)
        for k in range(3):
            assert np.array_equal(fg.qa[k].vol, g1.qa[k].vol, equal_nan=True) and np.array_equal(fg.peak[k].vol, g1.peak[k].vol)
        dref = F.dti_fit(*_mri(F, ph), ngpu=1)
        assert np.array_equal(fd.fa.vol, dref.fa.vol, equal_nan=True) and np.array_equal(fd.eigvec1.vol, dref.eigvec1.vol, equal_nan=True)
    finally:
        F.device.set_devices([0])


def test_gqi_edge_cases(F, sphere642, kernel):
    v, f = sphere642
    from fibers_jl_b200 import phantom
    bval, bvec = phantom.shells_table(1, [(2000.0, 15)])
    # all-zero mask, all-non-positive data, single voxel, ragged (non multiple of 64) sizes
    for shape in [(1, 1, 1), (5, 3, 2), (7, 9, 3)]:
        rng = np.random.default_rng(0)
        dwi = np.asfortranarray(rng.uniform(10, 100, shape + (16,)).astype(np.float32))
        mask = np.ones(shape, np.uint8, order="F")
        if shape != (1, 1, 1):
            mask[0, 0, 0] = 0
            dwi[-1, -1, -1, :] = -1
        got = F.gqi_rec(F.MRI(dwi, bval, bvec), F.MRI(mask))
        r64 = O.gqi_rec(dwi, mask, bval, bvec, v, f, 1.25, np.float64)
        _check_recon(got, None, r64, v, f, 321, f"gqi edge {shape}")
    # nothing computed: odfmax = 0 -> QA = 0/0 = NaN in the reference; ODF and peaks stay 0
    dwi = np.zeros((4, 4, 2, 16), np.float32, order="F")
    got = F.gqi_rec(F.MRI(dwi, bval, bvec), F.MRI(np.ones((4, 4, 2), np.uint8)))
    assert np.all(got.odf.vol == 0) and np.all(got.peak_idx == -1)
    assert np.all(np.isnan(got.qa[0].vol))


def test_missing_tables_raise(F):
    dwi = F.MRI(np.zeros((2, 2, 2, 4), np.float32))
    with pytest.raises(RuntimeError, match="Missing b-value table"):
        F.gqi_rec(dwi, F.MRI(np.ones((2, 2, 2), np.uint8)))
    dwi = F.MRI(np.zeros((2, 2, 2, 4), np.float32), bval=np.ones(4, np.float32))
    with pytest.raises(RuntimeError, match="Missing gradient table"):
        F.dti_fit(dwi, F.MRI(np.ones((2, 2, 2), np.uint8)))


# ---------------------------------------------------------------- DSI
def test_dsi_parity(F, sphere642, kernel):
    from fibers_jl_b200 import phantom
    v, f = sphere642
    ph = phantom.dsi_phantom((12, 10, 6), seed=3, mask_fill=0.7)
    dwi, mask = _mri(F, ph)
    got = F.dsi_rec(dwi, mask)
    r64 = O.dsi_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, f, 32, np.float64)
    _check_recon(got, None, r64, v, f, 321, "dsi")
    assert P.odf_rel_err(got.pdf.vol, r64["pdf"]) < P.ODF_TOL
    # hann_width = 0 (no window)
    got0 = F.dsi_rec(dwi, mask, F.sphere_642, 0)
    r0 = O.dsi_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, f, 0, np.float64)
    _check_recon(got0, None, r0, v, f, 321, "dsi hann0")
    assert P.odf_rel_err(got0.pdf.vol, r0["pdf"]) < P.ODF_TOL


def test_dsi_tc_overflow_fixup_and_uint16(F, sphere642):
    """DSI on the tensor-core kernel: voxels far brighter than the sampled maximum overflow the scaled fp16
    operand in every pass; the epilogue of the ODF pass lists their tiles and the SIMT kernel recomputes ODF,
    peaks AND the pdf rows.  Also exercises a non-float32 input type."""
    from fibers_jl_b200 import phantom
    v, f = sphere642
    ph = phantom.dsi_phantom((30, 22, 9), seed=41)            # 5940 voxels
    dwi = ph["dwi"]
    flat = dwi.reshape(-1, dwi.shape[3], order="F")
    for h in (300, 2500, 4200):
        assert h % 2048 >= 32
        flat[h] *= 40.0                                        # 1500 * 40 = 6e4 < 65535 keeps uint16 exact
    ph["dwi"] = np.asfortranarray(np.clip(np.round(flat), 0, 65535).astype(np.uint16).reshape(dwi.shape, order="F"))
    F.device.set_kernel("tc")
    try:
        got = F.dsi_rec(*_mri(F, ph))
    finally:
        F.device.set_kernel("auto")
    r64 = O.dsi_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, f, 32, np.float64)
    assert np.isfinite(got.odf.vol).all() and np.isfinite(got.pdf.vol).all()
    _check_recon(got, None, r64, v, f, 321, "dsi tc overflow fix-up")
    assert P.odf_rel_err(got.pdf.vol, r64["pdf"]) < P.ODF_TOL


def test_gqi_tc_candidate_list_overflow_falls_back(F, sphere642, monkeypatch):
    """A 128-voxel tile that lists more (voxel, vertex) pairs than the shared-memory list holds is handed to the
    fp32 fix-up kernel.  The capacity is shrunk to 8 entries so that every tile takes that route."""
    from fibers_jl_b200 import phantom
    v, f = sphere642
    ph = phantom.gqi_phantom((20, 16, 9), seed=23)
    r64 = O.gqi_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, f, 1.25, np.float64)
    monkeypatch.setenv("FIBERS_TC_CAND_CAP", "8")
    F.device.set_kernel("tc")
    try:
        got = F.gqi_rec(*_mri(F, ph))
    finally:
        F.device.set_kernel("auto")
    _check_recon(got, None, r64, v, f, 321, "gqi tc candidate-list overflow")


def test_fused_dti_gqi_equals_the_two_calls(F, sphere642):
    """fibers_dti_gqi_fit: one pass over the DWI volume, results bit-identical to dti_fit
    followed by gqi_rec.  Several chunks (FIBERS_CUDA_CHUNK_VOXELS), ragged mask, non-positive samples."""
    import os
    from fibers_jl_b200 import phantom
    ph = phantom.gqi_phantom((26, 22, 14), seed=31, mask_fill=0.6)     # 8008 voxels, 288 volumes
    os.environ["FIBERS_CUDA_CHUNK_VOXELS"] = "4096"
    try:
        F._lib.lib().fibers_cuda_release_cache()
        d1 = F.dti_fit(*_mri(F, ph))
        g1 = F.gqi_rec(*_mri(F, ph))
        d2, g2 = F.dti_gqi_fit(*_mri(F, ph))
    finally:
        del os.environ["FIBERS_CUDA_CHUNK_VOXELS"]
        F._lib.lib().fibers_cuda_release_cache()
    for name in ("s0", "eigval1", "eigval2", "eigval3", "eigvec1", "eigvec2", "eigvec3", "rd", "md", "fa"):
        a, b = getattr(d1, name).vol, getattr(d2, name).vol
        assert np.array_equal(a, b, equal_nan=True), name
    assert np.array_equal(g1.odf.vol, g2.odf.vol)
    for k in range(3):
        assert np.array_equal(g1.peak[k].vol, g2.peak[k].vol) and np.array_equal(g1.qa[k].vol, g2.qa[k].vol)
    assert np.count_nonzero(d2.fa.vol) > 1000 and np.count_nonzero(g2.qa[0].vol) > 1000


def test_two_live_plans_with_different_spheres_interleaved_and_concurrent(F):
    """Two tensor-core plans with different meshes live on ONE device: the neighbour table travels with every
    launch (kernel parameter) and the shared-memory limit of the kernel is never lowered, so alternating launches
    and two host threads reconstructing at the same time (the library is re-entrant, include/fibers_cuda.h) give
    the same bits as each call alone."""
    import threading
    from fibers_jl_b200 import phantom
    ph = phantom.gqi_phantom((24, 20, 12), seed=52, mask_fill=0.7)
    spheres = {n: F.ODF(*O.load_sphere(n)) for n in (642, 362, 724)}
    F.device.set_kernel("tc")
    try:
        alone = {n: F.gqi_rec(*_mri(F, ph), s) for n, s in spheres.items()}
        # creating the plan with the smallest tile LAST must not break the larger ones (642 needs ~215 KB, 362 ~145 KB)
        for n in (642, 362, 642, 724, 362, 724, 642):
            g = F.gqi_rec(*_mri(F, ph), spheres[n])
            assert np.array_equal(g.odf.vol, alone[n].odf.vol) and np.array_equal(g.peak_idx, alone[n].peak_idx)
        # device-resident plans kept alive side by side
        plans = [F.device.Plan("gqi", 0, ph["bval"], ph["bvec"], spheres[n]) for n in (642, 362, 724)]
        assert all(p.kernel == "tc" for p in plans)
        out, errs = {}, []

        def work(n, reps):
            try:
                for _ in range(reps):
                    out[n] = F.gqi_rec(*_mri(F, ph), spheres[n])
            except Exception as e:          # noqa: BLE001
                errs.append(e)
        for _ in range(3):
            th = [threading.Thread(target=work, args=(n, 4)) for n in (642, 724)]
            [t.start() for t in th]; [t.join() for t in th]
            assert not errs, errs
            for n in (642, 724):
                assert np.array_equal(out[n].odf.vol, alone[n].odf.vol)
                assert np.array_equal(out[n].peak_idx, alone[n].peak_idx), f"sphere_{n}: peaks differ under concurrency"
                for k in range(3):
                    assert np.array_equal(out[n].qa[k].vol, alone[n].qa[k].vol, equal_nan=True)
        del plans
    finally:
        F.device.set_kernel("auto")


def _same_recon(a, b, odf=True):
    if odf:
        assert np.array_equal(a.odf.vol, b.odf.vol)
    for k in range(3):
        assert np.array_equal(a.peak[k].vol, b.peak[k].vol) and np.array_equal(a.qa[k].vol, b.qa[k].vol, equal_nan=True)


def test_pageable_bounce_ring_equals_direct_copies(F, sphere642, monkeypatch):
    """Host transfer paths: the pinned bounce ring (what pageable caller arrays get) and the direct pitched DMA
    give the same bits; several chunks per slot so that the ring wraps, all four entry points, non-float input."""
    from fibers_jl_b200 import phantom
    ph = phantom.gqi_phantom((26, 22, 14), seed=61, mask_fill=0.6)     # 8008 voxels
    pd = phantom.dsi_phantom((14, 12, 7), seed=62, mask_fill=0.8)
    pd["dwi"] = np.asfortranarray(np.round(pd["dwi"]).astype(np.int16))
    monkeypatch.setenv("FIBERS_CUDA_CHUNK_VOXELS", "4096")
    res = {}
    for path in ("direct", "bounce"):
        monkeypatch.setenv("FIBERS_CUDA_HOST_PATH", path)
        F._lib.lib().fibers_cuda_release_cache()
        res[path] = (F.gqi_rec(*_mri(F, ph)), F.dti_fit(*_mri(F, ph)), F.adc_fit(*_mri(F, ph)), F.dsi_rec(*_mri(F, pd)),
                     F.dti_gqi_fit(*_mri(F, ph)))
    F._lib.lib().fibers_cuda_release_cache()
    g0, d0, a0, s0, f0 = res["direct"]; g1, d1, a1, s1, f1 = res["bounce"]
    _same_recon(g0, g1); assert np.array_equal(g0.peak_idx, g1.peak_idx)
    _same_recon(s0, s1); assert np.array_equal(s0.pdf.vol, s1.pdf.vol) and np.array_equal(s0.peak_idx, s1.peak_idx)
    for name in ("s0", "eigval1", "eigvec1", "eigvec3", "rd", "md", "fa"):
        assert np.array_equal(getattr(d0, name).vol, getattr(d1, name).vol, equal_nan=True), name
        assert np.array_equal(getattr(f0[0], name).vol, getattr(f1[0], name).vol, equal_nan=True), name
    assert np.array_equal(d0.valid, d1.valid)
    assert np.array_equal(a0[0].vol, a1[0].vol) and np.array_equal(a0[1].vol, a1[1].vol)
    _same_recon(f0[1], f1[1])
    # and against the oracle, so that "both paths agree" cannot mean "both are wrong"
    v, f = sphere642
    r64 = O.gqi_rec(ph["dwi"], ph["mask"], ph["bval"], ph["bvec"], v, f, 1.25, np.float64)
    _check_recon(g1, None, r64, v, f, 321, "gqi through the bounce ring")


def test_pinned_caller_buffers_take_the_direct_path(F, sphere642):
    """Caller arrays registered with fibers_cuda_host_register (what bench.py's e2e leg passes) and odf = NULL."""
    from fibers_jl_b200 import phantom, _lib
    ph = phantom.gqi_phantom((20, 18, 10), seed=63, mask_fill=0.7)
    ref = F.gqi_rec(*_mri(F, ph))
    L = _lib.lib()
    dwi = np.asfortranarray(ph["dwi"])
    _lib.check(L.fibers_cuda_host_register(_lib.ptr(dwi), dwi.nbytes))
    try:
        got = F.gqi_rec(F.MRI(dwi, ph["bval"], ph["bvec"]), F.MRI(ph["mask"]))
        _same_recon(ref, got)
        noodf = F.gqi_rec(F.MRI(dwi, ph["bval"], ph["bvec"]), F.MRI(ph["mask"]), want_odf=False)
        assert noodf.odf is None
        _same_recon(ref, noodf, odf=False)
        assert np.array_equal(ref.peak_idx, noodf.peak_idx)
    finally:
        _lib.check(L.fibers_cuda_host_unregister(_lib.ptr(dwi)))


def test_batch_of_subjects_equals_single_calls(F, monkeypatch):
    """fibers_dti_gqi_fit_batch (BASELINE cfg4 shape of work: DTI + GQI per subject, one protocol): the stream ring
    keeps rolling across subject boundaries; results must be bit-identical to one fused call per subject.
    Runs on every visible GPU count up to 4 (1 on the single-GPU box; gpurun --gpus 2/4 covers the queue)."""
    from fibers_jl_b200 import phantom
    monkeypatch.setenv("FIBERS_CUDA_CHUNK_VOXELS", "4096")
    phs = [phantom.gqi_phantom((22, 18, 11), seed=70 + i, mask_fill=0.5 + 0.1 * i) for i in range(5)]
    single = [F.dti_gqi_fit(*_mri(F, ph)) for ph in phs]
    ngpus = [n for n in (1, 2, 4) if n <= F.device_count()]
    F.device.set_devices(list(range(max(ngpus))))
    try:
        for ngpu in ngpus:
            for nsub in (5, 1):         # more subjects than GPUs (queue) / fewer (z-slab split with the host-side odfmax reduce)
                got = F.dti_gqi_fit_batch([_mri(F, ph)[0] for ph in phs[:nsub]], [_mri(F, ph)[1] for ph in phs[:nsub]], ngpu=ngpu)
                for (d1, g1), (d2, g2) in zip(single, got):
                    _same_recon(g1, g2)
                    for name in ("s0", "eigval1", "eigval2", "eigval3", "eigvec1", "eigvec2", "eigvec3", "rd", "md", "fa"):
                        assert np.array_equal(getattr(d1, name).vol, getattr(d2, name).vol, equal_nan=True), (ngpu, nsub, name)
        g_only = F.dti_gqi_fit_batch([_mri(F, ph)[0] for ph in phs[:2]], [_mri(F, ph)[1] for ph in phs[:2]], want_dti=False, want_odf=False)
        for (d1, g1), (d2, g2) in zip(single, g_only):
            assert d2 is None and g2.odf is None
            _same_recon(g1, g2, odf=False)
    finally:
        F.device.set_devices([0])
