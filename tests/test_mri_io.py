"""Volume I/O: the C-ABI readers / writers against an independent byte-level statement of the two formats
(struct + gzip from the standard library), following src/mri.jl load_nifti_hdr / load_nifti / load_mgh / save_nifti / save_mgh /
mri_write.  Host-only code: no GPU needed."""
import gzip
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import fibers_jl_b200 as Fb  # noqa: E402
from fibers_jl_b200 import io as fio  # noqa: E402

NIFTI_DT = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16, 768: np.uint32}


def py_nifti_bytes(vol, *, endian="<", code=16, sform=None, sform_code=1, qform_code=0, quatern=(0, 0, 0), qoffset=(0, 0, 0),
                   pixdim=(1, 1, 1, 1, 0, 0, 0, 0), units=2 | 16, slope=0.0, inter=0.0, dim=None, glmin=0, trailing=b""):
    """A NIfTI-1 file image built field by field (nifti1.h layout)."""
    e = endian
    shape = list(vol.shape)
    if dim is None:
        dim = [len(shape)] + shape + [1] * (7 - len(shape))
    h = bytearray(352)
    struct.pack_into(e + "i", h, 0, 348)
    struct.pack_into(e + "8h", h, 40, *dim)
    struct.pack_into(e + "h", h, 70, code)
    struct.pack_into(e + "h", h, 72, np.dtype(NIFTI_DT[code]).itemsize * 8)
    struct.pack_into(e + "8f", h, 76, *pixdim)
    struct.pack_into(e + "f", h, 108, 352.0)
    struct.pack_into(e + "f", h, 112, slope)
    struct.pack_into(e + "f", h, 116, inter)
    struct.pack_into("b", h, 123, units)
    struct.pack_into(e + "i", h, 144, glmin)
    struct.pack_into(e + "h", h, 252, qform_code)
    struct.pack_into(e + "h", h, 254, sform_code)
    struct.pack_into(e + "3f", h, 256, *quatern)
    struct.pack_into(e + "3f", h, 268, *qoffset)
    S = np.eye(4, dtype=np.float32) if sform is None else np.asarray(sform, np.float32)
    struct.pack_into(e + "12f", h, 280, *S[:3].reshape(-1))
    h[344:348] = b"n+1\0"
    data = np.asfortranarray(vol).astype(np.dtype(NIFTI_DT[code]).newbyteorder(e)).tobytes(order="F")
    return bytes(h) + data + trailing


def test_nifti_write_layout_and_roundtrip(tmp_path):
    g = np.random.default_rng(0)
    vol = np.asfortranarray(g.standard_normal((7, 5, 4, 3)).astype(np.float32))
    M = np.array([[-1.25, 0, 0, 90], [0, 0, 1.5, -126], [0, -2.0, 0, 72], [0, 0, 0, 1]], np.float32)   # det(Mdc) > 0 after the sign flips? checked below
    mri = Fb.MRI(vol, vox2ras0=M, tr=8800.0, niftihdr=dict(scl_slope=0.0, scl_inter=0.0))
    for name, opener in (("a.nii", open), ("a.nii.gz", gzip.open)):
        path = str(tmp_path / name)
        assert fio.mri_write(mri, path) is False
        raw = opener(path, "rb").read()
        assert len(raw) == 352 + vol.nbytes
        assert struct.unpack_from("<i", raw, 0)[0] == 348
        assert struct.unpack_from("<8h", raw, 40) == (4, 7, 5, 4, 3, 1, 1, 1)
        assert struct.unpack_from("<hh", raw, 70) == (16, 32)
        pix = struct.unpack_from("<8f", raw, 76)
        np.testing.assert_allclose(pix[1:5], [1.25, 2.0, 1.5, 8800.0], rtol=1e-6)
        assert abs(pix[0]) == 1.0
        assert struct.unpack_from("<f", raw, 108)[0] == 352.0
        assert raw[123] == (2 | 16)
        cal_max, cal_min = struct.unpack_from("<ff", raw, 124)
        assert cal_max == vol.max() and cal_min == vol.min()
        assert raw[148:228] == b"FreeSurfer julia".ljust(80)
        assert struct.unpack_from("<hh", raw, 252) == (1, 1)
        np.testing.assert_array_equal(np.array(struct.unpack_from("<12f", raw, 280), np.float32).reshape(3, 4), M[:3])
        assert raw[328:332] == b"huh?" and raw[344:348] == b"n+1\0" and raw[348:352] == b"\0\0\0\0"
        assert raw[352:] == vol.tobytes(order="F")
        # the quaternion must reproduce the rotation part: rebuild the qform the way load_nifti_hdr does
        back = fio.mri_read(path)
        np.testing.assert_array_equal(back.vol, vol)
        np.testing.assert_allclose(back.header["niftihdr"]["qform"], M, atol=2e-5)
        np.testing.assert_array_equal(back.header["vox2ras0"], M)                  # sform wins
        np.testing.assert_allclose(back.header["volres"], [1.25, 2.0, 1.5], rtol=1e-6)
        assert back.header["tr"] == 8800.0 and back.header["nframes"] == 3 and back.header["volsize"] == [7, 5, 4]


def test_nifti_read_big_endian_units_scaling_qform(tmp_path):
    vol = (np.arange(6 * 4 * 3, dtype=np.int16).reshape(6, 4, 3, order="F") - 20)
    b, c, d = 0.0, 0.0, np.sin(np.pi / 8)                         # 45 degrees about z
    raw = py_nifti_bytes(vol, endian=">", code=4, sform_code=0, qform_code=1, quatern=(b, c, d), qoffset=(0.01, -0.02, 0.03),
                         pixdim=(-1, 0.002, 0.003, 0.004, 2.5, 0, 0, 0), units=1 | 8, slope=2.0, inter=10.0)
    p = tmp_path / "be.nii"
    p.write_bytes(raw)
    m = fio.mri_read(str(p))
    assert m.vol.dtype == np.int16 and m.vol.shape == (6, 4, 3, 1)   # dim = (6, 4, 3, 1, 1, 1, 1): folded into a singleton 4th axis
    np.testing.assert_array_equal(m.vol[..., 0], vol * 2 + 10)       # vol .= Int16.(vol .* slope .+ inter)
    assert m.header["niftihdr"]["do_bswap"] is True
    a = np.sqrt(1 - d * d)
    R = np.array([[a * a - d * d, -2 * a * d, 0], [2 * a * d, a * a - d * d, 0], [0, 0, -(a * a + d * d)]])   # qfac = -1 flips the third column
    Q = np.eye(4); Q[:3, :3] = R @ np.diag([2.0, 3.0, 4.0]); Q[:3, 3] = [0.01, -0.02, 0.03]   # metres -> mm for pixdim only (:1470-1473)
    np.testing.assert_allclose(m.header["vox2ras0"], Q, atol=1e-5)
    assert m.header["tr"] == 2500.0                                 # seconds -> ms


def test_nifti_dims_beyond_four_fold_into_frames_and_errors(tmp_path):
    vol = np.arange(2 * 3 * 2 * 2 * 3, dtype=np.float32).reshape((2, 3, 2, 2, 3), order="F")
    raw = py_nifti_bytes(vol, dim=[5, 2, 3, 2, 2, 3, 1, 1])
    p = tmp_path / "five.nii.gz"
    p.write_bytes(gzip.compress(raw))
    m = fio.mri_read(str(p))
    assert m.vol.shape == (2, 3, 2, 6) and m.header["nframes"] == 6
    np.testing.assert_array_equal(m.vol, vol.reshape((2, 3, 2, 6), order="F"))
    bad = bytearray(raw); struct.pack_into("<i", bad, 0, 100)
    (tmp_path / "bad.nii").write_bytes(bytes(bad))
    with pytest.raises(Fb.FibersCudaError, match="Invalid header size"):
        fio.mri_read(str(tmp_path / "bad.nii"))
    bad = bytearray(raw); struct.pack_into("<h", bad, 70, 128)     # RGB24: not supported by the reference either
    (tmp_path / "rgb.nii").write_bytes(bytes(bad))
    with pytest.raises(Fb.FibersCudaError, match="Data type 128 not supported"):
        fio.mri_read(str(tmp_path / "rgb.nii"))
    (tmp_path / "long.nii").write_bytes(raw + b"xx")
    with pytest.raises(Fb.FibersCudaError, match="did not reach end of file"):
        fio.mri_read(str(tmp_path / "long.nii"))
    with pytest.raises(ValueError, match="Cannot determine format"):
        fio.mri_read(str(tmp_path / "volume.img"))


def py_mgh_bytes(vol, M, parms, type_code):
    nd = list(vol.shape) + [1] * (4 - vol.ndim)
    M = np.asarray(M, np.float64)
    delta = np.sqrt((M[:3, :3] ** 2).sum(0))
    Mdc = M[:3, :3] / delta
    pc = (M @ np.array([nd[0] / 2, nd[1] / 2, nd[2] / 2, 1.0]))[:3]
    h = struct.pack(">7i", 1, nd[0], nd[1], nd[2], nd[3], type_code, 1) + struct.pack(">h", 1)
    h += struct.pack(">3f", *delta) + struct.pack(">9f", *Mdc.reshape(-1, order="F")) + struct.pack(">3f", *pc)
    h += b"\0" * (256 - 2 - 60)
    dt = {3: ">f4", 0: "u1", 4: ">i2", 10: ">u2", 1: ">i4"}[type_code]
    return h + np.asfortranarray(vol).astype(dt).tobytes(order="F") + struct.pack(">4f", *parms)


def test_mgh_read_and_write(tmp_path):
    g = np.random.default_rng(1)
    vol = np.asfortranarray(g.integers(-500, 500, (5, 6, 4, 2)).astype(np.int16))
    M = np.array([[-1, 0, 0, 64.5], [0, 0, 2, -40], [0, -1.5, 0, 33.25], [0, 0, 0, 1]], np.float32)
    raw = py_mgh_bytes(vol, M, (2300.0, 0.1396, 2.9, 900.0), 4)
    (tmp_path / "a.mgh").write_bytes(raw)
    (tmp_path / "a.mgz").write_bytes(gzip.compress(raw))
    for name in ("a.mgh", "a.mgz"):
        m = fio.mri_read(str(tmp_path / name))
        np.testing.assert_array_equal(m.vol, vol)
        np.testing.assert_allclose(m.header["vox2ras0"], M, atol=1e-5)
        np.testing.assert_allclose([m.header[k] for k in ("tr", "flip_angle", "te", "ti")], [2300.0, 0.1396, 2.9, 900.0], rtol=1e-6)
    # writer: byte-identical to the independent statement of save_mgh
    mri = Fb.MRI(vol, vox2ras0=M, tr=2300.0, flip_angle=np.float32(0.1396), te=np.float32(2.9), ti=900.0)
    fio.mri_write(mri, str(tmp_path / "w.mgh"))
    assert (tmp_path / "w.mgh").read_bytes() == raw
    fio.mri_write(mri, str(tmp_path / "w.mgz"))
    assert gzip.open(tmp_path / "w.mgz", "rb").read() == raw
    fvol = np.asfortranarray(g.standard_normal((4, 3, 2)).astype(np.float32))
    fio.mri_write(Fb.MRI(fvol, vox2ras0=np.eye(4, dtype=np.float32)), str(tmp_path / "f.mgh"))
    assert (tmp_path / "f.mgh").read_bytes() == py_mgh_bytes(fvol, np.eye(4), (0, 0, 0, 0), 3)
    np.testing.assert_array_equal(fio.mri_read(str(tmp_path / "f.mgh")).vol, fvol)


def test_mri_read_tables_and_datatype_conversion(tmp_path):
    g = np.random.default_rng(2)
    vol = np.asfortranarray(np.round(g.uniform(0, 1000, (4, 4, 3, 5))).astype(np.float32))
    bval = np.array([0, 1000, 1000, 2000, 2000], np.float32)
    bvec = np.array([[0, 0, 0], [2, 0, 0], [0, 3, 4], [1, 1, 1], [0, 0, -0.5]], np.float32)
    mri = Fb.MRI(vol, bval, bvec, vox2ras0=np.diag([2, 2, 2, 1]).astype(np.float32))
    fio.mri_write(mri, str(tmp_path / "dwi.nii.gz"), datatype=np.int16)             # mri_write(mri, outfile, Int16): Int16.(vol)
    back = fio.mri_read(str(tmp_path / "dwi.nii.gz"))
    assert back.vol.dtype == np.int16
    np.testing.assert_array_equal(back.vol, vol.astype(np.int16))
    np.testing.assert_array_equal(back.bval, bval)
    with np.errstate(invalid="ignore"):
        want = bvec / np.sqrt((bvec ** 2).sum(1, keepdims=True))
    want[0] = 0          # normalised rows, b0 row NaN -> 0 (:701-703)
    np.testing.assert_allclose(back.bvec, want, rtol=1e-6)
    with pytest.raises(Fb.FibersCudaError, match="InexactError"):                  # Int16.(vol) raises for non-integers
        fio.mri_write(Fb.MRI(vol + np.float32(0.5), vox2ras0=np.eye(4, dtype=np.float32)), str(tmp_path / "x.nii"), datatype=np.int16)
    # tables given as rows / in the other order
    (tmp_path / "r.bvals").write_text(" ".join(str(x) for x in bval) + "\n")
    (tmp_path / "r.bvecs").write_text("\n".join(" ".join(str(x) for x in bvec[:, c]) for c in range(3)) + "\n")
    b, gtab = fio.mri_read_bfiles(str(tmp_path / "r.bvecs"), str(tmp_path / "r.bvals"))
    np.testing.assert_array_equal(b, bval); np.testing.assert_array_equal(gtab, bvec)
    # file stem without extension; permutedata
    m2 = fio.mri_read(str(tmp_path / "dwi"), permutedata=True)
    assert m2.vol.shape == (4, 4, 3, 5) and m2.header["ispermuted"]
    np.testing.assert_array_equal(m2.vol, np.swapaxes(back.vol, 0, 1))


def test_trk_write_layout(tmp_path):
    """fibers_trk_write against the TrackVis v2 layout trk_write produces (src/trk.jl:433-495, header :88-145)."""
    from fibers_jl_b200.stream import Tract, trk_write
    M = np.array([[-1.25, 0, 0, 90], [0, 0, 1.5, -126], [0, -2.0, 0, 72], [0, 0, 0, 1]], np.float32)
    lines = [np.asfortranarray(np.array([[1, 2, 3.5], [4, 5, 6.25], [7, 8, 9]], np.float32).T), np.asfortranarray(np.array([[10.5, 2, 3], [1, 1, 1]], np.float32).T)]
    tr = Tract(lines, np.array([3, 2], np.int32), None, dict(volsize=[40, 50, 60], volres=[1.25, 2.0, 1.5], vox2ras0=M))
    assert trk_write(tr, str(tmp_path / "a.trk")) is False
    raw = (tmp_path / "a.trk").read_bytes()
    assert len(raw) == 1000 + 4 * (2 + 3 * 5)
    assert raw[:6] == b"TRACK\0"
    assert struct.unpack_from("<3h", raw, 6) == (40, 50, 60)
    assert struct.unpack_from("<3f", raw, 12) == (1.25, 2.0, 1.5)
    assert struct.unpack_from("<3f", raw, 24) == (0.0, 0.0, 0.0)
    assert struct.unpack_from("<h", raw, 36)[0] == 0 and struct.unpack_from("<h", raw, 238)[0] == 0
    np.testing.assert_array_equal(np.array(struct.unpack_from("<16f", raw, 440), np.float32).reshape(4, 4), M)
    assert raw[948:952] == b"LIA\0" and raw[952:956] == b"LIA\0"           # columns of M: -x, -z, +y
    iop = np.array(struct.unpack_from("<6f", raw, 956))
    want = (np.diag([-1.0, -1.0, 1.0]) @ M[:3, :2].astype(np.float64) @ np.diag([1 / 1.25, 1 / 2.0])).reshape(-1, order="F")
    np.testing.assert_allclose(iop, want, rtol=1e-6)
    assert struct.unpack_from("<3i", raw, 988) == (2, 2, 1000)
    off = 1000
    for ln in lines:
        n = struct.unpack_from("<i", raw, off)[0]; off += 4
        assert n == ln.shape[1]
        pts = np.array(struct.unpack_from(f"<{3 * n}f", raw, off), np.float32).reshape(n, 3); off += 12 * n
        np.testing.assert_array_equal(pts, ((ln.T.astype(np.float64) + 0.5) * np.array([1.25, 2.0, 1.5])).astype(np.float32))


def test_trk_read_layout_and_round_trip(tmp_path):
    """fibers_trk_read_* against a file assembled byte by byte from the TrackVis v2 layout trk_read walks
    (src/trk.jl:358-425): scalars per point, properties per streamline, points back as xyz ./ voxel_size .- .5;
    then trk_write -> trk_read gives back the voxel coordinates stream() produced."""
    from fibers_jl_b200.stream import Tract, trk_read, trk_write
    vs = np.array([1.25, 2.0, 1.5], np.float32)
    M = np.array([[-1.25, 0, 0, 90], [0, 0, 1.5, -126], [0, -2.0, 0, 72], [0, 0, 0, 1]], np.float32)
    h = bytearray(1000)
    h[0:6] = b"TRACK\0"
    struct.pack_into("<3h", h, 6, 40, 50, 60)
    struct.pack_into("<3f", h, 12, *vs)
    struct.pack_into("<3f", h, 24, 1.0, 2.0, 3.0)
    struct.pack_into("<h", h, 36, 2); h[38:41] = b"fa\0"; h[58:61] = b"md\0"
    struct.pack_into("<h", h, 238, 1); h[240:244] = b"len\0"
    struct.pack_into("<16f", h, 440, *M.reshape(-1))
    h[948:952] = b"LIA\0"; h[952:956] = b"LAS\0"
    struct.pack_into("<6f", h, 956, 1, 0, 0, 0, 0, -1)
    struct.pack_into("<3i", h, 988, 3, 2, 1000)
    rng = np.random.default_rng(5)
    body = b""
    lines = []
    for n in (4, 0, 2):                                       # an empty streamline in the middle
        pts = rng.uniform(0, 60, size=(n, 3)).astype(np.float32)
        sc = rng.normal(size=(n, 2)).astype(np.float32)
        prop = rng.normal(size=1).astype(np.float32)
        lines.append((pts, sc, prop))
        body += struct.pack("<i", n) + np.concatenate([pts, sc], axis=1).astype("<f4").tobytes() + prop.astype("<f4").tobytes()
    (tmp_path / "b.trk").write_bytes(bytes(h) + body)
    tr = trk_read(str(tmp_path / "b.trk"))
    assert tr.n_count == 3 and tr.npts.tolist() == [4, 0, 2]
    assert tr.header["n_scalars"] == 2 and tr.header["n_properties"] == 1 and tr.header["version"] == 2 and tr.header["hdr_size"] == 1000
    assert tr.header["dim"].tolist() == [40, 50, 60] and tr.header["origin"].tolist() == [1.0, 2.0, 3.0]
    assert tr.header["voxel_order"] == b"LIA" and tr.header["voxel_order_original"] == b"LAS"
    np.testing.assert_array_equal(tr.header["vox_to_ras"], M)
    np.testing.assert_array_equal(tr.header["image_orientation_patient"], np.array([1, 0, 0, 0, 0, -1], np.float32))
    for i, (pts, sc, prop) in enumerate(lines):
        want = ((pts / vs).astype(np.float32).astype(np.float64) - 0.5).astype(np.float32)     # Float32 ./ Float32 .- .5 (Float64 literal), stored as Float32
        np.testing.assert_array_equal(tr.xyz[i], want.T)
        np.testing.assert_array_equal(tr.scalars[i], sc.T)
        np.testing.assert_array_equal(tr.properties[:, i], prop)
    # truncated file -> error, like the reference's read! throwing EOFError
    (tmp_path / "c.trk").write_bytes((bytes(h) + body)[:-6])
    with pytest.raises(Exception):
        trk_read(str(tmp_path / "c.trk"))
    # round trip of a written tract
    out = [np.asfortranarray(rng.uniform(0, 40, size=(3, n)).astype(np.float32)) for n in (5, 1, 7)]
    t0 = Tract(out, np.array([5, 1, 7], np.int32), None, dict(volsize=[40, 50, 60], volres=vs, vox2ras0=M))
    assert trk_write(t0, str(tmp_path / "d.trk")) is False
    t1 = trk_read(str(tmp_path / "d.trk"))
    assert t1.npts.tolist() == [5, 1, 7] and t1.header["n_scalars"] == 0 and t1.header["n_properties"] == 0
    for a, b in zip(out, t1.xyz):
        np.testing.assert_allclose(b, a, rtol=0, atol=2e-5)                # (x + .5) * vs / vs - .5: two roundings at |x| <= 60
    # scalars per point (the method-difference flags of an LCM run) and a property per streamline survive the round trip
    sc = [np.asfortranarray((rng.random((1, n)) < 0.5).astype(np.float32)) for n in (5, 1, 7)]
    t2 = Tract(out, np.array([5, 1, 7], np.int32), None, dict(volsize=[40, 50, 60], volres=vs, vox2ras0=M), scalars=sc,
               properties=np.array([[1.5, 2.5, 3.5]], np.float32))
    assert trk_write(t2, str(tmp_path / "e.trk")) is False
    raw = (tmp_path / "e.trk").read_bytes()
    assert len(raw) == 1000 + 4 * (3 + 4 * 13 + 3) and struct.unpack_from("<h", raw, 36)[0] == 1 and struct.unpack_from("<h", raw, 238)[0] == 1
    t3 = trk_read(str(tmp_path / "e.trk"))
    for a, b in zip(sc, t3.scalars):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(t3.properties, np.array([[1.5, 2.5, 3.5]], np.float32))


def test_result_writers_file_names_and_content(tmp_path):
    """dti_write / gqi_write / dsi_write / rumba_write (src/dti.jl:344, src/gqi.jl:210, src/dsi.jl:279, src/rusd.jl:645): one
    <basename>_<field>[i].nii.gz per volume, RUMBA-SD's scalars as text; the volumes read back unchanged."""
    rng = np.random.default_rng(2)
    M = np.diag([2.0, 2.0, 2.0, 1.0]).astype(np.float32)

    def vol(*shape):
        return Fb.MRI(np.asfortranarray(rng.normal(size=shape).astype(np.float32)), header=dict(vox2ras0=M))
    d = Fb.DTI(*(vol(3, 4, 2) for _ in range(4)), *(vol(3, 4, 2, 3) for _ in range(3)), *(vol(3, 4, 2) for _ in range(3)))
    Fb.dti_write(d, str(tmp_path / "sub_dti"))
    names = sorted(p.name for p in tmp_path.iterdir())
    assert names == sorted(f"sub_dti_{f}.nii.gz" for f in ("s0", "eigval1", "eigval2", "eigval3", "eigvec1", "eigvec2", "eigvec3", "rd", "md", "fa"))
    np.testing.assert_array_equal(fio.mri_read(str(tmp_path / "sub_dti_eigvec2.nii.gz")).vol, d.eigvec2.vol)
    g = Fb.GQI(vol(3, 4, 2, 5), [vol(3, 4, 2, 3) for _ in range(3)], [vol(3, 4, 2) for _ in range(3)])
    Fb.gqi_write(g, str(tmp_path / "g"))
    assert sorted(p.name for p in tmp_path.glob("g_*")) == sorted(["g_odf.nii.gz"] + [f"g_{n}{i}.nii.gz" for n in ("peak", "qa") for i in (1, 2, 3)])
    np.testing.assert_array_equal(fio.mri_read(str(tmp_path / "g_qa3.nii.gz")).vol.reshape(3, 4, 2), g.qa[2].vol)
    s = Fb.DSI(vol(3, 4, 2, 7), vol(3, 4, 2, 5), [vol(3, 4, 2, 3) for _ in range(3)], [vol(3, 4, 2) for _ in range(3)])
    Fb.dsi_write(s, str(tmp_path / "s"))
    assert (tmp_path / "s_pdf.nii.gz").exists() and (tmp_path / "s_peak2.nii.gz").exists() and len(list(tmp_path.glob("s_*"))) == 8
    r = Fb.RUMBASD(vol(3, 4, 2, 5), vol(3, 4, 2), vol(3, 4, 2), [vol(3, 4, 2, 3) for _ in range(5)], vol(3, 4, 2), vol(3, 4, 2),
                   np.float32(21.5), np.float32(3.25))
    Fb.rumba_write(r, str(tmp_path / "r"))
    assert len(list(tmp_path.glob("r_peak*.nii.gz"))) == 5 and (tmp_path / "r_fodf.nii.gz").exists()
    assert (tmp_path / "r_snr_mean.txt").read_text() == "21.5\n" and (tmp_path / "r_snr_std.txt").read_text() == "3.25\n"
