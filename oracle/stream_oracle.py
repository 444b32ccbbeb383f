"""CPU restatement of the reference's deterministic streamline tractography (TEST INFRASTRUCTURE: only
tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this; the product path never does).

Follows /root/reference/src/stream.jl line by line for the regime the GPU path covers: orientation vectors
given as 3-D vectors, in the macroscopic regime (voxel size > 50 um) and in the microscopy regime (`domicro`: the
next POSITION is searched in a box around the tentative step), and -- macroscopic regime only -- with local
connection matrices (`lcms`): that branch draws from `rand(Categorical(...))` on Julia's task-local generator, so
its answer depends on the thread schedule; here the uniform numbers come from a counter-based generator
(`lcm_uniform(seed, line, k)`: draw k of streamline `line`), an INPUT convention shared with the C ABI.

    StreamWork (mask / vector masking)        src/stream.jl:72-140
    stream_pick_by_angle!                     src/stream.jl:355-387
    stream_new_point!                         src/stream.jl:497-541
    stream_micro_new_point! (+ search area)   src/stream.jl:547-617, :266-300
    stream_pick_by_lcm! (+ LCM set-up)        src/stream.jl:380-494, :208-236
    stream_new_line                           src/stream.jl:621-690
    stream (seed order, len_min filter)       src/stream.jl:730-790

PARITY UNPINNED: the reference ships no tests or vectors for this path and Julia is not available.  Arithmetic
that lives outside /root/reference is restated from its published behaviour:
  * `round(Int, x)`                     round-half-to-even;
  * `dot(::Vector{Float32}, ::SubArray)` LinearAlgebra -> BLAS sdot; restated as the sequential fp32 sum
                                        (a1 b1 + a2 b2) + a3 b3 with every product and sum rounded to fp32 (the
                                        true kernel is platform dependent: OpenBLAS may keep a wider accumulator);
  * `norm(::Vector{Float32})` (n < 32)  LinearAlgebra.generic_norm2: squares in fp32, sum and sqrt in fp64,
                                        result converted to fp32;
  * `argmax`                            first maximum, NaN counts as the largest value;
  * `cosd(T(ang))`                      evaluated by the caller (the wrapper), passed in as `cosang_thresh`.
  * `rand(Categorical(p))`              Distributions.jl, DiscreteNonParametric: draw = rand(Float32); cp = p[1];
                                        i = 1; while cp <= draw && i < n: cp += p[i += 1]; return i  (recollection);
  * `sum(::Vector{Float32})` (n = 10)   restated as the sequential fp32 sum (Base's loop carries @simd, so the true
                                        association is up to LLVM; it only moves the thresholds of the draw by 1 ulp).
The sub-voxel offsets are random in the reference (`rand(Uniform(-.5+eps(), .5-eps()), 3)`, :177-183): they are
an INPUT here and in the C ABI, so that both sides track the same seeds.

Coordinates are the reference's: 1-based voxel indices, positions in voxel units.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def stream_work(ovecs, f=None, f_thresh=0.03, fa=None, fa_thresh=0.1, mask=None):
    """StreamWork constructor, src/stream.jl:72-140: returns (mask_array [nx,ny,nz] bool, ovec_array [3,nvec,nx,ny,nz] f32).
    ovecs: list of [nx,ny,nz,3] arrays; f: list of [nx,ny,nz] or None; fa, mask: [nx,ny,nz] or None."""
    ovecs = [np.asarray(o, dtype=F) for o in ovecs]
    nx, ny, nz = ovecs[0].shape[:3]
    if mask is None:                                         # :107-112
        m = np.zeros((nx, ny, nz), dtype=bool)
        for o in ovecs:
            m |= np.any(o != 0, axis=3)
    else:                                                    # :114
        mk = np.asarray(mask)
        m = (mk.reshape(nx, ny, nz, -1)[..., 0] > 0)
    if fa is not None:                                       # :117-128
        m = m & (np.asarray(fa, dtype=F).reshape(nx, ny, nz, -1)[..., 0] >= F(fa_thresh))
    arr = np.zeros((3, len(ovecs), nx, ny, nz), dtype=F)
    for i, o in enumerate(ovecs):                            # :133-147
        om = m if f is None else (m & (np.asarray(f[i], dtype=F).reshape(nx, ny, nz) >= F(f_thresh)))
        for d in range(3):
            arr[d, i] = o[..., d] * om
    return m, arr


def _dot3(a, b):
    return F(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2]))


def _rnd(x):
    return int(np.rint(x))                                   # half to even, like Julia's round(Int, x)


class _State:
    __slots__ = ("pos_now", "vec_now", "pos_next", "vec_next", "ivec_next", "isdiff")


def _pick_by_angle(st, ix, iy, iz, ovec):
    """src/stream.jl:355-387 (ix, iy, iz 1-based)."""
    nvec = ovec.shape[1]
    cos = np.empty(nvec, dtype=F)
    cab = np.empty(nvec, dtype=F)
    for i in range(nvec):
        v = ovec[:, i, ix - 1, iy - 1, iz - 1]
        if v[0] == 0 and v[1] == 0 and v[2] == 0:
            cos[i] = cab[i] = -np.inf
        else:
            cos[i] = _dot3(st.vec_now, v)
            cab[i] = abs(cos[i])
    nan = np.isnan(cab)
    k = int(np.argmax(nan)) if nan.any() else int(np.argmax(cab))     # Julia argmax: first NaN wins, else first maximum
    if not np.isfinite(cos[k]):
        return False
    v = ovec[:, k, ix - 1, iy - 1, iz - 1]
    st.vec_next = v.copy() if cos[k] > 0 else (-v).astype(F)
    st.ivec_next = k + 1
    return True


def _new_point(st, mask, ovec, step):
    """src/stream.jl:497-541 without LCMs."""
    st.pos_next = (st.pos_now + (st.vec_now * step).astype(F)).astype(F)
    ix, iy, iz = _rnd(st.pos_next[0]), _rnd(st.pos_next[1]), _rnd(st.pos_next[2])
    nx, ny, nz = mask.shape
    if not (1 <= ix <= nx and 1 <= iy <= ny and 1 <= iz <= nz):
        return False
    if not mask[ix - 1, iy - 1, iz - 1]:
        return False
    return _pick_by_angle(st, ix, iy, iz, ovec)


EDGETYPE = np.array([[1, 1, 1, 1, 2, 2, 2, 3, 3, 4],
                     [1, 2, 3, 4, 2, 3, 4, 3, 4, 4]])        # voxel edges connected by the i-th element of a vectorised LCM (:234-235)
_M64 = (1 << 64) - 1


def lcm_uniform(seed, line, k):
    """Draw k (0-based) of streamline `line` (0-based, reference order before the len_min filter): splitmix64 finaliser of
    seed + (line + 1) * 0x9E3779B97F4A7C15 + (k + 1) * 0xD1B54A32D192ED03, top 24 bits -> [0, 1) in fp32."""
    z = (int(seed) + (int(line) + 1) * 0x9E3779B97F4A7C15 + (int(k) + 1) * 0xD1B54A32D192ED03) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    z ^= z >> 31
    return F((z >> 40) * 2.0 ** -24)


def lcm_work(lcms, lcm_thresh, ovec0):
    """LCM part of the StreamWork constructor, src/stream.jl:208-236.  lcms: [nx,ny,nz,10]; ovec0: the FIRST orientation volume
    as given ([nx,ny,nz,3]).  Returns (lcm_array [10,nx,ny,nz], strdims (0-based), dxyz [3,4])."""
    arr = np.transpose(np.asarray(lcms, dtype=F), (3, 0, 1, 2)).copy()
    arr = np.where(arr.astype(np.float64) >= float(lcm_thresh), arr, F(0)).astype(F)   # .*= (>= thresh): false is a strong zero
    thru = [d for d in range(3) if np.all(np.asarray(ovec0)[..., d] == 0)]             # :224
    strdims = [d for d in range(3) if d not in thru]                                   # :226
    dxyz = np.zeros((3, 4), dtype=int)
    dxyz[strdims[0], :] = [-1, 0, 1, 0]                                                # :229-231 (IndexError = the reference's BoundsError)
    dxyz[strdims[1], :] = [0, -1, 0, 1]
    return arr, strdims, dxyz


def _pick_by_lcm(st, ix, iy, iz, ovec, L, draw):
    """src/stream.jl:380-494 (ix, iy, iz 1-based); `draw()` returns the next uniform number of this streamline."""
    arr, strdims, dxyz = L
    now = [_rnd(st.pos_now[0]), _rnd(st.pos_now[1]), _rnd(st.pos_now[2])]
    dvox = np.array([now[0] - ix, now[1] - iy, now[2] - iz])
    if not dvox.any():                                       # same voxel: keep the vector chosen last (:400-411)
        v = ovec[:, st.ivec_next - 1, ix - 1, iy - 1, iz - 1]
        st.vec_next = v.copy() if _dot3(st.vec_now, v) > 0 else (-v).astype(F)
        return True

    def edge(d):
        for j in range(4):
            if np.array_equal(d, dxyz[:, j]):
                return j + 1
        return 0
    entry = edge(dvox)                                       # :414-420
    if entry == 0:                                           # diagonal jump: which dimension changes faster (:422-437)
        a, b = strdims[0], strdims[1]
        if abs(F(st.pos_now[a] - st.pos_next[a])) < abs(F(st.pos_now[b] - st.pos_next[b])):
            dvox[b] = 0
        else:
            dvox[a] = 0
        entry = edge(dvox)
    lcm = arr[:, ix - 1, iy - 1, iz - 1].copy()              # :440-445
    for j in range(10):
        if entry not in EDGETYPE[:, j]:
            lcm[j] = 0
    while lcm.any():                                         # :447 (runs once: the acceptance test is `if true`)
        tot = F(0)
        for x in lcm:
            tot = F(tot + x)
        lcm = (lcm / tot).astype(F)
        u = draw()
        cp, i = lcm[0], 0
        while cp <= u and i < 9:
            i += 1
            cp = F(cp + lcm[i])
        ilcm = i
        exit_ = EDGETYPE[1, ilcm] if EDGETYPE[0, ilcm] == entry else EDGETYPE[0, ilcm]      # :453-454
        nvec = ovec.shape[1]
        cos = np.empty(nvec, dtype=F)
        cab = np.empty(nvec, dtype=F)
        dj = dxyz[:, exit_ - 1].astype(F)
        for i in range(nvec):                                # :460-469
            v = ovec[:, i, ix - 1, iy - 1, iz - 1]
            if v[0] == 0 and v[1] == 0 and v[2] == 0:
                cos[i] = cab[i] = -np.inf
            else:
                cos[i] = _dot3(dj, v)
                cab[i] = abs(cos[i])
        nan = np.isnan(cab)
        k = int(np.argmax(nan)) if nan.any() else int(np.argmax(cab))
        if not np.isfinite(cos[k]):
            return False
        v = ovec[:, k, ix - 1, iy - 1, iz - 1]
        st.vec_next = v.copy() if cos[k] > 0 else (-v).astype(F)
        st.ivec_next = k + 1
        return True
    return False                                             # :493


def _new_point_lcm(st, mask, ovec, step, L, draw):
    """src/stream.jl:497-541 with LCMs: the conventional pick first (for the method-difference flag), then the LCM pick."""
    st.pos_next = (st.pos_now + (st.vec_now * step).astype(F)).astype(F)
    ix, iy, iz = _rnd(st.pos_next[0]), _rnd(st.pos_next[1]), _rnd(st.pos_next[2])
    nx, ny, nz = mask.shape
    if not (1 <= ix <= nx and 1 <= iy <= ny and 1 <= iz <= nz):
        return False
    if not mask[ix - 1, iy - 1, iz - 1]:
        return False
    if not _pick_by_angle(st, ix, iy, iz, ovec):
        return False
    ivec_ang = st.ivec_next
    if not _pick_by_lcm(st, ix, iy, iz, ovec, L, draw):
        return False
    st.isdiff = st.ivec_next != ivec_ang
    return True


def search_area(dist):
    """Unit vectors from the centre of the (2 d1 + 1, 2 d2 + 1, 2 d3 + 1) search box to its voxels, zero outside the unit
    ellipsoid -- src/stream.jl:268-292.  The centre itself is 0 / 0 = NaN (the reference keeps it: every comparison with
    NaN is false, so the centre voxel always passes the cone test, :573-575)."""
    d = [int(x) for x in dist]
    A = np.zeros((2 * d[0] + 1, 2 * d[1] + 1, 2 * d[2] + 1, 3), dtype=F)
    with np.errstate(invalid="ignore", divide="ignore"):
        for iz in range(A.shape[2]):
            for iy in range(A.shape[1]):
                for ix in range(A.shape[0]):
                    rx = F(F(ix - d[0]) / F(d[0] + F(0.5)))
                    ry = F(F(iy - d[1]) / F(d[1] + F(0.5)))
                    rz = F(F(iz - d[2]) / F(d[2] + F(0.5)))
                    r = F(np.sqrt(F(F(F(rx * rx) + F(ry * ry)) + F(rz * rz))))
                    if r < 1:
                        A[ix, iy, iz] = [F(rx / r), F(ry / r), F(rz / r)]
    return A


def _micro_new_point(st, mask, ovec, step, dist, area, search_cosang):
    """stream_micro_new_point!, src/stream.jl:547-617: tentative step, then the voxel of the search box (inside the mask and
    inside the cone around the current direction) whose vector is most similar to the current one becomes the next POSITION."""
    st.pos_next = (st.pos_now + (st.vec_now * step).astype(F)).astype(F)
    ix, iy, iz = _rnd(st.pos_next[0]), _rnd(st.pos_next[1]), _rnd(st.pos_next[2])
    nx, ny, nz = mask.shape
    if not (1 <= ix <= nx and 1 <= iy <= ny and 1 <= iz <= nz):
        return False
    if not mask[ix - 1, iy - 1, iz - 1]:
        return False
    best = None                                          # (value, is_nan) of the first maximum in column-major order of the box
    bcos = F(-np.inf); bpos = None
    d = dist
    for kz in range(-d[2], d[2] + 1):
        for ky in range(-d[1], d[1] + 1):
            for kx in range(-d[0], d[0] + 1):
                x, y, z = ix + kx, iy + ky, iz + kz
                val = F(-np.inf); cos = F(-np.inf)
                if 1 <= x <= nx and 1 <= y <= ny and 1 <= z <= nz:
                    v = area[kx + d[0], ky + d[1], kz + d[2]]
                    skip = (not mask[x - 1, y - 1, z - 1]) or (v[0] == 0 and v[1] == 0 and v[2] == 0) or (_dot3(st.vec_now, v) <= search_cosang)
                    if not skip:
                        cos = _dot3(st.vec_now, ovec[:, 0, x - 1, y - 1, z - 1])
                        val = F(abs(cos))
                if best is None:
                    best = val; bcos = cos; bpos = (x, y, z)
                elif not np.isnan(best) and (np.isnan(val) or val > best):
                    best = val; bcos = cos; bpos = (x, y, z)
    if not np.isfinite(bcos):
        return False
    st.pos_next = np.array(bpos, dtype=F)
    v = ovec[:, 0, bpos[0] - 1, bpos[1] - 1, bpos[2] - 1]
    st.vec_next = v.copy() if bcos > 0 else (-v).astype(F)
    return True


def new_line(seed_vox, sub_vox, mask, ovec, len_max, cosang_thresh, step, smooth, micro=None, lcm=None):
    """src/stream.jl:621-690: returns the [3, npts] streamline of one (seed voxel, sub-voxel offset); with `lcm` =
    (lcm_work(...), draw) also the method-difference flag of every point (second return value)."""
    step, smooth, cosang_thresh = F(step), F(smooth), F(cosang_thresh)
    st = _State()
    st.ivec_next = 1                                          # :646 (NOT reset between the two directions)
    npts = 0
    fwd_pts, bwd_pts = [], []
    fwd_flag, bwd_flag = [], []
    seed = np.asarray(seed_vox, dtype=F)
    for fwd in (1, -1):
        st.pos_now = (seed + np.asarray(sub_vox, dtype=F)).astype(F)
        st.vec_now = (ovec[:, st.ivec_next - 1, seed_vox[0] - 1, seed_vox[1] - 1, seed_vox[2] - 1] * F(fwd)).astype(F)
        while True:
            if lcm is not None:
                ok = _new_point_lcm(st, mask, ovec, step, *lcm)
            else:
                ok = _micro_new_point(st, mask, ovec, step, *micro) if micro is not None else _new_point(st, mask, ovec, step)
            if not ok:
                break
            (fwd_pts if fwd == 1 else bwd_pts).append(st.pos_now.copy())     # prepend! / append! (:660, :666)
            npts += 1
            if lcm is not None:
                (fwd_flag if fwd == 1 else bwd_flag).append(bool(st.isdiff))  # :671-674; no angle threshold with LCMs (:676)
            elif _dot3(st.vec_now, st.vec_next) < cosang_thresh:              # :677
                break
            if npts > len_max:                                                # :681
                break
            if smooth != 0:                                                   # :684-688
                one_minus = F(F(1) - smooth)
                vn = ((smooth * st.vec_now).astype(F) + (one_minus * st.vec_next).astype(F)).astype(F)
                nrm = F(np.sqrt(np.float64(F(vn[0] * vn[0])) + np.float64(F(vn[1] * vn[1])) + np.float64(F(vn[2] * vn[2]))))
                st.vec_next = (vn / nrm).astype(F)
            st.pos_now = st.pos_next
            st.vec_now = st.vec_next
    pts = fwd_pts[::-1] + bwd_pts
    xyz = np.array(pts, dtype=F).reshape(-1, 3).T
    if lcm is not None:
        return xyz, np.array(fwd_flag[::-1] + bwd_flag, dtype=bool)
    return xyz


def stream(ovecs, sublist, f=None, f_thresh=0.03, fa=None, fa_thresh=0.1, mask=None, seed=None, len_min=3, len_max=None,
           cosang_thresh=None, step_size=0.5, smooth_coeff=0.2, micro_search_dist=None, micro_search_cosang=None,
           lcms=None, lcm_thresh=0.099, lcm_seed=0):
    """src/stream.jl:730-790.  Returns the list of [3, npts] streamlines in the reference's order (seed voxels in
    column-major order, sub-voxel samples innermost), lines shorter than len_min dropped; with `lcms` ([nx,ny,nz,10])
    a second list with the method-difference flags of every point (the scalars of the reference's Tract, :783)."""
    m, arr = stream_work(ovecs, f, f_thresh, fa, fa_thresh, mask)
    nx, ny, nz = m.shape
    if len_max is None:
        len_max = max(nx, ny, nz)
    if cosang_thresh is None:
        cosang_thresh = F(np.cos(np.deg2rad(45.0)))
    sm = m if seed is None else (np.asarray(seed).reshape(nx, ny, nz, -1)[..., 0] > 0)
    lin = np.flatnonzero(sm.reshape(-1, order="F"))           # findall: column-major order
    micro = None
    if micro_search_dist is not None:                        # microscopy regime (domicro, src/stream.jl:84-90, :266-300)
        dist = [int(x) for x in micro_search_dist]
        micro = (dist, search_area(dist), F(micro_search_cosang))
    L = None
    if lcms is not None:
        if micro is not None:
            raise ValueError("local connection matrices are only defined for the macroscopic regime")
        L = lcm_work(lcms, lcm_thresh, ovecs[0])
    out, flags = [], []
    line = 0
    for l in lin:
        vox = [int(l % nx) + 1, int((l // nx) % ny) + 1, int(l // (nx * ny)) + 1]
        for sub in sublist:
            if L is not None:
                cnt = [0]

                def draw(line=line, cnt=cnt):
                    u = lcm_uniform(lcm_seed, line, cnt[0]); cnt[0] += 1
                    return u
                s, fl = new_line(vox, sub, m, arr, len_max, cosang_thresh, step_size, smooth_coeff, None, (L, draw))
                if s.shape[1] >= len_min:
                    out.append(s); flags.append(fl)
            else:
                s = new_line(vox, sub, m, arr, len_max, cosang_thresh, step_size, smooth_coeff, micro)
                if s.shape[1] >= len_min:
                    out.append(s)
            line += 1
    return (out, flags) if L is not None else out
