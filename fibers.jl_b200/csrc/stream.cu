// Deterministic streamline tractography on the GPU (`stream`, the downstream consumer of
// the GQI / DSI / DTI peaks).  Replaces the `Threads.@threads` seed loop of the reference (src/stream.jl:730-790) and the
// per-seed propagation it calls (stream_new_line :621-690, stream_new_point! :497-541, stream_pick_by_angle! :355-387)
// for orientation VECTORS: macroscopic voxels, the microscopy regime (stream_micro_new_point! :547-617) and -- macroscopic
// only -- local connection matrices (stream_pick_by_lcm! :380-494).  That branch draws from rand(Categorical(...)) on Julia's
// task-local generator (its answer depends on the thread schedule); here draw k of streamline `line` is a counter-based
// uniform number (lcm_uniform), so the counting pass and the writing pass see the same draws and the oracle can follow.
//
//   stream_pack_kernel   the StreamWork constructor (:72-147): voxel mask (given, or "any vector component non-zero"),
//                        intersected with fa >= fa_thresh; vectors zeroed outside the mask / where f[ivec] < f_thresh;
//                        packed as [voxel][ivec][3] (= W.ovecs[3, nvec, nx, ny, nz]), so that one propagation step reads
//                        12 nvec contiguous bytes.
//   stream_track_kernel  one thread per (seed voxel, sub-voxel sample), seeds in column-major order, samples innermost
//                        (the reference's output order).  Run twice: a counting pass (points forward / backward), an
//                        exclusive scan over the lines that reach len_min, and a writing pass that puts every point at
//                        its final place -- forward points are PREPENDED by the reference (:660), so forward point i of nf
//                        lands at nf - 1 - i and backward point j at nf + j.  Nothing is compacted on the host.
// The reference's quirks are kept: the seed position is stored once per direction, the point counter runs on across
// the two directions (:681), and the backward pass starts along vector `ivec_next` of the SEED voxel where `ivec_next`
// is whatever the forward pass chose last (:646, :653).
//
// Arithmetic follows the reference's fp32 operation order (this file is compiled with -fmad=false): positions and
// vectors in fp32, dot products as (a1 b1 + a2 b2) + a3 b3, `norm` = squares in fp32, sum and square root in fp64
// (LinearAlgebra.generic_norm2), IEEE division, round-half-even for round(Int, x).
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <algorithm>
#include <cmath>
#include <vector>
#include "common.cuh"

namespace fibers {
namespace {

constexpr int MAX_NVEC = 8;

struct PackIn { const float* ovec[MAX_NVEC]; const float* f[MAX_NVEC]; };

// mask_out[v] (u8), ovec_out[v][i][3]; inputs are the reference's volumes: component c of vector volume i at ovec[i][c * nvox + v]
__global__ void stream_pack_kernel(PackIn in, int nvec, int has_f, float f_thresh, const float* __restrict__ fa, float fa_thresh,
                                   const uint8_t* __restrict__ mask, int64_t nvox, uint8_t* __restrict__ mask_out, float* __restrict__ ovec_out) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    float x[MAX_NVEC][3];
    bool m;
    if (mask) m = mask[v] != 0;
    else m = false;
    for (int i = 0; i < nvec; ++i)
        for (int c = 0; c < 3; ++c) {
            x[i][c] = in.ovec[i][(int64_t)c * nvox + v];
            if (!mask && x[i][c] != 0.f) m = true;                    // (:107-112) NaN != 0 counts, as in Julia
        }
    if (fa) m = m && (fa[v] >= fa_thresh);                            // (:128)
    mask_out[v] = m ? 1 : 0;
    for (int i = 0; i < nvec; ++i) {
        const bool om = has_f ? (m && in.f[i][v] >= f_thresh) : m;    // (:136-138)
        for (int c = 0; c < 3; ++c) ovec_out[(v * nvec + i) * 3 + c] = om ? x[i][c] : 0.f;
    }
}

struct TrackParams {
    const float* ovec; const uint8_t* mask; const int32_t* seeds; int64_t nseed; const float* sub; int nsub;
    int nx, ny, nz, nvec, len_min, len_max; float cos_thresh, step, smooth;
    int sd[3]; float search_cos;       // microscopy regime: half widths of the search box, cosine of the search angle
    const float* lcm; int s1, s2; unsigned long long lcm_seed;   // LCM branch: thresholded matrices [voxel][10], in-plane dimensions (0-based), generator seed
};

// draw k of streamline `line`: splitmix64 finaliser of seed + (line + 1) * golden + (k + 1) * 0xD1B54A32D192ED03, top 24 bits -> [0, 1)
__device__ __forceinline__ float lcm_uniform(unsigned long long seed, long long line, int k) {
    unsigned long long z = seed + (unsigned long long)(line + 1) * 0x9E3779B97F4A7C15ull + (unsigned long long)(k + 1) * 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * 5.9604644775390625e-8f;
}

// lcm_out[v][j] = lcms[j][v] where it is >= lcm_thresh (compared in fp64 like Float32 .>= Float64), else 0  (:209, :220)
__global__ void stream_pack_lcm_kernel(const float* __restrict__ lcms, double thresh, int64_t nvox, float* __restrict__ lcm_out) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    for (int j = 0; j < 10; ++j) {
        const float x = lcms[(int64_t)j * nvox + v];
        lcm_out[v * 10 + j] = ((double)x >= thresh) ? x : 0.f;
    }
}

__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2) { return (a0 * b0 + a1 * b1) + a2 * b2; }

// kWrite == false: counts -> nfb[line] = (points forward, points backward).
// kWrite == true : points -> xyz + 3 * off[line] for the lines with kept[line] != 0 (nfb gives nf).
// kLcm: the LCM branch of stream_new_point! (:523-538): conventional pick first (for the method-difference flag), then the
// pick by local connection matrix; every point also gets its flag (scal), and the angle threshold is not applied (:671-678).
template <bool kWrite, bool kLcm>
__global__ void __launch_bounds__(128) stream_track_kernel(TrackParams P, int2* __restrict__ nfb, const int64_t* __restrict__ off,
                                                            const int32_t* __restrict__ sidx, const uint8_t* __restrict__ kept,
                                                            float* __restrict__ xyz, int32_t* __restrict__ npts_out, float* __restrict__ scal) {
    const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= P.nseed * P.nsub) return;
    int nf_known = 0;
    float* out = nullptr;
    if (kWrite) {
        if (!kept[line]) return;
        const int2 c = nfb[line];
        nf_known = c.x;
        out = xyz + 3 * off[line];
        npts_out[sidx[line]] = c.x + c.y;
    }
    const int64_t si = line / P.nsub; const int isub = (int)(line - si * P.nsub);
    const int lin = P.seeds[si];
    const int sx = lin % P.nx, sy = (lin / P.nx) % P.ny, sz = lin / (P.nx * P.ny);      // 0-based seed voxel
    const float s0 = P.sub[isub * 3 + 0], s1 = P.sub[isub * 3 + 1], s2 = P.sub[isub * 3 + 2];
    int ivec = 0;                                          // W.ivec_next[tid] - 1 (:646): NOT reset between the directions
    int npts = 0, nf = 0, nb = 0;
    int ndraw = 0;                                         // LCM branch: uniform numbers this line has consumed
    for (int dir = 0; dir < 2; ++dir) {
        const float fwd = dir == 0 ? 1.f : -1.f;
        float px = (float)(sx + 1) + s0, py = (float)(sy + 1) + s1, pz = (float)(sz + 1) + s2;     // 1-based coordinates (:652)
        const float* sv = P.ovec + ((int64_t)lin * P.nvec + ivec) * 3;
        float vx = sv[0] * fwd, vy = sv[1] * fwd, vz = sv[2] * fwd;                                 // (:653)
        while (true) {
            // ---- stream_new_point! (:497-541) ----
            const float qx = px + vx * P.step, qy = py + vy * P.step, qz = pz + vz * P.step;
            const int ix = __float2int_rn(qx), iy = __float2int_rn(qy), iz = __float2int_rn(qz);
            if (ix < 1 || ix > P.nx || iy < 1 || iy > P.ny || iz < 1 || iz > P.nz) break;
            const int64_t nl = (int64_t)(ix - 1) + (int64_t)P.nx * ((iy - 1) + (int64_t)P.ny * (iz - 1));
            if (!P.mask[nl]) break;
            // ---- stream_pick_by_angle! (:355-387): argmax of |cos|, first maximum, NaN first ----
            const float* ov = P.ovec + nl * P.nvec * 3;
            int best = 0; float bcos = 0.f, babs = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
            for (int i = 0; i < P.nvec; ++i) {
                const float wx = ov[3 * i], wy = ov[3 * i + 1], wz = ov[3 * i + 2];
                float c, a;
                if (wx == 0.f && wy == 0.f && wz == 0.f) c = a = -INFINITY;
                else { c = dot3(vx, vy, vz, wx, wy, wz); a = fabsf(c); }
                if (i == 0 || (!isnan(babs) && (isnan(a) || a > babs))) { best = i; bcos = c; babs = a; bx = wx; by = wy; bz = wz; }
            }
            if (!isfinite(bcos)) break;
            float nxv, nyv, nzv;
            if (bcos > 0.f) { nxv = bx; nyv = by; nzv = bz; } else { nxv = -bx; nyv = -by; nzv = -bz; }
            ivec = best;
            bool isdiff = false;
            if (kLcm) {
                // ---- stream_pick_by_lcm! (:380-494).  In the voxel the line is already in, the vector chosen last is kept -- and
                //      "last" is the conventional pick a few lines up (:399-411), so there is nothing to change. ----
                const float pn[3] = {px, py, pz}, qn[3] = {qx, qy, qz};
                int d[3] = {__float2int_rn(px) - ix, __float2int_rn(py) - iy, __float2int_rn(pz) - iz};
                if (d[0] != 0 || d[1] != 0 || d[2] != 0) {
                    const int s1 = P.s1, s2 = P.s2, st = 3 - s1 - s2;
                    auto edge = [&]() {                       // dxyz columns (:229-231): 1 = (-1, 0), 2 = (0, -1), 3 = (1, 0), 4 = (0, 1) in (s1, s2)
                        if (d[st] != 0) return 0;
                        if (d[s2] == 0) return d[s1] == -1 ? 1 : d[s1] == 1 ? 3 : 0;
                        if (d[s1] == 0) return d[s2] == -1 ? 2 : d[s2] == 1 ? 4 : 0;
                        return 0;
                    };
                    int entry = edge();
                    if (entry == 0) {                          // diagonal jump: keep the dimension that changes faster (:422-437)
                        if (fabsf(pn[s1] - qn[s1]) < fabsf(pn[s2] - qn[s2])) d[s2] = 0; else d[s1] = 0;
                        entry = edge();
                    }
                    // connections of the entry edge (:440-445): element j joins edges E1[j], E2[j]
                    const int E1[10] = {1, 1, 1, 1, 2, 2, 2, 3, 3, 4}, E2[10] = {1, 2, 3, 4, 2, 3, 4, 3, 4, 4};
                    float lcm[10];
                    bool any = false;
                    float tot = 0.f;
#pragma unroll
                    for (int j = 0; j < 10; ++j) {
                        const float x = P.lcm[nl * 10 + j];
                        lcm[j] = (E1[j] == entry || E2[j] == entry) ? x : 0.f;
                        any = any || lcm[j] != 0.f;            // !iszero(lcm): NaN counts as non-zero
                        tot += lcm[j];
                    }
                    if (!any) break;                           // (:447, :493)
                    const float u = lcm_uniform(P.lcm_seed, line, ndraw++);
                    float cp = lcm[0] / tot;                  // lcm ./= sum(lcm); rand(Categorical(lcm)): first i with cumulative p > u
                    int il = 0;
#pragma unroll
                    for (int j = 1; j < 10; ++j)
                        if (il == j - 1 && cp <= u) { il = j; cp += lcm[j] / tot; }
                    const int ex = E1[il] == entry ? E2[il] : E1[il];                                // (:453-454)
                    float dj[3] = {0.f, 0.f, 0.f};
                    dj[ex == 1 || ex == 3 ? s1 : s2] = (ex == 1 || ex == 2) ? -1.f : 1.f;
                    int kb = 0; float kcos = 0.f, kabs = 0.f;
                    for (int i = 0; i < P.nvec; ++i) {                                                // (:460-471)
                        const float wx = ov[3 * i], wy = ov[3 * i + 1], wz = ov[3 * i + 2];
                        float c, a;
                        if (wx == 0.f && wy == 0.f && wz == 0.f) c = a = -INFINITY;
                        else { c = dot3(dj[0], dj[1], dj[2], wx, wy, wz); a = fabsf(c); }
                        if (i == 0 || (!isnan(kabs) && (isnan(a) || a > kabs))) { kb = i; kcos = c; kabs = a; }
                    }
                    if (!isfinite(kcos)) break;                                                       // (:473)
                    const float wx = ov[3 * kb], wy = ov[3 * kb + 1], wz = ov[3 * kb + 2];
                    if (kcos > 0.f) { nxv = wx; nyv = wy; nzv = wz; } else { nxv = -wx; nyv = -wy; nzv = -wz; }
                    ivec = kb;
                    isdiff = kb != best;                                                              // (:537)
                }
            }
            // ---- stream_new_line: save the CURRENT position (:660 / :666) ----
            if (kWrite) {
                const int64_t at = dir == 0 ? nf_known - 1 - nf : nf_known + nb;
                float* o = out + 3 * at;
                o[0] = px; o[1] = py; o[2] = pz;
                if (kLcm) scal[off[line] + at] = isdiff ? 1.f : 0.f;                                  // (:671-674)
            }
            if (dir == 0) ++nf; else ++nb;
            ++npts;
            if (!kLcm && dot3(vx, vy, vz, nxv, nyv, nzv) < P.cos_thresh) break;                     // (:677; not used with LCMs, :676)
            if (npts > P.len_max) break;                                                            // (:681)
            if (P.smooth != 0.f) {                                                                  // (:684-688)
                const float om = 1.f - P.smooth;
                const float tx = P.smooth * vx + om * nxv, ty = P.smooth * vy + om * nyv, tz = P.smooth * vz + om * nzv;
                const float nrm = (float)sqrt(((double)(tx * tx) + (double)(ty * ty)) + (double)(tz * tz));
                nxv = tx / nrm; nyv = ty / nrm; nzv = tz / nrm;
            }
            px = qx; py = qy; pz = qz; vx = nxv; vy = nyv; vz = nzv;
        }
    }
    if (!kWrite) nfb[line] = make_int2(nf, nb);
}

// ---- microscopy regime (stream_micro_new_point!, src/stream.jl:547-617): one WARP per line.  After the tentative step the
// next POSITION is the voxel of the (2 d1 + 1)(2 d2 + 1)(2 d3 + 1) search box -- inside the mask and inside the cone of the
// search angle around the current direction -- whose first vector is most similar to the current one: `argmax` over the box
// in column-major order (first maximum, NaN first; voxels outside the volume, the mask or the cone hold -Inf).  The unit vectors
// of the search area (:268-292) are recomputed per candidate in the reference's fp32 order; its centre is 0 / 0 = NaN, which
// fails every comparison, so the centre voxel always passes the cone test -- kept.
struct Cand { float val, cos; int idx; };
__device__ __forceinline__ bool cand_better(const Cand& a, const Cand& b) {       // does a come before b in Julia's argmax?
    const bool an = isnan(a.val), bn = isnan(b.val);
    if (an != bn) return an;
    if (an) return a.idx < b.idx;
    return a.val > b.val || (a.val == b.val && a.idx < b.idx);
}

template <bool kWrite>
__global__ void __launch_bounds__(128) stream_track_micro_kernel(TrackParams P, int2* __restrict__ nfb, const int64_t* __restrict__ off,
                                                                  const int32_t* __restrict__ sidx, const uint8_t* __restrict__ kept,
                                                                  float* __restrict__ xyz, int32_t* __restrict__ npts_out) {
    const int64_t line = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (line >= P.nseed * P.nsub) return;
    int nf_known = 0;
    float* out = nullptr;
    if (kWrite) {
        if (!kept[line]) return;
        const int2 c = nfb[line];
        nf_known = c.x;
        out = xyz + 3 * off[line];
        if (lane == 0) npts_out[sidx[line]] = c.x + c.y;
    }
    const int64_t si = line / P.nsub; const int isub = (int)(line - si * P.nsub);
    const int lin = P.seeds[si];
    const int sx = lin % P.nx, sy = (lin / P.nx) % P.ny, sz = lin / (P.nx * P.ny);
    const float s0 = P.sub[isub * 3 + 0], s1 = P.sub[isub * 3 + 1], s2 = P.sub[isub * 3 + 2];
    const int d0 = P.sd[0], d1 = P.sd[1], d2 = P.sd[2];
    const int w0 = 2 * d0 + 1, w1 = 2 * d1 + 1, nbox = w0 * w1 * (2 * d2 + 1);
    const float h0 = (float)d0 + 0.5f, h1 = (float)d1 + 0.5f, h2 = (float)d2 + 0.5f;
    int npts = 0, nf = 0, nb = 0;
    for (int dir = 0; dir < 2; ++dir) {
        const float fwd = dir == 0 ? 1.f : -1.f;
        float px = (float)(sx + 1) + s0, py = (float)(sy + 1) + s1, pz = (float)(sz + 1) + s2;
        const float* sv = P.ovec + (int64_t)lin * P.nvec * 3;          // first vector (W.ivec_next stays 1 in this regime)
        float vx = sv[0] * fwd, vy = sv[1] * fwd, vz = sv[2] * fwd;
        while (true) {
            const float qx = px + vx * P.step, qy = py + vy * P.step, qz = pz + vz * P.step;
            const int ix = __float2int_rn(qx), iy = __float2int_rn(qy), iz = __float2int_rn(qz);
            if (ix < 1 || ix > P.nx || iy < 1 || iy > P.ny || iz < 1 || iz > P.nz) break;
            if (!P.mask[(int64_t)(ix - 1) + (int64_t)P.nx * ((iy - 1) + (int64_t)P.ny * (iz - 1))]) break;
            Cand best{-INFINITY, -INFINITY, 0x7fffffff};
            for (int idx = lane; idx < nbox; idx += 32) {
                const int kx = idx % w0 - d0, ky = (idx / w0) % w1 - d1, kz = idx / (w0 * w1) - d2;
                const int x = ix + kx, y = iy + ky, z = iz + kz;
                Cand c{-INFINITY, -INFINITY, idx};
                if (x >= 1 && x <= P.nx && y >= 1 && y <= P.ny && z >= 1 && z <= P.nz) {
                    const int64_t nl = (int64_t)(x - 1) + (int64_t)P.nx * ((y - 1) + (int64_t)P.ny * (z - 1));
                    const float rx = (float)kx / h0, ry = (float)ky / h1, rz = (float)kz / h2;
                    const float r = sqrtf((rx * rx + ry * ry) + rz * rz);
                    float ax = 0.f, ay = 0.f, az = 0.f;
                    if (r < 1.f) { ax = rx / r; ay = ry / r; az = rz / r; }
                    const bool skip = !P.mask[nl] || (ax == 0.f && ay == 0.f && az == 0.f) || (dot3(vx, vy, vz, ax, ay, az) <= P.search_cos);
                    if (!skip) {
                        const float* ov = P.ovec + nl * P.nvec * 3;
                        c.cos = dot3(vx, vy, vz, ov[0], ov[1], ov[2]);
                        c.val = fabsf(c.cos);
                    }
                }
                if (cand_better(c, best)) best = c;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                Cand t{__shfl_xor_sync(0xffffffffu, best.val, o), __shfl_xor_sync(0xffffffffu, best.cos, o), __shfl_xor_sync(0xffffffffu, best.idx, o)};
                if (cand_better(t, best)) best = t;
            }
            if (!isfinite(best.cos)) break;
            const int kx = best.idx % w0 - d0, ky = (best.idx / w0) % w1 - d1, kz = best.idx / (w0 * w1) - d2;
            const int bxv = ix + kx, byv = iy + ky, bzv = iz + kz;
            const float* ov = P.ovec + ((int64_t)(bxv - 1) + (int64_t)P.nx * ((byv - 1) + (int64_t)P.ny * (bzv - 1))) * P.nvec * 3;
            float nxv = ov[0], nyv = ov[1], nzv = ov[2];
            if (!(best.cos > 0.f)) { nxv = -nxv; nyv = -nyv; nzv = -nzv; }
            if (kWrite && lane == 0) {
                float* o = out + 3 * (int64_t)(dir == 0 ? nf_known - 1 - nf : nf_known + nb);
                o[0] = px; o[1] = py; o[2] = pz;
            }
            if (dir == 0) ++nf; else ++nb;
            ++npts;
            if (dot3(vx, vy, vz, nxv, nyv, nzv) < P.cos_thresh) break;
            if (npts > P.len_max) break;
            if (P.smooth != 0.f) {
                const float om = 1.f - P.smooth;
                const float tx = P.smooth * vx + om * nxv, ty = P.smooth * vy + om * nyv, tz = P.smooth * vz + om * nzv;
                const float nrm = (float)sqrt(((double)(tx * tx) + (double)(ty * ty)) + (double)(tz * tz));
                nxv = tx / nrm; nyv = ty / nrm; nzv = tz / nrm;
            }
            px = (float)bxv; py = (float)byv; pz = (float)bzv;             // the position of the chosen voxel (:603-605)
            vx = nxv; vy = nyv; vz = nzv;
        }
    }
    if (!kWrite && lane == 0) nfb[line] = make_int2(nf, nb);
}

__global__ void stream_len_kernel(const int2* __restrict__ nfb, int64_t n, int len_min, int64_t* __restrict__ len, int32_t* __restrict__ keep32, uint8_t* __restrict__ kept) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = nfb[i].x + nfb[i].y;
    const bool k = t >= len_min;                                       // (:768)
    len[i] = k ? t : 0; keep32[i] = k ? 1 : 0; kept[i] = k ? 1 : 0;
}

struct DevBuf { std::vector<void*> p; ~DevBuf() { for (void* q : p) cudaFree(q); }
                template <class T> cudaError_t alloc(T** o, size_t n) { void* q = nullptr; cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)); if (e == cudaSuccess) p.push_back(q); *o = (T*)q; return e; } };

struct StreamResult { int device; int64_t nstr, npts; int32_t* d_npts; float* d_xyz; float* d_scal; };   // d_scal: one flag per point (LCM branch) or NULL
struct LcmArgs { const float* d_lcms; double thresh; int s1, s2; unsigned long long seed; };

}  // namespace
}  // namespace fibers

using namespace fibers;
#define T_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return fail(_e == cudaErrorMemoryAllocation ? FIBERS_ERR_NOMEM : FIBERS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)

// Device-resident core: every volume pointer is a DEVICE pointer (e.g. the peak / qa planes a reconstruction just wrote).
static int stream_core(const float* const* d_ovec, int nvec, int nx, int ny, int nz, const float* const* d_f, float f_thresh,
                       const float* d_fa, float fa_thresh, const uint8_t* d_mask, const uint8_t* d_seed,
                       const float* sublist /*host [nsub][3]*/, int nsub, int len_min, int len_max, float cosang_thresh,
                       float step_size, float smooth_coeff, const int32_t* micro_search_dist, float micro_search_cosang,
                       const LcmArgs* lcm, void** result, int64_t* nstr, int64_t* npts_total) {
    if (!d_ovec || !sublist || !result || !nstr || !npts_total) return fail(FIBERS_ERR_ARG, "NULL pointer");
    if (nvec < 1 || nvec > MAX_NVEC) return fail(FIBERS_ERR_ARG, "between 1 and 8 orientation-vector volumes are supported");
    if (nx <= 0 || ny <= 0 || nz <= 0 || nsub <= 0) return fail(FIBERS_ERR_ARG, "volume dimensions and the number of sub-voxel samples must be positive");
    const int64_t nvox = (int64_t)nx * ny * nz;
    if (nvox >= (1ll << 31)) return fail(FIBERS_ERR_ARG, "volume too large");
    *result = nullptr; *nstr = 0; *npts_total = 0;
    int device = 0;
    T_CUDA(cudaGetDevice(&device));
    DevBuf D;
    uint8_t* mask_arr; float* ovec_arr;
    T_CUDA(D.alloc(&mask_arr, (size_t)nvox)); T_CUDA(D.alloc(&ovec_arr, (size_t)nvox * nvec * 3));
    PackIn in{};
    for (int i = 0; i < nvec; ++i) { in.ovec[i] = d_ovec[i]; in.f[i] = d_f ? d_f[i] : nullptr; if (!in.ovec[i] || (d_f && !in.f[i])) return fail(FIBERS_ERR_ARG, "NULL volume pointer"); }
    const unsigned gv = (unsigned)((nvox + 255) / 256);
    stream_pack_kernel<<<gv, 256>>>(in, nvec, d_f ? 1 : 0, f_thresh, d_fa, fa_thresh, d_mask, nvox, mask_arr, ovec_arr);
    count_launch(1);
    T_CUDA(cudaGetLastError());
    // seed voxels in ascending (column-major) order: findall(W.mask .> 0) or findall(seed.vol .> 0)  (:744-754)
    int32_t* seeds; int* d_nseed;
    T_CUDA(D.alloc(&seeds, (size_t)nvox)); T_CUDA(D.alloc(&d_nseed, 1));
    const uint8_t* flags = d_seed ? d_seed : mask_arr;
    {
        size_t tb = 0;
        thrust::counting_iterator<int32_t> cnt(0);
        T_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, cnt, flags, seeds, d_nseed, (int)nvox));
        uint8_t* tmp; T_CUDA(D.alloc(&tmp, tb));
        T_CUDA(cub::DeviceSelect::Flagged(tmp, tb, cnt, flags, seeds, d_nseed, (int)nvox));
        count_launch(1);
    }
    int nseed = 0;
    T_CUDA(cudaMemcpy(&nseed, d_nseed, sizeof(int), cudaMemcpyDeviceToHost));
    const int64_t nline = (int64_t)nseed * nsub;
    if (nline == 0) return 0;
    float* d_sub; T_CUDA(D.alloc(&d_sub, (size_t)nsub * 3));
    T_CUDA(cudaMemcpy(d_sub, sublist, sizeof(float) * 3 * nsub, cudaMemcpyHostToDevice));
    int2* nfb; int64_t *len, *off; int32_t *keep32, *sidx; uint8_t* kept;
    T_CUDA(D.alloc(&nfb, (size_t)nline)); T_CUDA(D.alloc(&len, (size_t)nline + 1)); T_CUDA(D.alloc(&off, (size_t)nline + 1));
    T_CUDA(D.alloc(&keep32, (size_t)nline + 1)); T_CUDA(D.alloc(&sidx, (size_t)nline + 1)); T_CUDA(D.alloc(&kept, (size_t)nline));
    TrackParams P{ovec_arr, mask_arr, seeds, nseed, d_sub, nsub, nx, ny, nz, nvec, len_min, len_max, cosang_thresh, step_size, smooth_coeff, {0, 0, 0}, 0.f,
                  nullptr, 0, 1, 0ull};
    const bool micro = micro_search_dist != nullptr;
    if (lcm) {
        if (micro) return fail(FIBERS_ERR_ARG, "stream: local connection matrices are only defined for the macroscopic regime");
        if (!lcm->d_lcms || lcm->s1 < 0 || lcm->s1 > 2 || lcm->s2 < 0 || lcm->s2 > 2 || lcm->s1 == lcm->s2) return fail(FIBERS_ERR_ARG, "stream: bad LCM arguments");
        float* lcm_arr; T_CUDA(D.alloc(&lcm_arr, (size_t)nvox * 10));
        stream_pack_lcm_kernel<<<gv, 256>>>(lcm->d_lcms, lcm->thresh, nvox, lcm_arr);
        count_launch(1);
        T_CUDA(cudaGetLastError());
        P.lcm = lcm_arr; P.s1 = lcm->s1; P.s2 = lcm->s2; P.lcm_seed = lcm->seed;
    }
    if (micro) {
        for (int i = 0; i < 3; ++i) { if (micro_search_dist[i] < 0 || micro_search_dist[i] > 64) return fail(FIBERS_ERR_ARG, "micro_search_dist must be in 0..64"); P.sd[i] = micro_search_dist[i]; }
        P.search_cos = micro_search_cosang;
    }
    const unsigned gl = (unsigned)(micro ? (nline * 32 + 127) / 128 : (nline + 127) / 128);          // micro: one warp per line
    if (micro) stream_track_micro_kernel<false><<<gl, 128>>>(P, nfb, nullptr, nullptr, nullptr, nullptr, nullptr);
    else if (lcm) stream_track_kernel<false, true><<<gl, 128>>>(P, nfb, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    else stream_track_kernel<false, false><<<gl, 128>>>(P, nfb, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    stream_len_kernel<<<(unsigned)((nline + 255) / 256), 256>>>(nfb, nline, len_min, len, keep32, kept);
    count_launch(2);
    T_CUDA(cudaMemsetAsync(len + nline, 0, sizeof(int64_t))); T_CUDA(cudaMemsetAsync(keep32 + nline, 0, sizeof(int32_t)));
    {
        size_t t1 = 0, t2 = 0;
        T_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t1, len, off, (int)(nline + 1)));
        T_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t2, keep32, sidx, (int)(nline + 1)));
        uint8_t* tmp; T_CUDA(D.alloc(&tmp, std::max(t1, t2)));
        T_CUDA(cub::DeviceScan::ExclusiveSum(tmp, t1, len, off, (int)(nline + 1)));
        T_CUDA(cub::DeviceScan::ExclusiveSum(tmp, t2, keep32, sidx, (int)(nline + 1)));
        count_launch(2);
    }
    int64_t total = 0; int32_t nkeep = 0;
    T_CUDA(cudaMemcpy(&total, off + nline, sizeof(int64_t), cudaMemcpyDeviceToHost));
    T_CUDA(cudaMemcpy(&nkeep, sidx + nline, sizeof(int32_t), cudaMemcpyDeviceToHost));
    StreamResult* R = new StreamResult{device, nkeep, total, nullptr, nullptr, nullptr};
    if (cudaMalloc(&R->d_npts, sizeof(int32_t) * std::max<int64_t>(nkeep, 1)) != cudaSuccess ||
        cudaMalloc(&R->d_xyz, sizeof(float) * 3 * std::max<int64_t>(total, 1)) != cudaSuccess ||
        (lcm && cudaMalloc(&R->d_scal, sizeof(float) * std::max<int64_t>(total, 1)) != cudaSuccess)) {
        cudaFree(R->d_npts); cudaFree(R->d_xyz); cudaFree(R->d_scal); delete R; cudaGetLastError();
        return fail(FIBERS_ERR_NOMEM, "stream: device allocation of the streamline buffers failed");
    }
    if (nkeep > 0) {
        if (micro) stream_track_micro_kernel<true><<<gl, 128>>>(P, nfb, off, sidx, kept, R->d_xyz, R->d_npts);
        else if (lcm) stream_track_kernel<true, true><<<gl, 128>>>(P, nfb, off, sidx, kept, R->d_xyz, R->d_npts, R->d_scal);
        else stream_track_kernel<true, false><<<gl, 128>>>(P, nfb, off, sidx, kept, R->d_xyz, R->d_npts, nullptr);
        count_launch(1);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(R->d_npts); cudaFree(R->d_xyz); cudaFree(R->d_scal); delete R; return fail(FIBERS_ERR_CUDA, std::string("stream: ") + cudaGetErrorString(e)); }
    *result = R; *nstr = nkeep; *npts_total = total;
    return 0;
}

extern "C" int fibers_stream_device(const float* const* d_ovec, int nvec, int nx, int ny, int nz, const float* const* d_f, float f_thresh,
                                    const float* d_fa, float fa_thresh, const uint8_t* d_mask, const uint8_t* d_seed,
                                    const float* sublist /*host [nsub][3]*/, int nsub, int len_min, int len_max, float cosang_thresh,
                                    float step_size, float smooth_coeff, const int32_t* micro_search_dist, float micro_search_cosang,
                                    void** result, int64_t* nstr, int64_t* npts_total) {
    return stream_core(d_ovec, nvec, nx, ny, nz, d_f, f_thresh, d_fa, fa_thresh, d_mask, d_seed, sublist, nsub, len_min, len_max, cosang_thresh,
                       step_size, smooth_coeff, micro_search_dist, micro_search_cosang, nullptr, result, nstr, npts_total);
}

static int stream_host(const float* const* ovec, int nvec, int nx, int ny, int nz, const float* const* f, float f_thresh,
                       const float* fa, float fa_thresh, const uint8_t* mask, const uint8_t* seed, const float* sublist, int nsub,
                       int len_min, int len_max, float cosang_thresh, float step_size, float smooth_coeff,
                       const int32_t* micro_search_dist, float micro_search_cosang, const float* lcms, LcmArgs lcm, int device,
                       void** result, int64_t* nstr, int64_t* npts_total) {
    if (!ovec || !sublist || !result || !nstr || !npts_total) return fail(FIBERS_ERR_ARG, "NULL pointer");
    if (nvec < 1 || nvec > MAX_NVEC) return fail(FIBERS_ERR_ARG, "between 1 and 8 orientation-vector volumes are supported");
    if (nx <= 0 || ny <= 0 || nz <= 0) return fail(FIBERS_ERR_ARG, "volume dimensions must be positive");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return fail(FIBERS_ERR_NODEV, "no CUDA device available (libfibers_cuda has no CPU fallback)"); }
    if (device < 0 || device >= n) return fail(FIBERS_ERR_ARG, "device ordinal out of range");
    T_CUDA(cudaSetDevice(device));
    const int64_t nvox = (int64_t)nx * ny * nz;
    DevBuf D;
    const float* d_ovec[MAX_NVEC]; const float* d_f[MAX_NVEC];
    auto up = [&](const void* h, size_t bytes, const void** d) -> int {
        uint8_t* q; T_CUDA(D.alloc(&q, bytes)); T_CUDA(cudaMemcpy(q, h, bytes, cudaMemcpyHostToDevice)); *d = q; return 0;
    };
    for (int i = 0; i < nvec; ++i) {
        if (!ovec[i] || (f && !f[i])) return fail(FIBERS_ERR_ARG, "NULL volume pointer");
        if (int rc = up(ovec[i], sizeof(float) * 3 * nvox, (const void**)&d_ovec[i])) return rc;
        if (f) if (int rc = up(f[i], sizeof(float) * nvox, (const void**)&d_f[i])) return rc;
    }
    const float* d_fa = nullptr; const uint8_t *d_mask = nullptr, *d_seed = nullptr;
    if (fa) if (int rc = up(fa, sizeof(float) * nvox, (const void**)&d_fa)) return rc;
    if (mask) if (int rc = up(mask, (size_t)nvox, (const void**)&d_mask)) return rc;
    if (seed) if (int rc = up(seed, (size_t)nvox, (const void**)&d_seed)) return rc;
    if (lcms) if (int rc = up(lcms, sizeof(float) * 10 * nvox, (const void**)&lcm.d_lcms)) return rc;
    return stream_core(d_ovec, nvec, nx, ny, nz, f ? d_f : nullptr, f_thresh, d_fa, fa_thresh, d_mask, d_seed, sublist, nsub,
                       len_min, len_max, cosang_thresh, step_size, smooth_coeff, micro_search_dist, micro_search_cosang, lcms ? &lcm : nullptr,
                       result, nstr, npts_total);
}

extern "C" int fibers_stream(const float* const* ovec, int nvec, int nx, int ny, int nz, const float* const* f, float f_thresh,
                             const float* fa, float fa_thresh, const uint8_t* mask, const uint8_t* seed, const float* sublist, int nsub,
                             int len_min, int len_max, float cosang_thresh, float step_size, float smooth_coeff,
                             const int32_t* micro_search_dist, float micro_search_cosang, int device,
                             void** result, int64_t* nstr, int64_t* npts_total) {
    return stream_host(ovec, nvec, nx, ny, nz, f, f_thresh, fa, fa_thresh, mask, seed, sublist, nsub, len_min, len_max, cosang_thresh, step_size,
                       smooth_coeff, micro_search_dist, micro_search_cosang, nullptr, LcmArgs{}, device, result, nstr, npts_total);
}

// stream(...; lcms, lcm_thresh): the branch of stream_new_point! that follows local connection matrices (src/stream.jl:523-538, :380-494).
// lcms: [nx, ny, nz, 10] like lcms.vol; strdim1 / strdim2: the in-plane dimensions (0-based; the reference derives them from the
// all-zero component of the first orientation volume, :224-226); lcm_seed: seed of the counter-based uniform generator.
extern "C" int fibers_stream_lcm(const float* const* ovec, int nvec, int nx, int ny, int nz, const float* const* f, float f_thresh,
                                 const float* fa, float fa_thresh, const uint8_t* mask, const uint8_t* seed, const float* sublist, int nsub,
                                 int len_min, int len_max, float step_size, float smooth_coeff, const float* lcms, double lcm_thresh,
                                 int strdim1, int strdim2, uint64_t lcm_seed, int device, void** result, int64_t* nstr, int64_t* npts_total) {
    if (!lcms) return fail(FIBERS_ERR_ARG, "NULL pointer");
    return stream_host(ovec, nvec, nx, ny, nz, f, f_thresh, fa, fa_thresh, mask, seed, sublist, nsub, len_min, len_max, 0.f, step_size,
                       smooth_coeff, nullptr, 0.f, lcms, LcmArgs{nullptr, lcm_thresh, strdim1, strdim2, (unsigned long long)lcm_seed}, device, result, nstr, npts_total);
}

extern "C" int fibers_stream_fetch(void* result, int32_t* npts, float* xyz) {
    StreamResult* R = (StreamResult*)result;
    if (!R) return fail(FIBERS_ERR_ARG, "NULL result handle");
    T_CUDA(cudaSetDevice(R->device));
    if (npts && R->nstr > 0) T_CUDA(cudaMemcpy(npts, R->d_npts, sizeof(int32_t) * R->nstr, cudaMemcpyDeviceToHost));
    if (xyz && R->npts > 0) T_CUDA(cudaMemcpy(xyz, R->d_xyz, sizeof(float) * 3 * R->npts, cudaMemcpyDeviceToHost));
    return 0;
}

// one value per point in the order of fibers_stream_fetch's xyz: 1 where the LCM pick differed from the conventional pick
// (the scalars of the reference's Tract, src/stream.jl:783); an error for results of the other entry points
extern "C" int fibers_stream_fetch_scalars(void* result, float* scalars) {
    StreamResult* R = (StreamResult*)result;
    if (!R || !scalars) return fail(FIBERS_ERR_ARG, "NULL pointer");
    if (!R->d_scal) return fail(FIBERS_ERR_ARG, "this result carries no scalars (not an LCM run)");
    T_CUDA(cudaSetDevice(R->device));
    if (R->npts > 0) T_CUDA(cudaMemcpy(scalars, R->d_scal, sizeof(float) * R->npts, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" void fibers_stream_free(void* result) {
    StreamResult* R = (StreamResult*)result;
    if (!R) return;
    cudaSetDevice(R->device);
    cudaFree(R->d_npts); cudaFree(R->d_xyz); cudaFree(R->d_scal);
    delete R;
}
