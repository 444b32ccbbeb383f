// RUMBA-SD (robust and unbiased model-based spherical deconvolution) on the GPU: SURVEY.md section 8(f) rank 2.
// Replaces rumba_rec of the reference (src/rusd.jl:419-636): set-up :444-540, the Richardson-Lucy iteration
// rumba_sd_iterate! :266-340 with the total-variation term rumba_tv! :214-234 (sd_grad! :181-186, sd_div! :192-205),
// the Bessel ratio :167-175, and the final normalisation, GFA and peak search (:560-633, rumba_peaks! :348-373).
//
// Data layout: the work matrices keep the reference's shape, [rows x nmask] column-major: a mask voxel is a COLUMN
// (its ndir or ncomp values are contiguous).  Element-wise passes and the per-voxel reductions are coalesced along
// a column; the TV stencil of component c reads the same component of the neighbouring voxels' columns, so a warp
// that owns 32 consecutive components of one voxel reads 128 contiguous bytes per stencil point.
//
// Per iteration: 2 matrix products K' x [signal .* Iratio | dodf] (library SGEMM: plain products, cuBLAS is loaded at
// run time), ONE fused kernel for rl / (rl2 + eps), the TV term (13-point stencil on the previous estimate, computed on
// the fly: no gradient / divergence volumes exist) and the multiplicative update, 1 product K x fodf, and ONE fused
// warp-per-voxel kernel for dodf_sig, the noise-variance estimate (sum over directions), its clamp and the NEXT
// iteration's Bessel ratio and signal .* Iratio; a two-stage deterministic reduction gives mean(sigma^2) for the
// regularisation weight without a host synchronisation.
//
// The TV term couples neighbouring voxels in every iteration, so this path does NOT shard by z-slab without a halo
// exchange per iteration: it runs on one GPU ("replicas only"; a batch of subjects uses one GPU per subject).
#include <dlfcn.h>
#include <math_constants.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace fibers {
namespace {

constexpr float EPS32 = 1.1920929e-7f;          // eps(Float32)
constexpr int NPEAK = 5;
constexpr int NB_W = 16;                        // angular-neighbourhood table width (12.5 deg on sphere_724: <= 8 entries)

// ---- cuBLAS, loaded lazily (the library itself does not link against it) -------------------------------------------
typedef int (*cublasCreate_t)(void**);
typedef int (*cublasDestroy_t)(void*);
typedef int (*cublasSetStream_t)(void*, cudaStream_t);
typedef int (*cublasSgemm_t)(void*, int, int, int, int, int, const float*, const float*, int, const float*, int, const float*, float*, int);
struct Cublas {
    void* lib = nullptr; cublasCreate_t create = nullptr; cublasDestroy_t destroy = nullptr; cublasSetStream_t set_stream = nullptr;
    cublasSgemm_t sgemm = nullptr; bool tried = false;
};
Cublas g_blas; std::mutex g_blas_mu;
bool load_cublas() {
    std::lock_guard<std::mutex> lk(g_blas_mu);
    if (g_blas.tried) return g_blas.sgemm != nullptr;
    g_blas.tried = true;
    for (const char* n : {"libcublas.so.12", "libcublas.so", "/usr/local/cuda/lib64/libcublas.so.12"}) {
        g_blas.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_blas.lib) break;
    }
    if (!g_blas.lib) return false;
    g_blas.create = (cublasCreate_t)dlsym(g_blas.lib, "cublasCreate_v2");
    g_blas.destroy = (cublasDestroy_t)dlsym(g_blas.lib, "cublasDestroy_v2");
    g_blas.set_stream = (cublasSetStream_t)dlsym(g_blas.lib, "cublasSetStream_v2");
    g_blas.sgemm = (cublasSgemm_t)dlsym(g_blas.lib, "cublasSgemm_v2");
    if (!g_blas.create || !g_blas.destroy || !g_blas.set_stream || !g_blas.sgemm) { g_blas.sgemm = nullptr; return false; }
    return true;
}

// I_nu(z) / I_{nu-1}(z), Perron's continued fraction, the reference's expression term by term (src/rusd.jl:169-174)
__device__ __forceinline__ float besseli_ratio(float n2 /* 2*nu */, float z) {
    return z / ((n2 + z) - ((n2 + 1.f) * z / (2.f * z + (n2 + 1.f) - ((n2 + 3.f) * z / ((n2 + 2.f) + 2.f * z - ((n2 + 5.f) * z / ((n2 + 3.f) + 2.f * z)))))));
}

// ---- set-up: signal_mat [ndir x nmask] from the DWI volume (src/rusd.jl:448-466), warp per mask voxel ---------------
// vol_row[k]: 0 for a minimum-b volume, else the 1-based row of volume k (rows 1 .. ndir-1 in volume order)
__global__ void rumba_signal_kernel(const float* __restrict__ dwi, int64_t pitch, int nvol, const int* __restrict__ vol_row, int nb0,
                                    const int* __restrict__ ind, int nmask, int ndir, float* __restrict__ sig) {
    const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (col >= nmask) return;
    const int64_t v = ind[col];
    float s0 = 0.f;                                            // mean over the minimum-b volumes of max(s, 0), in volume order
    for (int k = 0; k < nvol; ++k) if (vol_row[k] == 0) s0 += fmaxf(dwi[(int64_t)k * pitch + v], 0.f);
    s0 /= (float)nb0;
    float* out = sig + (int64_t)col * ndir;
    for (int k = lane; k < nvol; k += 32) {
        const int r = vol_row[k];
        if (r == 0) continue;
        float x = fmaxf(dwi[(int64_t)k * pitch + v], 0.f) / s0;      // x / 0 = Inf, 0 / 0 = NaN
        if (x != x) x = 0.f;                                         // NaN -> 0 (:461)
        if (x > 1.f) x = 1.f;                                        // (:465; Inf -> 1)
        out[r] = x;
    }
    if (lane == 0) out[0] = s0 > 0.f ? 1.f : 0.f;                    // signal = 1 if b = 0 (:464)
}

// ---- initial state (rumba_sd_initialize!, :240-255) + the first Bessel ratio ------------------------------------------
__global__ void rumba_init_kernel(const float* __restrict__ sig, const float* __restrict__ kf0 /*[ndir] = K f0*/, float f0, float s20, float n2,
                                  int ndir, int ncomp, int nmask, float* __restrict__ fodf, float* __restrict__ dodf, float* __restrict__ ir,
                                  float* __restrict__ t1, float* __restrict__ s2) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)ncomp * nmask) fodf[i] = f0;
    if (i < nmask) s2[i] = s20;
    if (i < (int64_t)ndir * nmask) {
        const int r = (int)(i % ndir);
        const float d = kf0[r], s = sig[i];
        dodf[i] = d;
        const float dsig = (s * d) / s20;
        const float q = besseli_ratio(n2, dsig);
        ir[i] = q; t1[i] = s * q;
    }
}

// ---- fused update: rl / (rl2 + eps), TV term on the previous estimate, fodf = max(fodf * rl * tv, 0) -------------------
// One block per mask voxel, threads over the ncomp components.  The zero-embedded volume of component c is
// V_c(q) = fodf_old[col(q)][c] inside the mask, 0 elsewhere; gradient G(q) = forward differences (0 at the far border),
// normalised by sqrt(|G|^2 + eps); divergence = backward differences of the normalised gradient with the reference's
// border rules (Div_x(p) = Gx(p) - Gx(p - x) with Gx(-1) = 0; at the far border Gx(p) = 0); tv = 1 / (|1 - lambda div| + eps).
struct TvGeom { int nx, ny, nz; };
__global__ void __launch_bounds__(384) rumba_update_kernel(const float* __restrict__ fold, float* __restrict__ fnew, const float* __restrict__ rl,
                                                           const float* __restrict__ rl2, const int* __restrict__ ind, const int* __restrict__ colmap,
                                                           TvGeom g, int ncomp, int use_tv, const float* __restrict__ lam_scalar,
                                                           const float* __restrict__ lam_col /* per-column lambda (ipat > 1) or null */) {
    const int col = blockIdx.x;
    __shared__ int nb[13];          // columns of: p, p+x, p+y, p+z, p-x, p-x+y, p-x+z, p-y, p-y+x, p-y+z, p-z, p-z+x, p-z+y  (-1: outside / zero)
    __shared__ int has[4];          // p-x, p-y, p-z exist
    if (use_tv && threadIdx.x < 13) {
        const int64_t v = ind[col];
        const int x = (int)(v % g.nx), y = (int)((v / g.nx) % g.ny), z = (int)(v / ((int64_t)g.nx * g.ny));
        const int dx[13] = {0, 1, 0, 0, -1, -1, -1, 0, 1, 0, 0, 1, 0};
        const int dy[13] = {0, 0, 1, 0, 0, 1, 0, -1, -1, -1, 0, 0, 1};
        const int dz[13] = {0, 0, 0, 1, 0, 0, 1, 0, 0, 1, -1, -1, -1};
        const int t = threadIdx.x;
        const int xx = x + dx[t], yy = y + dy[t], zz = z + dz[t];
        int c = -1;
        if (xx >= 0 && xx < g.nx && yy >= 0 && yy < g.ny && zz >= 0 && zz < g.nz) c = colmap[((int64_t)zz * g.ny + yy) * g.nx + xx];
        nb[t] = c;
        if (t == 0) { has[0] = x > 0; has[1] = y > 0; has[2] = z > 0; has[3] = (x < g.nx - 1) | ((y < g.ny - 1) << 1) | ((z < g.nz - 1) << 2); }
    }
    __syncthreads();
    const float lam = use_tv ? (lam_col ? lam_col[col] : *lam_scalar) : 0.f;
    for (int c = threadIdx.x; c < ncomp; c += blockDim.x) {
        const int64_t i = (int64_t)col * ncomp + c;
        const float f = fold[i];
        float tv = 1.f;
        if (use_tv) {
            auto V = [&](int k) { const int cc = nb[k]; return cc >= 0 ? fold[(int64_t)cc * ncomp + c] : 0.f; };
            // forward differences exist only where the +neighbour is inside the volume (else 0: V[end] - V[end])
            auto grad = [&](float v0, float vx, float vy, float vz, bool ex, bool ey, bool ez, float& gx, float& gy, float& gz) {
                gx = ex ? vx - v0 : 0.f; gy = ey ? vy - v0 : 0.f; gz = ez ? vz - v0 : 0.f;
                const float n = sqrtf(gx * gx + gy * gy + gz * gz + EPS32);
                gx /= n; gy /= n; gz /= n;
            };
            const int far = has[3];
            float gx, gy, gz, t0, t1, t2;
            grad(f, V(1), V(2), V(3), far & 1, far & 2, far & 4, gx, gy, gz);                       // at p
            float div = gx;
            if (has[0]) { grad(V(4), f, V(5), V(6), true, far & 2, far & 4, t0, t1, t2); div -= t0; }   // at p - x
            float dy = gy;
            if (has[1]) { grad(V(7), V(8), f, V(9), far & 1, true, far & 4, t0, t1, t2); dy -= t1; }    // at p - y
            float dz = gz;
            if (has[2]) { grad(V(10), V(11), V(12), f, far & 1, far & 2, true, t0, t1, t2); dz -= t2; } // at p - z
            div = (div + dy) + dz;
            tv = 1.f / (fabsf(1.f - lam * div) + EPS32);
        }
        const float r = rl[i] / (rl2[i] + EPS32);
        fnew[i] = fmaxf((f * r) * tv, 0.f);
    }
}

// ---- fused noise step, warp per voxel: dodf_sig, sigma^2 (sum over directions, clamp), next Bessel ratio and signal .* Iratio
__global__ void rumba_noise_kernel(const float* __restrict__ sig, const float* __restrict__ dodf, float* __restrict__ ir, float* __restrict__ t1,
                                   float* __restrict__ s2, int ndir, int nmask, float n2, float norder, float* __restrict__ block_sum) {
    const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    float s2new = 0.f;
    if (col < nmask) {
        const float s2old = s2[col];
        const int64_t base = (int64_t)col * ndir;
        float acc = 0.f;
        for (int r = lane; r < ndir; r += 32) {
            const float s = sig[base + r], d = dodf[base + r];
            const float dsig = (s * d) / s2old;                                  // dodf_sig of the NEXT iteration (old sigma^2, :309)
            acc += (s * s + d * d) / 2.f - (s2old * dsig) * ir[base + r];      // (:312-313)
            const float q = besseli_ratio(n2, dsig);
            ir[base + r] = q; t1[base + r] = s * q;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        s2new = acc / (norder * (float)ndir);
        s2new = fminf(fmaxf(s2new, 1.5625e-4f), 0.015625f);      // clamp to [(1/80)^2, (1/8)^2] (:318)
        if (lane == 0) s2[col] = s2new;
    }
    // deterministic partial sum of sigma^2 over the block's voxels (for lambda = max(mean sigma^2, (1/30)^2), :327)
    __shared__ float part[32];
    if (lane == 0) part[threadIdx.x >> 5] = col < nmask ? s2new : 0.f;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w]; block_sum[blockIdx.x] = t; }
}

__global__ void rumba_lambda_kernel(const float* __restrict__ block_sum, int nblocks, int nmask, float* __restrict__ lam) {
    __shared__ double sh[256];
    double t = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) t += (double)block_sum[i];
    sh[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) *lam = fmaxf((float)(sh[0] / nmask), 0.0011111111111111111f);          // max(mean sigma^2, (1/30)^2)
}

// ---- final: normalise, embed, GFA, peaks (src/rusd.jl:560-633), warp per voxel of the WHOLE volume --------------------
struct RumbaOut { float* fodf; float* fgm; float* fcsf; float* peak[NPEAK]; float* gfa; float* var; int16_t* peak_idx; };
__global__ void rumba_final_kernel(const float* __restrict__ fodf_mat, const float* __restrict__ s2, const int* __restrict__ colmap,
                                   const uint8_t* __restrict__ mask_any, int64_t nvox, int ncomp, const uint16_t* __restrict__ nbr /*[nvert][NB_W]*/,
                                   const float* __restrict__ vert /*[nvert][3]*/, RumbaOut o) {
    extern __shared__ float sm[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31, nvert = ncomp - 2;
    float* f = sm + (size_t)wib * (nvert + 2);
    const int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    if (v >= nvox) return;
    const int col = colmap[v];
    float fiso = 0.f;
    if (col >= 0) {
        const float* src = fodf_mat + (int64_t)col * ncomp;
        float sum = 0.f;
        for (int c = lane; c < ncomp; c += 32) sum += src[c];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
        const float inv = sum + EPS32;                                           // energy preservation (:561)
        const float fcsf = src[nvert] / inv, fgm = src[nvert + 1] / inv;
        fiso = fgm + fcsf;
        float tot = 0.f;
        for (int c = lane; c < nvert; c += 32) { const float x = src[c] / inv + fiso; f[c] = x; tot += x; }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, s);
        for (int c = lane; c < nvert; c += 32) { float x = f[c] / tot; if (x != x) x = 0.f; f[c] = x; }     // (:589-590)
        if (lane == 0) { o.fcsf[v] = fcsf; o.fgm[v] = fgm; o.var[v] = s2[col]; }
    } else {
        for (int c = lane; c < nvert; c += 32) f[c] = 0.f;                         // 0 / 0 -> NaN -> 0 outside the mask
        if (lane == 0) { o.fcsf[v] = 0.f; o.fgm[v] = 0.f; o.var[v] = 0.f; }
    }
    __syncwarp();
    // fodf frames + GFA = std (corrected) / sqrt(mean of squares)  (:598-599)
    float s1 = 0.f, sq = 0.f, mx = 0.f;
    for (int c = lane; c < nvert; c += 32) { const float x = f[c]; o.fodf[(int64_t)c * nvox + v] = x; s1 += x; sq += x * x; mx = fmaxf(mx, x); }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, s); sq += __shfl_xor_sync(0xffffffffu, sq, s); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s)); }
    const float mean = s1 / nvert;
    float dev = 0.f;
    for (int c = lane; c < nvert; c += 32) { const float d = f[c] - mean; dev += d * d; }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) dev += __shfl_xor_sync(0xffffffffu, dev, s);
    float gfa = sqrtf(dev / (nvert - 1)) / sqrtf(sq / nvert);
    if (gfa != gfa) gfa = 0.f;
    if (lane == 0) o.gfa[v] = gfa;
    // peaks: local maxima over the angular neighbourhood, amplitude >= thr / (1 - f_iso) * max, 5 largest, stable order
    float pv[NPEAK]; int pi[NPEAK];
    int npk = 0;
    if (mask_any[v]) {
        const float thr_abs = (0.1f / (1.f - fiso)) * mx;
        // every lane collects its candidates' best (value, index); 5 rounds of warp arg-max with "already taken" exclusion
        unsigned long long taken[6] = {0, 0, 0, 0, 0, 0};                        // nvert <= 384 bits
        for (int k = 0; k < NPEAK; ++k) {
            float best = 0.f; int bi = 0x7fffffff;
            for (int c = lane; c < nvert; c += 32) {
                if ((taken[c >> 6] >> (c & 63)) & 1ull) continue;
                const float x = f[c];
                if (!(x > 0.f) || x < thr_abs) continue;
                bool ok = true;
                float nmax = -CUDART_INF_F;
                for (int j = 0; j < NB_W; ++j) { const uint16_t n = nbr[c * NB_W + j]; if (n == 0xFFFF) break; nmax = fmaxf(nmax, f[n]); }
                ok = x > nmax;                                                   // fodf <= max(neighbours) -> not a peak (:361-364)
                if (ok && (x > best || (x == best && c < bi))) { best = x; bi = c; }
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, s); const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (!(best > 0.f)) break;
            pv[k] = best; pi[k] = bi; ++npk;
            taken[bi >> 6] |= 1ull << (bi & 63);
        }
    }
    float psum = 0.f;
    for (int k = 0; k < npk; ++k) psum += pv[k];
    const float fnorm = (1.f - fiso) / psum;
    if (lane < 3) {
        for (int k = 0; k < NPEAK; ++k)
            o.peak[k][(int64_t)lane * nvox + v] = k < npk ? vert[pi[k] * 3 + lane] * (pv[k] * fnorm) : 0.f;
    }
    if (o.peak_idx && lane == 0) for (int k = 0; k < NPEAK; ++k) o.peak_idx[(int64_t)k * nvox + v] = k < npk ? (int16_t)pi[k] : (int16_t)-1;
}

struct DevBuf {                        // frees what it owns on scope exit
    std::vector<void*> p;
    template <typename T> cudaError_t alloc(T** out, size_t n) { void* q = nullptr; cudaError_t e = cudaMalloc(&q, n * sizeof(T)); if (e == cudaSuccess) p.push_back(q); *out = (T*)q; return e; }
    ~DevBuf() { for (void* q : p) cudaFree(q); }
};

void ang2rot(double phi, double th, double R[3][3]) {
    const double cz = cos(phi), sz = sin(phi), cy = cos(th), sy = sin(th);
    const double Rz[3][3] = {{cz, -sz, 0}, {sz, cz, 0}, {0, 0, 1}}, Ry[3][3] = {{cy, 0, sy}, {0, 1, 0}, {-sy, 0, cy}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double t = 0; for (int k = 0; k < 3; ++k) t += Rz[i][k] * Ry[k][j]; R[i][j] = t; }
}

}  // namespace

// Builds Kernel [ndir x ncomp] column-major (src/rusd.jl:497-520) in float64, rounded to float32; vol_row as in rumba_signal_kernel.
std::string build_rumba_kernel(int nvol, const float* bval, const float* bvec /*[nvol x 3] column-major*/, const float* vertices, int nvert2,
                               float lpar, float lperp, float lcsf, float lgm, std::vector<float>& K, std::vector<int>& vol_row, int& ndir, int& nb0) {
    if (nvol <= 0 || !bval) return "Missing b-value table from input DWI structure";
    if (!bvec) return "Missing gradient table from input DWI structure";
    if (nvert2 <= 0 || nvert2 % 2) return "odf_dirs.vertices must have an even, positive number of rows";
    float bmin = bval[0];
    for (int k = 1; k < nvol; ++k) bmin = std::min(bmin, bval[k]);
    vol_row.assign(nvol, 0); ndir = 1; nb0 = 0;
    for (int k = 0; k < nvol; ++k) { if (bval[k] == bmin) ++nb0; else vol_row[k] = ndir++; }
    const int nvert = nvert2 / 2, ncomp = nvert + 2;
    std::vector<double> b(ndir, 0.0), g((size_t)ndir * 3, 0.0);
    for (int k = 0; k < nvol; ++k) {
        const int r = vol_row[k];
        if (!r) continue;
        const double gx = bvec[k], gy = bvec[nvol + k], gz = bvec[2 * nvol + k], n = sqrt(gx * gx + gy * gy + gz * gz);
        b[r] = bval[k]; g[r * 3 + 0] = gx / n; g[r * 3 + 1] = gy / n; g[r * 3 + 2] = gz / n;      // (a zero vector gives NaN, as in the reference)
    }
    K.assign((size_t)ndir * ncomp, 0.f);
    auto column = [&](int c, double phi, double th, double l1, double l2, double l3) {
        double R[3][3]; ang2rot(phi, th, R);
        double D[3][3]; const double lam[3] = {l1, l2, l3};
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double t = 0; for (int k = 0; k < 3; ++k) t += R[i][k] * lam[k] * R[j][k]; D[i][j] = t; }
        for (int r = 0; r < ndir; ++r) {
            double q = 0;
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) q += g[r * 3 + i] * D[i][j] * g[r * 3 + j];
            K[(size_t)c * ndir + r] = (float)exp(-b[r] * q);
        }
    };
    for (int i = 0; i < nvert; ++i) {
        const double x = vertices[nvert + i], y = vertices[nvert2 + nvert + i], z = vertices[2 * nvert2 + nvert + i];   // second half of the vertices (:507-509)
        const double hxy = hypot(x, y);
        column(i, atan2(y, x), -atan2(z, hxy), lpar, lperp, lperp);
    }
    column(nvert, 0, 0, lcsf, lcsf, lcsf);
    column(nvert + 1, 0, 0, lgm, lgm, lgm);
    return "";
}

// Angular neighbourhoods (src/rusd.jl:478-492): folded angle < ang_neig degrees, no self; table [nvert][NB_W], 0xFFFF terminated.
std::string build_rumba_neighbours(const float* vertices, int nvert2, float ang_neig, std::vector<uint16_t>& nbr) {
    const int nvert = nvert2 / 2;
    nbr.assign((size_t)nvert * NB_W, 0xFFFF);
    for (int i = 0; i < nvert; ++i) {
        int n = 0;
        for (int j = 0; j < nvert; ++j) {
            if (i == j) continue;
            float c = 0.f;
            for (int d = 0; d < 3; ++d) c += vertices[d * nvert2 + i] * vertices[d * nvert2 + j];
            c = std::min(1.f, std::max(-1.f, c));
            float a = acosf(c) * 57.29577951308232f;
            a = std::min(a, 180.f - a);
            if (a < ang_neig) { if (n >= NB_W) return "angular neighbourhood too large for the neighbour table"; nbr[(size_t)i * NB_W + n++] = (uint16_t)j; }
        }
        if (n == 0) return "a vertex has no neighbour within the angular neighbourhood (the reference's maximum over an empty set throws)";
    }
    return "";
}

}  // namespace fibers

using namespace fibers;

// Host-only set-up helper (no device needed; exercised by the CPU test-suite): Kernel [ndir x ncomp] column-major, the volume ->
// row map and the angular-neighbour table [nvert][16] (0xFFFF terminated).  Returns ndir, or < 0 on error.
extern "C" int fibers_host_build_rumba(int nvol, const float* bval, const float* bvec, const float* vertices, int nvert2, float ang_neig,
                                       float lambda_para, float lambda_perp, float lambda_csf, float lambda_gm,
                                       float* kernel, int64_t capacity, int32_t* vol_row, uint16_t* nbr) {
    std::vector<float> K; std::vector<int> vr; std::vector<uint16_t> nb; int ndir = 0, nb0 = 0;
    std::string e = build_rumba_kernel(nvol, bval, bvec, vertices, nvert2, lambda_para, lambda_perp, lambda_csf, lambda_gm, K, vr, ndir, nb0);
    if (e.empty()) e = build_rumba_neighbours(vertices, nvert2, ang_neig, nb);
    if (!e.empty()) { fail(FIBERS_ERR_ARG, e); return -FIBERS_ERR_ARG; }
    if (kernel) { if (capacity < (int64_t)K.size()) { fail(FIBERS_ERR_ARG, "output buffer too small"); return -FIBERS_ERR_ARG; } memcpy(kernel, K.data(), sizeof(float) * K.size()); }
    if (vol_row) for (int k = 0; k < nvol; ++k) vol_row[k] = vr[k];
    if (nbr) memcpy(nbr, nb.data(), sizeof(uint16_t) * nb.size());
    return ndir;
}

#define R_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return fail(_e == cudaErrorMemoryAllocation ? FIBERS_ERR_NOMEM : FIBERS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)

extern "C" int fibers_rumba_rec(const float* dwi, const uint8_t* mask_pos, const uint8_t* mask_any, int nx, int ny, int nz, int nvol,
                                const float* bval, const float* bvec, const float* vertices, int nvert2, float ang_neig, int niter,
                                float lambda_para, float lambda_perp, float lambda_csf, float lambda_gm, int ncoils, int coil_combine,
                                int ipat_factor, int use_tv, float* fodf, float* fgm, float* fcsf, float* peak1, float* peak2, float* peak3,
                                float* peak4, float* peak5, float* gfa, float* var, float* snr_mean, float* snr_std, int16_t* peak_idx,
                                int device) {
    if (!bval || nvol <= 0) return fail(FIBERS_ERR_TABLE, "Missing b-value table from input DWI structure");
    if (!bvec) return fail(FIBERS_ERR_TABLE, "Missing gradient table from input DWI structure");
    if (coil_combine != 0 && coil_combine != 1) return fail(FIBERS_ERR_ARG, "Unknown coil combine mode");
    if (ipat_factor < 1) return fail(FIBERS_ERR_ARG, "iPAT factor must be a positive integer");
    if (nx < 2 || ny < 2 || nz < 2) return fail(FIBERS_ERR_ARG, "volume dimensions must be at least 2 (the reference's divergence operator indexes end-1)");
    if (!dwi || !mask_pos || !vertices || !fodf || !fgm || !fcsf || !peak1 || !peak2 || !peak3 || !peak4 || !peak5 || !gfa || !var)
        return fail(FIBERS_ERR_ARG, "NULL pointer");
    if (niter < 0) return fail(FIBERS_ERR_ARG, "niter must be >= 0");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return fail(FIBERS_ERR_NODEV, "no CUDA device available (libfibers_cuda has no CPU fallback)"); }
    if (device < 0 || device >= ndev) return fail(FIBERS_ERR_ARG, "device ordinal out of range");
    if (!mask_any) mask_any = mask_pos;
    const int64_t nvox = (int64_t)nx * ny * nz;
    const int nvert = nvert2 / 2, ncomp = nvert + 2;
    if (nvert > 384) return fail(FIBERS_ERR_ARG, "more than 384 half-sphere vertices");
    std::vector<float> K; std::vector<int> vol_row; int ndir = 0, nb0 = 0;
    std::string e = build_rumba_kernel(nvol, bval, bvec, vertices, nvert2, lambda_para, lambda_perp, lambda_csf, lambda_gm, K, vol_row, ndir, nb0);
    if (!e.empty()) return fail(e.find("Missing") == 0 ? FIBERS_ERR_TABLE : FIBERS_ERR_ARG, e);
    std::vector<uint16_t> nbr;
    e = build_rumba_neighbours(vertices, nvert2, ang_neig, nbr);
    if (!e.empty()) return fail(FIBERS_ERR_ARG, e);
    std::vector<int> ind; std::vector<int> colmap((size_t)nvox, -1);
    for (int64_t v = 0; v < nvox; ++v) if (mask_pos[v]) { colmap[(size_t)v] = (int)ind.size(); ind.push_back((int)v); }
    const int nmask = (int)ind.size();
    const float norder = coil_combine == 1 ? (float)ncoils : 1.f, n2 = 2.f * norder;
    // start value and K f0 (:527-529, :246)
    std::vector<float> f0(ncomp, 1.f);
    for (auto& x : f0) x = x / (float)(2 * nvert + 2);
    float fsum = 0.f; for (float x : f0) fsum += x;
    for (auto& x : f0) x = x / fsum;
    std::vector<float> kf0(ndir, 0.f);
    for (int r = 0; r < ndir; ++r) { float t = 0.f; for (int c = 0; c < ncomp; ++c) t += K[(size_t)c * ndir + r] * f0[c]; kf0[r] = t; }
    const float s20 = (1.f / 15.f) * (1.f / 15.f);

    R_CUDA(cudaSetDevice(device));
    if (nmask > 0 && niter > 0 && !load_cublas()) return fail(FIBERS_ERR_CUDA, "cuBLAS (libcublas.so.12) could not be loaded for the RUMBA-SD matrix products");
    DevBuf B;
    cudaStream_t st = nullptr;
    R_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } sg{st};
    float *d_dwi, *d_sig, *d_ir, *d_t1, *d_dodf, *d_f[2], *d_rl, *d_rl2, *d_s2, *d_K, *d_kf0, *d_lam, *d_bsum, *d_vert;
    int *d_ind, *d_colmap, *d_volrow; uint8_t* d_maskany; uint16_t* d_nbr;
    const size_t nd = (size_t)ndir * std::max(nmask, 1), nc = (size_t)ncomp * std::max(nmask, 1);
    R_CUDA(B.alloc(&d_colmap, (size_t)nvox)); R_CUDA(B.alloc(&d_maskany, (size_t)nvox));
    R_CUDA(B.alloc(&d_ind, (size_t)std::max(nmask, 1))); R_CUDA(B.alloc(&d_volrow, (size_t)nvol));
    R_CUDA(B.alloc(&d_sig, nd)); R_CUDA(B.alloc(&d_ir, nd)); R_CUDA(B.alloc(&d_t1, nd)); R_CUDA(B.alloc(&d_dodf, nd));
    R_CUDA(B.alloc(&d_f[0], nc)); R_CUDA(B.alloc(&d_f[1], nc)); R_CUDA(B.alloc(&d_rl, nc)); R_CUDA(B.alloc(&d_rl2, nc));
    R_CUDA(B.alloc(&d_s2, (size_t)std::max(nmask, 1))); R_CUDA(B.alloc(&d_K, K.size())); R_CUDA(B.alloc(&d_kf0, (size_t)ndir));
    R_CUDA(B.alloc(&d_lam, (size_t)1)); R_CUDA(B.alloc(&d_nbr, nbr.size())); R_CUDA(B.alloc(&d_vert, (size_t)nvert * 3));
    const int wpb = 8, nblk_vox = (nmask + wpb - 1) / wpb;
    R_CUDA(B.alloc(&d_bsum, (size_t)std::max(nblk_vox, 1)));
    R_CUDA(cudaMemcpyAsync(d_colmap, colmap.data(), sizeof(int) * nvox, cudaMemcpyHostToDevice, st));
    R_CUDA(cudaMemcpyAsync(d_maskany, mask_any, (size_t)nvox, cudaMemcpyHostToDevice, st));
    if (nmask) R_CUDA(cudaMemcpyAsync(d_ind, ind.data(), sizeof(int) * nmask, cudaMemcpyHostToDevice, st));
    R_CUDA(cudaMemcpyAsync(d_volrow, vol_row.data(), sizeof(int) * nvol, cudaMemcpyHostToDevice, st));
    R_CUDA(cudaMemcpyAsync(d_K, K.data(), sizeof(float) * K.size(), cudaMemcpyHostToDevice, st));
    R_CUDA(cudaMemcpyAsync(d_kf0, kf0.data(), sizeof(float) * ndir, cudaMemcpyHostToDevice, st));
    R_CUDA(cudaMemcpyAsync(d_nbr, nbr.data(), sizeof(uint16_t) * nbr.size(), cudaMemcpyHostToDevice, st));
    std::vector<float> hv((size_t)nvert * 3);
    for (int i = 0; i < nvert; ++i) for (int d = 0; d < 3; ++d) hv[(size_t)i * 3 + d] = vertices[(size_t)d * nvert2 + i];
    R_CUDA(cudaMemcpyAsync(d_vert, hv.data(), sizeof(float) * hv.size(), cudaMemcpyHostToDevice, st));
    R_CUDA(cudaMemcpyAsync(d_lam, &s20, sizeof(float), cudaMemcpyHostToDevice, st));
    int cur = 0;
    if (nmask > 0) {
        // the DWI volume is only needed to build signal_mat: it is uploaded, reduced to the mask and freed
        R_CUDA(cudaMalloc(&d_dwi, sizeof(float) * (size_t)nvol * nvox));
        struct Free { float* p; ~Free() { cudaFree(p); } } fr{d_dwi};
        R_CUDA(cudaMemcpyAsync(d_dwi, dwi, sizeof(float) * (size_t)nvol * nvox, cudaMemcpyHostToDevice, st));
        rumba_signal_kernel<<<nblk_vox, wpb * 32, 0, st>>>(d_dwi, nvox, nvol, d_volrow, nb0, d_ind, nmask, ndir, d_sig);
        const int64_t tot = (int64_t)std::max(ndir, ncomp) * nmask;
        rumba_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d_sig, d_kf0, f0[0], s20, n2, ndir, ncomp, nmask, d_f[0], d_dodf, d_ir, d_t1, d_s2);
        count_launch(2);
        R_CUDA(cudaGetLastError());
        R_CUDA(cudaStreamSynchronize(st));
    }
    if (nmask > 0 && niter > 0) {
        void* h = nullptr;
        if (g_blas.create(&h) != 0) return fail(FIBERS_ERR_CUDA, "cublasCreate failed");
        struct HG { void* h; ~HG() { g_blas.destroy(h); } } hg{h};
        g_blas.set_stream(h, st);
        const float one = 1.f, zero = 0.f;
        const TvGeom geo{nx, ny, nz};
        for (int it = 0; it < niter; ++it) {
            // rl = K' (signal .* Iratio), rl2 = K' dodf            (CUBLAS_OP_T = 1, CUBLAS_OP_N = 0)
            if (g_blas.sgemm(h, 1, 0, ncomp, nmask, ndir, &one, d_K, ndir, d_t1, ndir, &zero, d_rl, ncomp) != 0 ||
                g_blas.sgemm(h, 1, 0, ncomp, nmask, ndir, &one, d_K, ndir, d_dodf, ndir, &zero, d_rl2, ncomp) != 0)
                return fail(FIBERS_ERR_CUDA, "cublasSgemm failed");
            rumba_update_kernel<<<nmask, 384, 0, st>>>(d_f[cur], d_f[cur ^ 1], d_rl, d_rl2, d_ind, d_colmap, geo, ncomp, use_tv ? 1 : 0, d_lam,
                                                      (use_tv && ipat_factor > 1) ? d_s2 : nullptr);
            cur ^= 1;
            // dodf = K fodf
            if (g_blas.sgemm(h, 0, 0, ndir, nmask, ncomp, &one, d_K, ndir, d_f[cur], ncomp, &zero, d_dodf, ndir) != 0)
                return fail(FIBERS_ERR_CUDA, "cublasSgemm failed");
            rumba_noise_kernel<<<nblk_vox, wpb * 32, 0, st>>>(d_sig, d_dodf, d_ir, d_t1, d_s2, ndir, nmask, n2, norder, d_bsum);
            if (use_tv && ipat_factor == 1) rumba_lambda_kernel<<<1, 256, 0, st>>>(d_bsum, nblk_vox, nmask, d_lam);
            count_launch(use_tv && ipat_factor == 1 ? 3 : 2);
        }
        R_CUDA(cudaGetLastError());
    }
    // ---- outputs ----
    float *o_fodf, *o_fgm, *o_fcsf, *o_peak, *o_gfa, *o_var; int16_t* o_idx = nullptr;
    R_CUDA(B.alloc(&o_fodf, (size_t)nvert * nvox)); R_CUDA(B.alloc(&o_fgm, (size_t)nvox)); R_CUDA(B.alloc(&o_fcsf, (size_t)nvox));
    R_CUDA(B.alloc(&o_peak, (size_t)NPEAK * 3 * nvox)); R_CUDA(B.alloc(&o_gfa, (size_t)nvox)); R_CUDA(B.alloc(&o_var, (size_t)nvox));
    if (peak_idx) R_CUDA(B.alloc(&o_idx, (size_t)NPEAK * nvox));
    RumbaOut o{};
    o.fodf = o_fodf; o.fgm = o_fgm; o.fcsf = o_fcsf; o.gfa = o_gfa; o.var = o_var; o.peak_idx = o_idx;
    for (int k = 0; k < NPEAK; ++k) o.peak[k] = o_peak + (size_t)k * 3 * nvox;
    const int fw = 8;
    rumba_final_kernel<<<(unsigned)((nvox + fw - 1) / fw), fw * 32, sizeof(float) * fw * (nvert + 2), st>>>(d_f[cur], d_s2, d_colmap, d_maskany, nvox, ncomp, d_nbr,
                                                                                                          d_vert, o);
    count_launch(1);
    R_CUDA(cudaGetLastError());
    float* peaks[NPEAK] = {peak1, peak2, peak3, peak4, peak5};
    R_CUDA(cudaMemcpyAsync(fodf, o_fodf, sizeof(float) * (size_t)nvert * nvox, cudaMemcpyDeviceToHost, st));
    R_CUDA(cudaMemcpyAsync(fgm, o_fgm, sizeof(float) * nvox, cudaMemcpyDeviceToHost, st));
    R_CUDA(cudaMemcpyAsync(fcsf, o_fcsf, sizeof(float) * nvox, cudaMemcpyDeviceToHost, st));
    R_CUDA(cudaMemcpyAsync(gfa, o_gfa, sizeof(float) * nvox, cudaMemcpyDeviceToHost, st));
    R_CUDA(cudaMemcpyAsync(var, o_var, sizeof(float) * nvox, cudaMemcpyDeviceToHost, st));
    for (int k = 0; k < NPEAK; ++k) R_CUDA(cudaMemcpyAsync(peaks[k], o.peak[k], sizeof(float) * 3 * nvox, cudaMemcpyDeviceToHost, st));
    if (peak_idx) R_CUDA(cudaMemcpyAsync(peak_idx, o_idx, sizeof(int16_t) * NPEAK * nvox, cudaMemcpyDeviceToHost, st));
    std::vector<float> s2h((size_t)std::max(nmask, 1));
    if (nmask) R_CUDA(cudaMemcpyAsync(s2h.data(), d_s2, sizeof(float) * nmask, cudaMemcpyDeviceToHost, st));
    R_CUDA(cudaStreamSynchronize(st));
    // SNR statistics of the last iteration (:549-550): mean and corrected standard deviation of 1 / sqrt(sigma^2)
    float m = 0.f, sd = 0.f;
    if (niter > 0 && nmask > 0) {
        double acc = 0; for (int i = 0; i < nmask; ++i) acc += 1.0 / sqrt((double)s2h[i]);
        m = (float)(acc / nmask);
        double dv = 0; for (int i = 0; i < nmask; ++i) { const double d = 1.0 / sqrt((double)s2h[i]) - m; dv += d * d; }
        sd = nmask > 1 ? (float)sqrt(dv / (nmask - 1)) : NAN;
    }
    if (snr_mean) *snr_mean = m;
    if (snr_std) *snr_std = sd;
    return 0;
}
