// DTI / ADC log-linear fit kernels (reference: src/dti.jl:164-213 adc_fit, :243-316 dti_fit_ls,
// :325-335 dti_maps; eigen-decomposition = StaticArrays.jl closed form, call site :311).
//
// One thread per voxel.  The DWI array is frame-major with contiguous voxels, so a warp reads
// 128 contiguous bytes per volume: fully coalesced, purely HBM-bound (4N + 64 + 1 B / voxel).
// pinv(A) (7 x N) sits in shared memory and is read with warp-uniform (broadcast) addresses.
// This translation unit is compiled with -fmad=false so that the eigen-solver follows the
// reference's unfused fp32 operation order; the dot products use explicit fmaf().
#include <math_constants.h>
#include "common.cuh"
#include "eig3.cuh"

namespace fibers {

namespace {

constexpr int DTI_THREADS = 256;
// Samples per batch of the software-pipelined hot loop (two batches of registers are live: one in flight, one being consumed).
// Same-box A/B (profiles/r2_dti_ab.txt): DTI 8 -> 65.2 % of the HBM roof (16 spills at 64 registers: 49 %), ADC 16 -> 80.0 % (8: 68.6 %).
#ifndef DTI_U
#define DTI_U 8
#endif
#ifndef ADC_U
#define ADC_U 16
#endif

struct DtiOut {
    float* p[10];   // s0, l1, l2, l3, v1, v2, v3, rd, md, fa
    int64_t pitch;
    uint8_t* valid;
};

// s0 = exp(d7), eigen, maps, stores  (src/dti.jl:305-315, :325-335)
__device__ void dti_finish(const float d[7], int64_t vox, const DtiOut& o) {
    float s0 = expf(d[6]);
    float w[3], v[3][3];
    eig3_sym(d[0], d[1], d[2], d[3], d[4], d[5], w, v);
    float l1 = w[2], l2 = w[1], l3 = w[0];
    float rd = l2 + l3;
    float md = (l1 + rd) / 3.f;
    rd = rd / 2.f;
    float a = l1 - md, b = l2 - md, c = l3 - md;
    float fa = sqrtf((a * a + b * b + c * c) / (l1 * l1 + l2 * l2 + l3 * l3) * 1.5f);
    o.p[0][vox] = s0; o.p[1][vox] = l1; o.p[2][vox] = l2; o.p[3][vox] = l3;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        o.p[4][vox + r * o.pitch] = v[r][2];
        o.p[5][vox + r * o.pitch] = v[r][1];
        o.p[6][vox + r * o.pitch] = v[r][0];
    }
    o.p[7][vox] = rd; o.p[8][vox] = md; o.p[9][vox] = fa;
    if (o.valid) o.valid[vox] = 1;
}

__device__ void dti_zero(int64_t vox, const DtiOut& o) {
    o.p[0][vox] = 0.f; o.p[1][vox] = 0.f; o.p[2][vox] = 0.f; o.p[3][vox] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        o.p[4][vox + r * o.pitch] = 0.f; o.p[5][vox + r * o.pitch] = 0.f; o.p[6][vox + r * o.pitch] = 0.f;
    }
    o.p[7][vox] = 0.f; o.p[8][vox] = 0.f; o.p[9][vox] = 0.f;
    if (o.valid) o.valid[vox] = 0;
}

// NC = 7 (DTI) or 2 (ADC).  One thread per voxel streams its N samples (coalesced across the warp).
//
// Full-sample voxels: d = pinv(A) * ln s.
// Partly non-positive voxels (src/dti.jl:297-298: d = pinv(A[ipos,:]) * ln s[ipos]): with r removed
// rows U = A[removed,:]' the normal matrix is a rank-r downdate, G_pos = G - U U', so by Woodbury
//     d = x0 + P_r (I - U' P_r)^-1 U' x0,   x0 = pinv(A) * (ln s with zeros at removed rows),
// where P_r = pinv(A)[:, removed] (= G^-1 U for a full-column-rank design).  x0 falls out of the main
// loop for free; the r x r system (r <= RMAX) is solved per thread in fp32.  Voxels with more
// removed samples, or a singular downdate, go to the general per-voxel pinv kernel via a device list.
constexpr int RMAX = 8;     // (one voxel with 7 dropped samples on the per-voxel pinv path costs 0.27 ms of latency: fp64 Jacobi in ONE thread; with 8 the cfg4 phantom never leaves the inline path)
constexpr int CW = 8;     // coefficient row width in shared memory: [nvol][CW] = pinv(A)' padded (2 x LDS.128 per sample)

template <int NC>
__global__ void __launch_bounds__(DTI_THREADS, 4)      // <= 64 registers (4 CTAs = 32 warps per SM: the loop is bound by loads in flight); only the rare downdate may spill
fit_full_kernel(const float* __restrict__ dwi, int64_t pitch, const uint8_t* __restrict__ mask, int64_t nvox,
                int nvol, int nb0, const float* __restrict__ pinv, const float* __restrict__ design, const uint8_t* __restrict__ ib0,
                DtiOut out, float* __restrict__ adc, float* __restrict__ adc_s0,
                int* __restrict__ list, int* __restrict__ count) {
    extern __shared__ __align__(16) float sm[];
    // [nvol][CW]: column j of pinv(A) times ln 2 (the loop works on log2 s: MUFU.LG2 without the conversion multiply), zero padded
    float* spa = sm;
    for (int i = threadIdx.x; i < CW * nvol; i += blockDim.x) {
        const int j = i / CW, k = i - j * CW;
        spa[i] = k < NC ? pinv[k * nvol + j] * 0.693147182f : 0.f;
    }
    __syncthreads();
    const int64_t vox = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vox >= nvox) return;
    const bool inside = mask[vox] != 0;
    float d[NC];
    // accumulators as fp32 PAIRS (d0 d1)(d2 d3)(d4 d5)(d6 -): one FFMA2 (two IEEE fma per instruction, sm_100) per pair and sample
    unsigned long long dp[(NC + 1) / 2];
#pragma unroll
    for (int k = 0; k < (NC + 1) / 2; ++k) dp[k] = 0ull;
#pragma unroll
    for (int k = 0; k < NC; ++k) d[k] = 0.f;
    int nrem = 0, b0rem = 0;                           // removed (non-positive) samples, and how many of them are minimum-b volumes
    int rem[RMAX];
    // Hot loop: a non-positive sample is replaced by 1 (log = 0: it contributes nothing,
    // src/dti.jl:297-298 drops its row) with one compare + select; WHICH samples were dropped is only looked at when the
    // minimum of a group of UNROLL samples is not positive (rare), so the loop carries no counters.
    auto sample = [&](float s, int j) {
        // MUFU.LG2 (<= 2 ulp of log2 outside [0.5, 2], 2^-22 absolute inside).  A denormal sample is flushed to zero
        // (-inf) and an infinite one gives +inf: both leave a non-finite fit, which is detected after the loop and
        // sent to the exact-log path -- no per-sample range test.
        float l2;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(s > 0.f ? s : 1.f));
        if (NC == 2) {                                  // ADC: two scalar FFMA (the packed form measured 0.9 % slower here, 1.3 % faster for DTI)
            const float2 c = *reinterpret_cast<const float2*>(spa + j * CW);
            d[0] = fmaf(c.x, l2, d[0]); d[1 % NC] = fmaf(c.y, l2, d[1 % NC]);
            return;
        }
        unsigned long long l22;
        asm("mov.b64 %0, {%1, %1};" : "=l"(l22) : "f"(l2));
        const ulonglong2 c0 = *reinterpret_cast<const ulonglong2*>(spa + j * CW);
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dp[0]) : "l"(c0.x), "l"(l22));
        if (NC > 2) {
            const ulonglong2 c1 = *reinterpret_cast<const ulonglong2*>(spa + j * CW + 4);
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dp[1 % ((NC + 1) / 2)]) : "l"(c0.y), "l"(l22));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dp[2 % ((NC + 1) / 2)]) : "l"(c1.x), "l"(l22));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dp[3 % ((NC + 1) / 2)]) : "l"(c1.y), "l"(l22));     // (second half: the zero pad)
        }
    };
    auto dropped = [&](int j) { if (nrem < RMAX) rem[nrem] = j; ++nrem; b0rem += ib0[j] ? 1 : 0; };
    if (inside) {
        // the row base (volume j) is warp-uniform and the voxel offset fits 32 bits: the address is formed as
        // uniform base + 32-bit thread offset, without per-thread 64-bit arithmetic
        const uint32_t vox32 = (uint32_t)vox;
        const float* rb = dwi;
        int j = 0;
        // Software-pipelined by one batch: the loads of batch b + 1 are issued BEFORE the arithmetic of batch b, so a warp keeps
        // UNROLL loads in flight while it computes (two register sets, the loop is unrolled by two batches: no moves).
        constexpr int UNROLL = NC == 7 ? DTI_U : ADC_U;
        auto load = [&](float (&s)[UNROLL]) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) s[u] = __ldg(rb + (int64_t)u * pitch + vox32);
            rb += (int64_t)UNROLL * pitch;
        };
        auto process = [&](const float (&s)[UNROLL], int j0) {
            float mn = s[0];
#pragma unroll
            for (int u = 1; u < UNROLL; ++u) asm("min.NaN.f32 %0, %0, %1;" : "+f"(mn) : "f"(s[u]));   // (a NaN sample counts as dropped, like `s > 0` in the reference)
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) sample(s[u], j0 + u);
            if (!(mn > 0.f)) {                                              // rare: remember which samples were dropped
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) if (!(s[u] > 0.f)) dropped(j0 + u);
            }
        };
        const int nbatch = nvol / UNROLL;
        float sa[UNROLL], sb[UNROLL];
        if (nbatch > 0) load(sa);
        int b = 0;
        for (; b + 2 <= nbatch; b += 2) {
            load(sb);
            process(sa, b * UNROLL);
            if (b + 2 < nbatch) load(sa);
            process(sb, (b + 1) * UNROLL);
        }
        if (b < nbatch) { process(sa, b * UNROLL); ++b; }
        j = nbatch * UNROLL;
        for (int u = 0; j < nvol; ++j, ++u) {
            const float sv = __ldg(rb + (int64_t)u * pitch + vox32);
            sample(sv, j);
            if (!(sv > 0.f)) dropped(j);
        }
    }
    if (NC > 2) {
#pragma unroll
        for (int k = 0; k < (NC + 1) / 2; ++k) {
            float lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(dp[k]));
            d[2 * k] = lo;
            if (2 * k + 1 < NC) d[2 * k + 1] = hi;
        }
    }
    const int npos = nvol - nrem;
    const bool b0pos = nb0 - b0rem > 0;                             // some minimum-b volume is positive
    const bool full = inside && npos == nvol;                       // src/dti.jl:294
    bool part = inside && !full && npos > 6 && b0pos;               // :297 (ADC keeps the same rule, :206)
    bool solved = full;
    {
        bool finite = true;
#pragma unroll
        for (int k = 0; k < NC; ++k) finite = finite && (fabsf(d[k]) < CUDART_INF_F);
        if (!finite && (full || part)) { solved = false; part = true; nrem = RMAX + 1; }   // denormal / inf sample: exact-log path
    }
    if (part && nrem <= RMAX) {
        // rank-nrem downdate (fp32: the correction is a small, well-conditioned update; fp64 runs at 1/64 rate here)
        float S[RMAX][RMAX], z[RMAX];
        for (int i = 0; i < nrem; ++i) {
            const float* ai = design + (size_t)rem[i] * NC;
            float r = 0.f;
            for (int k = 0; k < NC; ++k) r = fmaf(ai[k], d[k], r);
            z[i] = r;                                                // U' x0
            for (int jj = 0; jj < nrem; ++jj) {
                float t = 0.f;
                for (int k = 0; k < NC; ++k) t = fmaf(ai[k], pinv[k * nvol + rem[jj]], t);
                S[i][jj] = (i == jj ? 1.f : 0.f) - t;                // I - U' P_r
            }
        }
        bool ok = true;
        for (int c = 0; c < nrem && ok; ++c) {                       // Gaussian elimination, partial pivoting
            int piv = c; float best = fabsf(S[c][c]);
            for (int i = c + 1; i < nrem; ++i) if (fabsf(S[i][c]) > best) { best = fabsf(S[i][c]); piv = i; }
            if (!(best > 1e-5f)) { ok = false; break; }               // (near-)singular downdate: general path decides
            if (piv != c) { for (int k = 0; k < nrem; ++k) { float t = S[c][k]; S[c][k] = S[piv][k]; S[piv][k] = t; } float t = z[c]; z[c] = z[piv]; z[piv] = t; }
            const float inv = 1.f / S[c][c];
            for (int i = c + 1; i < nrem; ++i) {
                const float f = S[i][c] * inv;
                for (int k = c; k < nrem; ++k) S[i][k] -= f * S[c][k];
                z[i] -= f * z[c];
            }
        }
        if (ok) {
            for (int c = nrem - 1; c >= 0; --c) {
                float t = z[c];
                for (int k = c + 1; k < nrem; ++k) t -= S[c][k] * z[k];
                z[c] = t / S[c][c];
            }
            for (int i = 0; i < nrem; ++i) {
#pragma unroll
                for (int k = 0; k < NC; ++k) d[k] = fmaf(pinv[k * nvol + rem[i]], z[i], d[k]);
            }
            solved = true; part = false;
        }
    }
    if (solved) {
        if (NC == 7) dti_finish(d, vox, out);
        else { adc[vox] = d[0]; adc_s0[vox] = expf(d[1]); }
    } else {
        if (NC == 7) dti_zero(vox, out);
        else { adc[vox] = 0.f; adc_s0[vox] = 0.f; }
        if (part) { int slot = atomicAdd(count, 1); list[slot] = (int)vox; }
    }
}

// Partial-sample path: d = pinv(A[ipos,:]) * log(s[ipos])  (src/dti.jl:298, :207).
// One thread per listed voxel; pinv through the eigen-decomposition of the NCxNC normal matrix in
// float64 (cyclic Jacobi) with LAPACK-pinv truncation (sigma <= eps32*min(m,n)*sigma_max dropped).
template <int NC>
__global__ void fit_partial_kernel(const float* __restrict__ dwi, int64_t pitch, int nvol,
                                   const float* __restrict__ design /*[nvol][NC]*/,
                                   DtiOut out, float* __restrict__ adc, float* __restrict__ adc_s0,
                                   const int* __restrict__ list, const int* __restrict__ count, int* __restrict__ next_count) {
  // the counter of the NEXT launch of this plan is reset here (the two alternate), which saves a memset per call
  if (blockIdx.x == 0 && threadIdx.x == 0) *next_count = 0;
  const int total = *count;
  // One WARP per listed voxel: the lanes split the samples (a single thread walking 288 dependent, uncoalesced loads took 0.27 ms
  // for ONE voxel: the latency of the whole call), the normal matrix is reduced with shuffles, every lane then runs the small
  // eigen-solve redundantly and lane 0 stores.
  const int lane = threadIdx.x & 31;
  const int nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < total; i += nwarp) {
    const int64_t vox = list[i];
    double G[NC][NC], V[NC][NC], rhs[NC];
    for (int a = 0; a < NC; ++a) { rhs[a] = 0; for (int b = 0; b < NC; ++b) { G[a][b] = 0; V[a][b] = (a == b); } }
    int npos = 0;
    for (int j = lane; j < nvol; j += 32) {
        float s = dwi[vox + (int64_t)j * pitch];
        if (!(s > 0.f)) continue;
        ++npos;
        double lg = (double)logf(s);
        double row[NC];
        for (int a = 0; a < NC; ++a) row[a] = (double)design[j * NC + a];
        for (int a = 0; a < NC; ++a) {
            rhs[a] += row[a] * lg;
            for (int b = a; b < NC; ++b) G[a][b] += row[a] * row[b];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        npos += __shfl_xor_sync(0xffffffffu, npos, o);
        for (int a = 0; a < NC; ++a) {
            rhs[a] += __shfl_xor_sync(0xffffffffu, rhs[a], o);
            for (int b = a; b < NC; ++b) G[a][b] += __shfl_xor_sync(0xffffffffu, G[a][b], o);
        }
    }
    for (int a = 0; a < NC; ++a) for (int b = 0; b < a; ++b) G[a][b] = G[b][a];
    // cyclic Jacobi: G = V diag(lam) V'
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0, dg = 0;
        for (int a = 0; a < NC; ++a) { dg += G[a][a] * G[a][a]; for (int b = a + 1; b < NC; ++b) off += G[a][b] * G[a][b]; }
        if (off <= 1e-36 * dg) break;
        for (int p = 0; p < NC - 1; ++p)
            for (int q = p + 1; q < NC; ++q) {
                double apq = G[p][q];
                if (apq == 0.0) continue;
                double theta = (G[q][q] - G[p][p]) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < NC; ++k) {
                    double gkp = G[k][p], gkq = G[k][q];
                    G[k][p] = c * gkp - s * gkq; G[k][q] = s * gkp + c * gkq;
                }
                for (int k = 0; k < NC; ++k) {
                    double gpk = G[p][k], gqk = G[q][k];
                    G[p][k] = c * gpk - s * gqk; G[q][k] = s * gpk + c * gqk;
                }
                for (int k = 0; k < NC; ++k) {
                    double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    double lmax = 0;
    for (int a = 0; a < NC; ++a) lmax = fmax(lmax, G[a][a]);
    const int mn = npos < NC ? npos : NC;
    const double rtol = 1.1920929e-7 * mn;
    float d[NC];
    for (int a = 0; a < NC; ++a) d[a] = 0.f;
    double acc[NC];
    for (int a = 0; a < NC; ++a) acc[a] = 0;
    for (int k = 0; k < NC; ++k) {
        double lam = G[k][k];
        if (!(lam > rtol * rtol * lmax) || lam <= 0) continue;
        double proj = 0;
        for (int a = 0; a < NC; ++a) proj += V[a][k] * rhs[a];
        proj /= lam;
        for (int a = 0; a < NC; ++a) acc[a] += V[a][k] * proj;
    }
    for (int a = 0; a < NC; ++a) d[a] = (float)acc[a];
    if (lane == 0) {
        if (NC == 7) dti_finish(d, vox, out);
        else { adc[vox] = d[0]; adc_s0[vox] = expf(d[1]); }
    }
  }
}

int ensure_list(Plan* p, int64_t nvox) {
    if (p->list_cap >= nvox && p->d_count) return 0;
    if (p->d_list) cudaFree(p->d_list);
    p->d_list = nullptr;
    FB_CUDA(cudaMalloc(&p->d_list, sizeof(int) * (size_t)nvox));
    if (!p->d_count) { FB_CUDA(cudaMalloc(&p->d_count, 2 * sizeof(int))); FB_CUDA(cudaMemset(p->d_count, 0, 2 * sizeof(int))); }
    p->list_cap = nvox;
    return 0;
}

template <int NC>
int launch_fit(Plan* p, const float* d_dwi, int64_t dwi_pitch, const uint8_t* d_mask, int64_t nvox,
               DtiOut out, float* adc, float* adc_s0, cudaStream_t st) {
    if (nvox <= 0) return 0;
    if (nvox > 0x7FFFFFFFLL) return fail(FIBERS_ERR_ARG, "slab too large (nvox must fit in int32)");
    int rc = ensure_list(p, nvox);
    if (rc) return rc;
    if (int rc2 = plan_enter(p, st)) return rc2;                    // the partial-path list / counter belong to the plan
    int* cnt = p->d_count + (p->count_flip & 1); int* cnt_next = p->d_count + ((p->count_flip + 1) & 1);
    ++p->count_flip;                                               // (launches of one plan are serialised by plan_enter / plan_leave)
    size_t smem = sizeof(float) * CW * p->nvol + p->nvol + 16;
    if (smem > 48 * 1024)
        FB_CUDA(cudaFuncSetAttribute(fit_full_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned blocks = (unsigned)((nvox + DTI_THREADS - 1) / DTI_THREADS);
    fit_full_kernel<NC><<<blocks, DTI_THREADS, smem, st>>>(d_dwi, dwi_pitch, d_mask, nvox, p->nvol, p->nb0, p->d_pinv, p->d_design,
                                                           p->d_ib0, out, adc, adc_s0, p->d_list, cnt);
    // The partial path is rare: a fixed 4-CTA-per-SM grid strides over the device-side list
    // (no host sync needed to learn the count).  64-thread blocks: heavy per-thread state.
    unsigned pblocks = (unsigned)std::min<int64_t>((nvox + 1) / 2, 148 * 4);        // two warps (= two listed voxels at a time) per block
    fit_partial_kernel<NC><<<pblocks, 64, 0, st>>>(d_dwi, dwi_pitch, p->nvol, p->d_design, out, adc, adc_s0,
                                                   p->d_list, cnt, cnt_next);
    count_launch(2);
    FB_CUDA(cudaGetLastError());
    return plan_leave(p, st);
}

}  // namespace

int launch_dti(Plan* p, const float* d_dwi, int64_t dwi_pitch, const uint8_t* d_mask, int64_t nvox,
               int64_t out_pitch, float* const outp[10], uint8_t* d_valid, cudaStream_t st) {
    DtiOut o;
    for (int i = 0; i < 10; ++i) o.p[i] = outp[i];
    o.pitch = out_pitch; o.valid = d_valid;
    return launch_fit<7>(p, d_dwi, dwi_pitch, d_mask, nvox, o, nullptr, nullptr, st);
}

int launch_adc(Plan* p, const float* d_dwi, int64_t dwi_pitch, const uint8_t* d_mask, int64_t nvox,
               float* d_adc, float* d_s0, cudaStream_t st) {
    DtiOut o{};
    return launch_fit<2>(p, d_dwi, dwi_pitch, d_mask, nvox, o, d_adc, d_s0, st);
}

}  // namespace fibers
