// SIMT (fp32 CUDA-core) reconstruction kernel: out[r, voxel] = sum_k Mt[k, r] * max(s[k, voxel], 0)
// fused with the GQI/DSI voxel epilogue (reference: src/gqi.jl:139-159 + find_peaks! :180-201,
// src/dsi.jl:205-258 in matrix form).  It is the correctness anchor and the fallback for shapes or
// value ranges the tensor-core kernel (recon_tc.cu) does not take; both are CUDA paths.
//
// Tile: 64 voxels x 16*J matrix rows per CTA, K streamed in chunks of 16 through a double-buffered
// shared-memory pipeline (cp.async for the matrix, register prefetch + clamp for the signal).
// 256 threads; thread (tx, ty) owns voxels 4tx..4tx+3 and rows ty + 16 j, j < J  (4J accumulators).
// ODF mode: the 64 x M ODF tile is staged in shared memory (aliasing the pipeline buffers), stored
// coalesced, and searched for local maxima of the folded mesh: candidates {v : o[v] > 0 and
// o[v] > o[n] for every neighbour n} ranked by (value desc, index asc), top 3 kept.
#include <math_constants.h>
#include <algorithm>
#include "common.cuh"

namespace fibers {
namespace {

constexpr int VT = 64;        // voxels per CTA
constexpr int KC = 16;        // K chunk
constexpr int NT = 256;       // threads
constexpr int NPART = NT / VT;

struct SimtParams {
    const float* dwi; int64_t dwi_pitch; const uint8_t* mask; int64_t nvox;
    int K;
    const float* mt; int rows_pad;      // [K][rows_pad]
    int row_base;                       // first matrix row handled by panel 0 of this launch
    int nrows;                          // rows produced by this launch
    int cvol; float dscale;             // per-voxel divisor den = dscale * s+[cvol]  (cvol < 0: none)
    int64_t out_pitch; float* out;      // odf (ODF mode) or pdf (plain mode)
    float* peak[3]; float* qa[3]; int16_t* peak_idx; int32_t* stats;
    const uint16_t* nbr; const float* vert; int M;
    const int* tile_list; const int* tile_count; int tile_cap;   // list mode (fix-up of selected 64-voxel tiles)
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ void top3_insert(float val, int idx, float tv[3], int ti[3]) {
    if (val > tv[2]) {
        if (val > tv[1]) {
            tv[2] = tv[1]; ti[2] = ti[1];
            if (val > tv[0]) { tv[1] = tv[0]; ti[1] = ti[0]; tv[0] = val; ti[0] = idx; }
            else { tv[1] = val; ti[1] = idx; }
        } else { tv[2] = val; ti[2] = idx; }
    }
}

template <int J, bool ODF>
__device__ __forceinline__ void simt_tile(const SimtParams& p, const int64_t tile_x, const int panel_y) {
    constexpr int RP = 16 * J;                       // rows per panel
    extern __shared__ __align__(16) float sm[];
    // pipeline view
    float* sS = sm;                                  // [2][KC][VT]
    float* sA = sm + 2 * KC * VT;                    // [2][KC][RP]
    // epilogue view (ODF mode): stage [RP][VT] aliases the pipeline buffers
    float* stage = sm;
    constexpr int PIPE_FLOATS = 2 * KC * VT + 2 * KC * RP;
    constexpr int MAIN_FLOATS = ODF ? (PIPE_FLOATS > RP * VT ? PIPE_FLOATS : RP * VT) : PIPE_FLOATS;
    float* s_topv = sm + MAIN_FLOATS;                // [NPART][VT][3]
    int*   s_topi = (int*)(s_topv + NPART * VT * 3); // [NPART][VT][3]
    float* s_min  = (float*)(s_topi + NPART * VT * 3);   // [NPART][VT]
    float* s_sum  = s_min + NPART * VT;              // [NPART][VT]
    int*   s_pos  = (int*)(s_sum + NPART * VT);      // [VT]
    uint16_t* s_nbr = (uint16_t*)(s_pos + VT);       // [M][NBR_W]  (ODF mode)

    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const int lv = t & (VT - 1), part = t >> 6;      // loader / epilogue mapping
    const int64_t v0 = tile_x * VT;
    const int row0 = p.row_base + panel_y * RP;      // first matrix row of this panel
    const int64_t myvox = v0 + lv;
    const bool inside = myvox < p.nvox && p.mask[myvox] != 0;

    if (t < VT) s_pos[t] = 0;
    if (ODF) for (int i = t; i < p.M * NBR_W; i += NT) s_nbr[i] = p.nbr[i];

    float acc[J][4];
#pragma unroll
    for (int j = 0; j < J; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }

    const int nchunk = (p.K + KC - 1) / KC;
    bool anypos = false;
    float sreg[KC / NPART];

    auto load_signal = [&](int kc) {
#pragma unroll
        for (int i = 0; i < KC / NPART; ++i) {
            int k = kc * KC + part + NPART * i;
            float v = 0.f;
            if (inside && k < p.K) v = fmaxf(__ldg(p.dwi + (int64_t)k * p.dwi_pitch + myvox), 0.f);   // s[s<0] = 0
            anypos |= v > 0.f;
            sreg[i] = v;
        }
    };
    auto store_signal = [&](int buf) {
#pragma unroll
        for (int i = 0; i < KC / NPART; ++i) sS[(buf * KC + part + NPART * i) * VT + lv] = sreg[i];
    };
    auto issue_matrix = [&](int kc, int buf) {
        for (int i = t; i < KC * (RP / 4); i += NT) {
            int kk = i / (RP / 4), c4 = i - kk * (RP / 4);
            int k = kc * KC + kk;
            int r = row0 + c4 * 4;
            bool ok = k < p.K && r < p.rows_pad;
            const float* src = ok ? p.mt + (int64_t)k * p.rows_pad + r : p.mt;
            cp_async16(&sA[(buf * KC + kk) * RP + c4 * 4], src, ok);
        }
    };

    issue_matrix(0, 0);
    load_signal(0);
    store_signal(0);
    cp_async_wait_all();
    __syncthreads();

    for (int kc = 0; kc < nchunk; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < nchunk) { issue_matrix(kc + 1, buf ^ 1); load_signal(kc + 1); }
        const float* cs = sS + buf * KC * VT + tx * 4;
        const float* ca = sA + buf * KC * RP + ty;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            float4 s4 = *reinterpret_cast<const float4*>(cs + kk * VT);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                float a = ca[kk * RP + 16 * j];
                acc[j][0] = fmaf(a, s4.x, acc[j][0]);
                acc[j][1] = fmaf(a, s4.y, acc[j][1]);
                acc[j][2] = fmaf(a, s4.z, acc[j][2]);
                acc[j][3] = fmaf(a, s4.w, acc[j][3]);
            }
        }
        if (kc + 1 < nchunk) store_signal(buf ^ 1);
        cp_async_wait_all();
        __syncthreads();
    }
    if (anypos) s_pos[lv] = 1;      // benign race: all writers store 1
    __syncthreads();

    // per-voxel scale: 1/den for DSI, 1 for GQI; 0 for voxels that are not computed
    float scale[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        int64_t vox = v0 + tx * 4 + c;
        float sc = s_pos[tx * 4 + c] ? 1.f : 0.f;       // mask == 0 or max(s+) == 0 -> untouched (zeros)
        if (sc != 0.f && p.cvol >= 0) {
            // q = 0 sample <= 0: zeros (the reference either skips the voxel or divides by zero: undefined)
            const float den = p.dscale * fmaxf(vox < p.nvox ? __ldg(p.dwi + (int64_t)p.cvol * p.dwi_pitch + vox) : 0.f, 0.f);
            sc = den > 0.f ? 1.f / den : 0.f;
        }
        scale[c] = sc;
    }

    if (!ODF) {
        // plain rows (DSI pdf): direct store, rows ty + 16 j of this panel
#pragma unroll
        for (int j = 0; j < J; ++j) {
            int r = row0 - p.row_base + ty + 16 * j;      // output frame
            if (r < p.nrows) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    int64_t vox = v0 + tx * 4 + c;
                    if (vox < p.nvox) p.out[(int64_t)r * p.out_pitch + vox] = scale[c] != 0.f ? acc[j][c] * scale[c] : 0.f;
                }
            }
        }
        return;
    }

    // ---- ODF mode: stage the tile, store it, find peaks ---------------------------------
#pragma unroll
    for (int j = 0; j < J; ++j) {
        float4 o;
        o.x = scale[0] != 0.f ? acc[j][0] * scale[0] : 0.f;
        o.y = scale[1] != 0.f ? acc[j][1] * scale[1] : 0.f;
        o.z = scale[2] != 0.f ? acc[j][2] * scale[2] : 0.f;
        o.w = scale[3] != 0.f ? acc[j][3] * scale[3] : 0.f;
        *reinterpret_cast<float4*>(stage + (ty + 16 * j) * VT + tx * 4) = o;
    }
    __syncthreads();
    const int M = p.M;
    if (myvox < p.nvox)
        for (int r = part; r < M; r += NPART) p.out[(int64_t)r * p.out_pitch + myvox] = stage[r * VT + lv];

    {
        const int per = (M + NPART - 1) / NPART;
        const int va = part * per, vb = min(M, va + per);
        float tv[3] = {0.f, 0.f, 0.f}; int ti[3] = {-1, -1, -1};
        float mn = CUDART_INF_F, sum = 0.f;
        for (int v = va; v < vb; ++v) {
            float val = stage[v * VT + lv];
            mn = fminf(mn, val);
            sum += val;
            bool cand = val > 0.f;
#pragma unroll
            for (int k = 0; k < NBR_W; ++k) {                       // branch-free: loads are unconditional
                const unsigned n = s_nbr[v * NBR_W + k];
                const float x = stage[(n != NBR_NONE ? n : (unsigned)v) * VT + lv];
                cand = cand & ((n == NBR_NONE) | (val > x));
            }
            if (cand) top3_insert(val, v, tv, ti);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) { s_topv[(part * VT + lv) * 3 + k] = tv[k]; s_topi[(part * VT + lv) * 3 + k] = ti[k]; }
        s_min[part * VT + lv] = mn; s_sum[part * VT + lv] = sum;
    }
    __syncthreads();
    if (part == 0) {
        float tv[3] = {0.f, 0.f, 0.f}; int ti[3] = {-1, -1, -1};
        float mn = CUDART_INF_F, sum = 0.f;
        for (int q = 0; q < NPART; ++q) {
            mn = fminf(mn, s_min[q * VT + lv]); sum += s_sum[q * VT + lv];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                int idx = s_topi[(q * VT + lv) * 3 + k];
                if (idx >= 0) top3_insert(s_topv[(q * VT + lv) * 3 + k], idx, tv, ti);
            }
        }
        float mean = sum / (float)M;
        if (myvox < p.nvox) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const bool ok = ti[k] >= 0;
                const int id = ok ? ti[k] : 0;
                p.peak[k][myvox]                  = ok ? __ldg(p.vert + id * 3 + 0) : 0.f;
                p.peak[k][myvox + p.out_pitch]     = ok ? __ldg(p.vert + id * 3 + 1) : 0.f;
                p.peak[k][myvox + 2 * p.out_pitch] = ok ? __ldg(p.vert + id * 3 + 2) : 0.f;
                p.qa[k][myvox] = ok ? tv[k] - mn : 0.f;                 // unnormalised; /= odfmax later
                if (p.peak_idx) p.peak_idx[myvox + k * p.out_pitch] = (int16_t)ti[k];
            }
        } else mean = -CUDART_INF_F;
        // max over the tile of the per-voxel mean (voxels left at zero contribute 0, as in
        // maximum(mean(odf.vol, dims=4)) over the zero-filled array, src/gqi.jl:164)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mean = fmaxf(mean, __shfl_xor_sync(0xffffffffu, mean, o));
        if ((t & 31) == 0 && mean > -CUDART_INF_F) atomicMax(p.stats, f2ord(mean));
    }
}

template <int J, bool ODF>
__global__ void __launch_bounds__(NT, (J <= 23 ? 2 : 1)) recon_simt_kernel(const SimtParams p) {
    if (p.tile_list == nullptr) { simt_tile<J, ODF>(p, blockIdx.x, blockIdx.y); return; }
    const int n = min(*p.tile_count, p.tile_cap);
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        simt_tile<J, ODF>(p, p.tile_list[i], blockIdx.y);
        __syncthreads();
    }
}

template <int J, bool ODF>
size_t simt_smem(int M) {
    constexpr int RP = 16 * J;
    size_t pipe = 2 * KC * VT + 2 * KC * RP;
    size_t mainf = ODF ? std::max<size_t>(pipe, (size_t)RP * VT) : pipe;
    size_t fl = mainf + NPART * VT * 3 * 2 + NPART * VT * 2 + VT;
    return fl * 4 + (ODF ? (size_t)M * NBR_W * 2 : 0) + 16;
}

template <int J, bool ODF>
int launch_one(const SimtParams& sp, int npanels, cudaStream_t st) {
    size_t smem = simt_smem<J, ODF>(sp.M);
    FB_CUDA(cudaFuncSetAttribute(recon_simt_kernel<J, ODF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((sp.nvox + VT - 1) / VT), (unsigned)npanels);
    if (sp.tile_list) grid = dim3((unsigned)std::min<int64_t>((sp.nvox + VT - 1) / VT, 148 * 2), (unsigned)npanels);
    recon_simt_kernel<J, ODF><<<grid, NT, smem, st>>>(sp);
    count_launch(1);
    FB_CUDA(cudaGetLastError());
    return 0;
}

__global__ void stats_init_kernel(int32_t* stats) { stats[0] = ORD_NEG_INF; stats[1] = 0; }

__global__ void qa_scale_kernel(float* q1, float* q2, float* q3, int64_t nvox, const int32_t* stats, float odfmax) {
    const float m = stats ? ord2f(stats[0]) : odfmax;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += (int64_t)gridDim.x * blockDim.x) {
        q1[i] = q1[i] / m; q2[i] = q2[i] / m; q3[i] = q3[i] / m;          // qa[ipeak].vol /= odfmax
    }
}

template <typename T>
__global__ void convert_kernel(const T* __restrict__ src, float* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = (float)src[i];
}

}  // namespace

static int launch_recon_simt_impl(Plan* p, const ReconArgs& a, const int* d_list, const int* d_count, cudaStream_t st) {
    if (a.nvox <= 0) return 0;
    SimtParams sp{};
    sp.tile_list = d_list; sp.tile_count = d_count; sp.tile_cap = (int)std::min<int64_t>(2 * ((a.nvox + VT - 1) / VT), 0x7FFFFFFF);
    sp.dwi = a.dwi; sp.dwi_pitch = a.dwi_pitch; sp.mask = a.mask; sp.nvox = a.nvox;
    sp.K = p->nvol; sp.mt = p->d_mt; sp.rows_pad = p->rows_pad;
    sp.row_base = 0; sp.nrows = p->nvert;
    sp.cvol = p->kind == PLAN_DSI ? p->cvol : -1; sp.dscale = p->dscale;
    if (p->kind == PLAN_DSI && p->cvol < 0)
        return fail(FIBERS_ERR_ARG, "DSI: no volume maps to the q-space origin (sum(p) would be 0)");
    sp.out_pitch = a.out_pitch; sp.out = a.odf;
    for (int k = 0; k < 3; ++k) { sp.peak[k] = a.peak[k]; sp.qa[k] = a.qa[k]; }
    sp.peak_idx = a.peak_idx; sp.stats = a.stats;
    sp.nbr = p->d_nbr; sp.vert = p->d_vert; sp.M = p->nvert;
    int rc;
    const int M = p->nvert;
    if (M <= 16 * 12) rc = launch_one<12, true>(sp, 1, st);
    else if (M <= 16 * 21) rc = launch_one<21, true>(sp, 1, st);
    else if (M <= 16 * 23) rc = launch_one<23, true>(sp, 1, st);
    else if (M <= 16 * 32) rc = launch_one<32, true>(sp, 1, st);
    else return fail(FIBERS_ERR_ARG, "ODF tessellations with more than 512 half-sphere vertices are not supported");
    if (rc) return rc;
    if (p->kind == PLAN_DSI && a.pdf) {
        // pdf rows: matrix rows M .. M+nvol-1.  Row panels must start on a 16-row boundary of the
        // padded matrix; the plan stores Mp starting at row_pdf = round_up(M, 16).
        SimtParams pp = sp;
        pp.row_base = (M + 15) / 16 * 16; pp.nrows = p->nvol; pp.out = a.pdf;
        int npan = (p->nvol + 16 * 21 - 1) / (16 * 21);
        rc = launch_one<21, false>(pp, npan, st);
    }
    return rc;
}

int launch_recon_simt(Plan* p, const ReconArgs& a, cudaStream_t st) { return launch_recon_simt_impl(p, a, nullptr, nullptr, st); }

// Recompute the 64-voxel tiles listed in d_list[0 .. *d_count) (device-side list; GQI ODF mode only).
int launch_recon_simt_list(Plan* p, const ReconArgs& a, const int* d_list, const int* d_count, cudaStream_t st) {
    return launch_recon_simt_impl(p, a, d_list, d_count, st);
}

int launch_stats_init(int32_t* d_stats, cudaStream_t st) {
    stats_init_kernel<<<1, 1, 0, st>>>(d_stats);
    count_launch(1);
    FB_CUDA(cudaGetLastError());
    return 0;
}

int launch_qa_scale(float* q1, float* q2, float* q3, int64_t nvox, const int32_t* d_stats, float odfmax,
                    cudaStream_t st) {
    if (nvox <= 0) return 0;
    unsigned blocks = (unsigned)std::min<int64_t>((nvox + 255) / 256, 148 * 8);
    qa_scale_kernel<<<blocks, 256, 0, st>>>(q1, q2, q3, nvox, d_stats, odfmax);
    count_launch(1);
    FB_CUDA(cudaGetLastError());
    return 0;
}

int launch_convert(const void* src, int dtype, float* dst, int64_t n, cudaStream_t st) {
    if (n <= 0) return 0;
    unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
    switch (dtype) {
        case FIBERS_F64: convert_kernel<double><<<blocks, 256, 0, st>>>((const double*)src, dst, n); break;
        case FIBERS_I16: convert_kernel<int16_t><<<blocks, 256, 0, st>>>((const int16_t*)src, dst, n); break;
        case FIBERS_U16: convert_kernel<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t*)src, dst, n); break;
        case FIBERS_I32: convert_kernel<int32_t><<<blocks, 256, 0, st>>>((const int32_t*)src, dst, n); break;
        case FIBERS_U8:  convert_kernel<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t*)src, dst, n); break;
        default: return fail(FIBERS_ERR_ARG, "unsupported dwi element type");
    }
    count_launch(1);
    FB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace fibers
