// Tensor-core (tcgen05 / UMMA) GQI reconstruction kernel for sm_100a, fused with the voxel epilogue.
// Replaces the reference's per-voxel `mul!(o, A, s)` + find_peaks! + QA (src/gqi.jl:139-159, :180-201).
//
//   ODF[voxel, vertex] = sum_k s+[voxel, k] * A[vertex, k]          (a genuine dense contraction)
//
// fp32-accurate split operands on the fp16 tensor pipe (kind::f16, fp32 accumulate in TMEM):
//   s+ * scale = s_hi + s_lo,  A = A_hi + A_lo  (each fp16),  ODF ~= s_hi A_hi + s_lo A_hi + s_hi A_lo
//   (scale is a power of two chosen from a strided sample of the slab; voxels whose scaled signal
//   overflows fp16 produce non-finite ODFs, are detected in the epilogue and recomputed by the SIMT
//   kernel -- see launch_recon_tc.)
//
// One persistent CTA PAIR (cta_group::2, M = 256) per two SMs; every CTA owns 128 voxels of a tile:
//   warp 0      TMA producer: streams K16 chunks of the split matrix (this CTA's half of the rows)
//               from L2 into a 4-stage SWIZZLE_32B shared-memory ring
//   warp 1      MMA issuer (leader CTA): tcgen05.mma.cta_group::2, A operand from TENSOR MEMORY,
//               B from shared memory; accumulators D[128 x Npad] fp32 in TMEM columns [0, Npad)
//   warps 2-9   converters: coalesced fp32 loads of the DWI slab (voxel-contiguous), clamp, scale,
//               hi/lo fp16 split, tcgen05.st into a 4-slot TMEM ring (columns 384..511)
//   warps 10-17 epilogue: tcgen05.ld, un-scale, coalesced ODF store, stage the 128 x M tile in shared
//               memory, local-maximum search on the folded mesh + top-3 + QA, per-voxel mean -> atomicMax
// The full ODF never round-trips HBM: it is written once and the peaks come from the staged tile.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math_constants.h>
#include <algorithm>
#include <cmath>
#include "common.cuh"

namespace fibers {

int launch_recon_simt_list(Plan* p, const ReconArgs& a, const int* d_list, const int* d_count, cudaStream_t st);

namespace {

constexpr int TC_THREADS = 576;
constexpr int W_MMA = 1, W_CONV0 = 2, W_EPI0 = 10;
constexpr int NSTAGE = 4;            // B ring: K16 chunks
constexpr int ASLOT = 4;             // A ring in TMEM: K32 chunks, 32 columns each
constexpr int TMEM_A_COL = 384;
constexpr int VOX_CTA = 128;
constexpr int EPI_THREADS = 256;
constexpr float FP16_TARGET = 8192.f;   // the sampled maximum is scaled to <= 8192 (8x headroom to 65504)

struct TcParams {
    const float* dwi; int64_t dwi_pitch; const uint8_t* mask; int64_t nvox;
    int K, Kpad;                 // Kpad: multiple of 32
    int M, Npad, N1, N2;         // D columns: [0,N1) block 1, [N1, N1+N2) block 2; N1, N2 multiples of 16
    float* odf; int64_t out_pitch; float* peak[3]; float* qa[3]; int16_t* peak_idx; int32_t* stats;
    const uint16_t* nbr; const float* vert;
    const int* maxbits;          // device: bit pattern of the sampled max(s) (>= 0)
    int* fix_list; int* fix_count; int fix_cap;
    int ntiles;                  // 256-voxel tiles
};

struct TcState {
    __half* d_split = nullptr;   // [2 ranks][hi Nh rows | lo Nh rows][Kpad]
    CUtensorMap tmap;
    int Kpad = 0, Npad = 0, N1 = 0, N2 = 0;
    size_t smem = 0;
    int* d_scratch = nullptr;    // [0] maxbits, [1] fix_count, [2..] fix list
    int64_t scratch_cap = 0;
};

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// arrive on the barrier at cluster-shared address `addr` (own CTA or the pair's leader)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();        // never hang the GPU: a lost signal becomes a launch error
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem], CTA pair
__device__ __forceinline__ void mma_ts2(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit2(uint64_t* bar) {     // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__host__ __device__ inline uint32_t make_idesc_f16(int Mdim, int Ndim) {       // f16 x f16 -> f32, K-major A and B
    return (1u << 4) | ((uint32_t)(Ndim >> 3) << 17) | ((uint32_t)(Mdim >> 4) << 24);
}
__device__ __forceinline__ uint64_t make_sdesc_sw32(uint32_t saddr) {          // K-major, SWIZZLE_32B, 8-row groups 256 B apart
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
}

__device__ __forceinline__ void top3_insert(float val, int idx, float tv[3], int ti[3]) {
    if (val > tv[2]) {
        if (val > tv[1]) {
            tv[2] = tv[1]; ti[2] = ti[1];
            if (val > tv[0]) { tv[1] = tv[0]; ti[1] = ti[0]; tv[0] = val; ti[0] = idx; }
            else { tv[1] = val; ti[1] = idx; }
        } else { tv[2] = val; ti[2] = idx; }
    }
}

// ---------------------------------------------------------------------------------------------
// strided sample of the slab: max over 32-voxel runs every 2048 voxels of every volume
// ---------------------------------------------------------------------------------------------
__global__ void sample_max_kernel(const float* __restrict__ dwi, int64_t pitch, int64_t nvox, int nvol, int* maxbits) {
    const int64_t nrun = (nvox + 2047) / 2048;
    const int nkg = (nvol + 15) / 16;                 // 16 volumes per task: 16 independent loads in flight
    const int lane = threadIdx.x & 31;
    float m = 0.f;
    for (int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < nrun * nkg; w += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int64_t run = w % nrun; const int k0 = (int)(w / nrun) * 16;
        const int64_t v = run * 2048 + lane;
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = (v < nvox && k0 + j < nvol) ? __ldg(dwi + (int64_t)(k0 + j) * pitch + v) : 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) m = fmaxf(m, x[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && m > 0.f && m < CUDART_INF_F) atomicMax(maxbits, __float_as_int(m));
}

// ---------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
recon_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap tmapB) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int cluster_id = blockIdx.x >> 1, ncluster = gridDim.x >> 1;
    const int Nh = (p.N1 + p.N2) >> 1, N1h = p.N1 >> 1;
    const uint32_t stage_bytes = (uint32_t)(2 * Nh * 32);
    const int nk16 = p.Kpad >> 4, nk32 = p.Kpad >> 5;

    // ---- shared memory carve-up -------------------------------------------------------------
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sB = base;                                                   // NSTAGE * stage_bytes
    float* stage = (float*)(sB + NSTAGE * stage_bytes);                   // [M][128]
    float* s_topv = stage + (size_t)p.M * VOX_CTA;                        // [2][128][3]
    int* s_topi = (int*)(s_topv + 2 * VOX_CTA * 3);                       // [2][128][3]
    float* s_min = (float*)(s_topi + 2 * VOX_CTA * 3);                    // [2][128]
    float* s_sum = s_min + 2 * VOX_CTA;                                   // [2][128]
    uint16_t* s_nbr = (uint16_t*)(s_sum + 2 * VOX_CTA);                   // [M][NBR_W], 16-byte aligned rows
    uint64_t* bars = (uint64_t*)(s_nbr + (size_t)p.M * NBR_W);
    uint64_t* b_full = bars, *b_empty = bars + NSTAGE, *a_full = bars + 2 * NSTAGE, *a_empty = bars + 2 * NSTAGE + ASLOT;
    uint64_t* d_full = bars + 2 * NSTAGE + 2 * ASLOT, *d_empty = d_full + 1;
    uint32_t* tmem_ptr_s = (uint32_t*)(d_empty + 1);

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < ASLOT; ++i) { mbar_init(&a_full[i], 8); mbar_init(&a_empty[i], 1); }   // 4 converter warps x 2 CTAs
        mbar_init(d_full, 1); mbar_init(d_empty, 16);                                             // 8 epilogue warps x 2 CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.M * NBR_W; i += TC_THREADS) s_nbr[i] = p.nbr[i];
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    float scale = 1.f, inv_scale = 1.f;
    {
        const float mx = __int_as_float(*p.maxbits);
        if (mx > 0.f) {
            int e = (int)floorf(log2f(FP16_TARGET / mx));
            e = max(-100, min(100, e));
            scale = exp2f((float)e); inv_scale = exp2f((float)-e);
        }
    }

    if (warp == 0) {
        // ===== TMA producer: this CTA's half of the split matrix rows, K16 per stage ===========
        if (lane == 0) {
            const uint32_t full0 = mapa(smem_u32(&b_full[0]), 0);
            uint32_t g = 0;
            for (int tile = cluster_id; tile < p.ntiles; tile += ncluster) {
                for (int c = 0; c < nk16; ++c, ++g) {
                    const int s = g % NSTAGE; const uint32_t use = g / NSTAGE;
                    mbar_wait(&b_empty[s], (use & 1) ^ 1);
                    if (rank == 0) mbar_expect_tx(&b_full[s], 2 * stage_bytes);       // both CTAs' bytes land on the leader's barrier
                    const uint32_t dst = smem_u32(sB + s * stage_bytes);
                    const uint32_t bar = full0 + s * 8;
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                     ::"r"(dst + h * Nh * 32), "l"(&tmapB), "r"(c * 16), "r"((int)(rank * 2 * Nh + h * Nh)), "r"(bar) : "memory");
                }
            }
        }
    } else if (warp == W_MMA) {
        // ===== MMA issuer (leader CTA, one lane) ================================================
        if (rank == 0 && lane == 0) {
            const uint32_t idesc1 = make_idesc_f16(256, p.N1);
            const uint32_t idesc2 = p.N2 ? make_idesc_f16(256, p.N2) : 0u;
            uint32_t g16 = 0, g32 = 0, it = 0;
            for (int tile = cluster_id; tile < p.ntiles; tile += ncluster, ++it) {
                mbar_wait(d_empty, (it & 1) ^ 1);                    // epilogue of the previous tile has drained TMEM
                tc_fence_after();
                for (int c = 0; c < nk16; ++c, ++g16) {
                    const int s = g16 % NSTAGE;
                    const int slot = g32 % ASLOT;
                    mbar_wait(&b_full[s], (g16 / NSTAGE) & 1);
                    if ((c & 1) == 0) mbar_wait(&a_full[slot], (g32 / ASLOT) & 1);
                    tc_fence_after();
                    const uint32_t a_hi = tmem_base + TMEM_A_COL + slot * 32 + (c & 1) * 8;
                    const uint32_t a_lo = a_hi + 16;
                    const uint32_t bs = smem_u32(sB + s * stage_bytes);
                    const uint64_t bhi1 = make_sdesc_sw32(bs), blo1 = make_sdesc_sw32(bs + Nh * 32);
                    const uint32_t acc = c > 0 ? 1u : 0u;
                    mma_ts2(tmem_base, a_lo, bhi1, idesc1, acc);       // small terms first
                    mma_ts2(tmem_base, a_hi, blo1, idesc1, 1u);
                    mma_ts2(tmem_base, a_hi, bhi1, idesc1, 1u);
                    if (p.N2) {
                        const uint64_t bhi2 = make_sdesc_sw32(bs + N1h * 32), blo2 = make_sdesc_sw32(bs + Nh * 32 + N1h * 32);
                        mma_ts2(tmem_base + p.N1, a_lo, bhi2, idesc2, acc);
                        mma_ts2(tmem_base + p.N1, a_hi, blo2, idesc2, 1u);
                        mma_ts2(tmem_base + p.N1, a_hi, bhi2, idesc2, 1u);
                    }
                    mma_commit2(&b_empty[s]);                          // stage reusable once these MMAs retire
                    if (c & 1) { mma_commit2(&a_empty[slot]); ++g32; }
                }
                mma_commit2(d_full);
            }
        }
    } else if (warp < W_EPI0) {
        // ===== converters: DWI fp32 -> clamp -> scale -> fp16 hi/lo -> TMEM ring ================
        const int cw = warp - W_CONV0, grp = cw >> 2, q = warp & 3;
        const int vl = q * 32 + lane;                                   // TMEM lane == voxel within the CTA's 128
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t afull0 = mapa(smem_u32(&a_full[0]), 0);
        uint32_t it = 0;
        for (int tile = cluster_id; tile < p.ntiles; tile += ncluster, ++it) {
            const int64_t vox = (int64_t)tile * 256 + rank * VOX_CTA + vl;
            const bool inside = vox < p.nvox && p.mask[vox] != 0;
            const float* src = p.dwi + vox;
            for (int c = grp; c < nk32; c += 2) {
                float x[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int k = c * 32 + j;
                    x[j] = (inside && k < p.K) ? __ldg(src + (int64_t)k * p.dwi_pitch) : 0.f;
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float v0 = fmaxf(x[2 * j], 0.f) * scale, v1 = fmaxf(x[2 * j + 1], 0.f) * scale;   // s[s<0] = 0
                    const __half2 h = __floats2half2_rn(v0, v1);
                    const float2 hf = __half22float2(h);
                    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                    hi[j] = *reinterpret_cast<const uint32_t*>(&h);
                    lo[j] = *reinterpret_cast<const uint32_t*>(&l);
                }
                const uint32_t g32 = it * nk32 + c;
                const int slot = g32 % ASLOT;
                mbar_wait(&a_empty[slot], ((g32 / ASLOT) & 1) ^ 1);
                tc_fence_after();
                const uint32_t col = lane_addr + TMEM_A_COL + slot * 32;
                tmem_st8(col, hi); tmem_st8(col + 8, hi + 8); tmem_st8(col + 16, lo); tmem_st8(col + 24, lo + 8);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(afull0 + slot * 8);
            }
        }
    } else {
        // ===== epilogue ==========================================================================
        const int ew = warp - W_EPI0, part = ew >> 2, q = warp & 3;
        const int vl = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t dempty0 = mapa(smem_u32(d_empty), 0);
        const int csplit = p.N2 ? p.N1 : ((p.Npad / 2 + 15) & ~15);
        const int c_begin = part ? csplit : 0, c_end = part ? p.Npad : csplit;
        const int M = p.M;
        const int vhalf = (M + 1) >> 1;
        uint32_t it = 0;
        for (int tile = cluster_id; tile < p.ntiles; tile += ncluster, ++it) {
            const int64_t vox = (int64_t)tile * 256 + rank * VOX_CTA + vl;
            const bool vok = vox < p.nvox;
            mbar_wait(d_full, it & 1);
            tc_fence_after();
            float mn = CUDART_INF_F, sum = 0.f;
            for (int c0 = c_begin; c0 < c_end; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(lane_addr + c0, r);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int col = c0 + j;
                    if (col < M) {
                        const float val = __uint_as_float(r[j]) * inv_scale;
                        stage[col * VOX_CTA + vl] = val;
                        if (vok) p.odf[(int64_t)col * p.out_pitch + vox] = val;
                        mn = fminf(mn, val); sum += val;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(dempty0);               // TMEM may be overwritten by the next tile
            s_min[part * VOX_CTA + vl] = mn; s_sum[part * VOX_CTA + vl] = sum;
            named_bar(1, EPI_THREADS);
            // ---- local maxima of the folded mesh: strictly greater than every neighbour, > 0 ----
            {
                const int va = part * vhalf, vb = min(M, va + vhalf);
                float tv[3] = {0.f, 0.f, 0.f}; int ti[3] = {-1, -1, -1};
                for (int v = va; v < vb; ++v) {
                    const float val = stage[v * VOX_CTA + vl];
                    const uint4 nb = *reinterpret_cast<const uint4*>(s_nbr + v * NBR_W);
                    bool cand = val > 0.f;
                    const uint32_t w[4] = {nb.x, nb.y, nb.z, nb.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t n0 = w[k] & 0xFFFFu, n1 = w[k] >> 16;
                        if (n0 != NBR_NONE) cand = cand && (val > stage[n0 * VOX_CTA + vl]);
                        if (n1 != NBR_NONE) cand = cand && (val > stage[n1 * VOX_CTA + vl]);
                    }
                    if (cand) top3_insert(val, v, tv, ti);
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) { s_topv[(part * VOX_CTA + vl) * 3 + k] = tv[k]; s_topi[(part * VOX_CTA + vl) * 3 + k] = ti[k]; }
            }
            named_bar(1, EPI_THREADS);
            if (part == 0) {
                float tv[3], ti_f; int ti[3];
                (void)ti_f;
#pragma unroll
                for (int k = 0; k < 3; ++k) { tv[k] = s_topv[vl * 3 + k]; ti[k] = s_topi[vl * 3 + k]; }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int idx = s_topi[(VOX_CTA + vl) * 3 + k];
                    if (idx >= 0) top3_insert(s_topv[(VOX_CTA + vl) * 3 + k], idx, tv, ti);
                }
                const float omin = fminf(s_min[vl], s_min[VOX_CTA + vl]);
                const float osum = s_sum[vl] + s_sum[VOX_CTA + vl];
                float mean = osum / (float)M;
                const bool bad = vok && !(fabsf(osum) < CUDART_INF_F);   // fp16 overflow of the scaled signal (or non-finite input)
                if (vok) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const bool ok = ti[k] >= 0;
                        const int id = ok ? ti[k] : 0;
                        p.peak[k][vox]                   = ok ? __ldg(p.vert + id * 3 + 0) : 0.f;
                        p.peak[k][vox + p.out_pitch]     = ok ? __ldg(p.vert + id * 3 + 1) : 0.f;
                        p.peak[k][vox + 2 * p.out_pitch] = ok ? __ldg(p.vert + id * 3 + 2) : 0.f;
                        p.qa[k][vox] = ok ? tv[k] - omin : 0.f;
                        if (p.peak_idx) p.peak_idx[vox + k * p.out_pitch] = (int16_t)ti[k];
                    }
                }
                if (!vok || bad) mean = -CUDART_INF_F;
                const unsigned anybad = __ballot_sync(0xffffffffu, bad);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mean = fmaxf(mean, __shfl_xor_sync(0xffffffffu, mean, o));
                if (lane == 0) {
                    if (mean > -CUDART_INF_F) atomicMax(p.stats, f2ord(mean));
                    if (anybad) {                                       // recompute this 64-voxel tile with the SIMT kernel
                        const int slot = atomicAdd(p.fix_count, 1);
                        if (slot < p.fix_cap) p.fix_list[slot] = (int)(((int64_t)tile * 256 + rank * VOX_CTA + q * 32) >> 6);
                    }
                }
            }
            named_bar(1, EPI_THREADS);                                  // staging / scratch free for the next tile
        }
    }

    // ---- teardown --------------------------------------------------------------------------
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == W_MMA)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

size_t tc_smem_bytes(int M, int Nh) {
    size_t b = (size_t)NSTAGE * 2 * Nh * 32 + (size_t)M * VOX_CTA * 4 + 2 * VOX_CTA * 3 * 8 + 2 * VOX_CTA * 2 * 4 +
               (2 * NSTAGE + 2 * ASLOT + 2) * 8 + 16 + (size_t)M * NBR_W * 2;
    return b + 1024 + 64;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// Decide whether the tensor-core kernel can take this plan; build the split fp16 operand and its
// tensor map.  Returns 0 when usable.
int tc_plan_init(Plan* p) {
    if (p->kind != PLAN_GQI) { set_error("tensor-core path: GQI only (DSI uses the SIMT kernel)"); return 1; }
    const int M = p->nvert, K = p->nvol;
    const int Npad = (M + 15) / 16 * 16;
    if (Npad > TMEM_A_COL) { set_error("tensor-core path: more than 384 half-sphere vertices"); return 1; }
    int N1 = Npad, N2 = 0;
    if (Npad > 256) { N1 = (Npad / 2 + 15) / 16 * 16; N2 = Npad - N1; }
    const int Nh = (N1 + N2) / 2, N1h = N1 / 2, N2h = N2 / 2;
    if (Nh > 256) { set_error("tensor-core path: TMA box too tall"); return 1; }
    const size_t smem = tc_smem_bytes(M, Nh);
    int dev_smem = 0;
    if (cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, p->device) != cudaSuccess ||
        smem > (size_t)dev_smem) { set_error("tensor-core path: tile does not fit in shared memory"); return 1; }
    int cc_major = 0;
    cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, p->device);
    if (cc_major != 10) { set_error("tensor-core path needs sm_100"); return 1; }
    const int Kpad = (K + 31) / 32 * 32;
    // split operand, row order: rank 0 [hi: blk1 rows 0..N1h, blk2 rows 0..N2h][lo: same], then rank 1
    std::vector<__half> split((size_t)4 * Nh * Kpad, __float2half(0.f));
    auto row_of = [&](int n, int& rank, int& local) {
        if (n < N1) { rank = n / N1h; local = n % N1h; }
        else { int m = n - N1; rank = m / N2h; local = N1h + m % N2h; }
    };
    for (int n = 0; n < M; ++n) {
        int rank, local; row_of(n, rank, local);
        for (int k = 0; k < K; ++k) {
            const float a = p->h_matrix[(size_t)n * K + k];
            const __half h = __float2half_rn(a);
            const __half l = __float2half_rn(a - __half2float(h));
            split[((size_t)rank * 2 * Nh + local) * Kpad + k] = h;
            split[((size_t)rank * 2 * Nh + Nh + local) * Kpad + k] = l;
        }
    }
    TcState* st = new TcState();
    st->Kpad = Kpad; st->Npad = Npad; st->N1 = N1; st->N2 = N2; st->smem = smem;
    if (cudaMalloc(&st->d_split, split.size() * sizeof(__half)) != cudaSuccess ||
        cudaMemcpy(st->d_split, split.data(), split.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("tensor-core path: device allocation failed"); cudaGetLastError(); delete st; return 1;
    }
    void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
        set_error("tensor-core path: cuTensorMapEncodeTiled unavailable"); cudaGetLastError(); cudaFree(st->d_split); delete st; return 1;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)Kpad, (cuuint64_t)(4 * Nh)};
    cuuint64_t gstr[1] = {(cuuint64_t)Kpad * 2};
    cuuint32_t box[2] = {16, (cuuint32_t)Nh};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&st->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, st->d_split, gdim, gstr, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tensor-core path: cuTensorMapEncodeTiled failed"); cudaFree(st->d_split); delete st; return 1; }
    if (cudaFuncSetAttribute(recon_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        set_error("tensor-core path: cannot raise the shared-memory limit"); cudaGetLastError(); cudaFree(st->d_split); delete st; return 1;
    }
    p->tc = st;
    return 0;
}

void tc_plan_free(Plan* p) {
    TcState* st = reinterpret_cast<TcState*>(p->tc);
    if (!st) return;
    cudaFree(st->d_split); cudaFree(st->d_scratch);
    delete st;
    p->tc = nullptr;
}

int launch_recon_tc(Plan* p, const ReconArgs& a, cudaStream_t stream) {
    TcState* st = reinterpret_cast<TcState*>(p->tc);
    if (!st) return fail(FIBERS_ERR_ARG, "plan has no tensor-core state");
    if (a.nvox <= 0) return 0;
    if (a.nvox > 0x7FFFFFFFLL * 32) return fail(FIBERS_ERR_ARG, "slab too large");
    const int64_t ntile64 = (a.nvox + 63) / 64;
    const int64_t need = 2 + 2 * ntile64 + 8;
    if (st->scratch_cap < need) {
        if (st->d_scratch) cudaFree(st->d_scratch);
        st->d_scratch = nullptr; st->scratch_cap = 0;
        FB_CUDA(cudaMalloc(&st->d_scratch, sizeof(int) * (size_t)need));
        st->scratch_cap = need;
    }
    FB_CUDA(cudaMemsetAsync(st->d_scratch, 0, 2 * sizeof(int), stream));
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, p->device);
    sample_max_kernel<<<nsm * 8, 256, 0, stream>>>(a.dwi, a.dwi_pitch, a.nvox, p->nvol, st->d_scratch);
    TcParams tp{};
    tp.dwi = a.dwi; tp.dwi_pitch = a.dwi_pitch; tp.mask = a.mask; tp.nvox = a.nvox;
    tp.K = p->nvol; tp.Kpad = st->Kpad; tp.M = p->nvert; tp.Npad = st->Npad; tp.N1 = st->N1; tp.N2 = st->N2;
    tp.odf = a.odf; tp.out_pitch = a.out_pitch;
    for (int k = 0; k < 3; ++k) { tp.peak[k] = a.peak[k]; tp.qa[k] = a.qa[k]; }
    tp.peak_idx = a.peak_idx; tp.stats = a.stats; tp.nbr = p->d_nbr; tp.vert = p->d_vert;
    tp.maxbits = st->d_scratch; tp.fix_count = st->d_scratch + 1; tp.fix_list = st->d_scratch + 2;
    tp.fix_cap = (int)(2 * ntile64);
    tp.ntiles = (int)((a.nvox + 255) / 256);
    const int nclusters = std::max(1, std::min(nsm / 2, tp.ntiles));
    recon_tc_kernel<<<2 * nclusters, TC_THREADS, st->smem, stream>>>(tp, st->tmap);
    count_launch(2);
    FB_CUDA(cudaGetLastError());
    // voxels whose scaled signal overflowed fp16 (rare): recompute their 64-voxel tiles in fp32
    return launch_recon_simt_list(p, a, tp.fix_list, tp.fix_count, stream);
}

}  // namespace fibers
