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
//   warp 0      TMA producer: the split matrix lives in HBM as a ready-made shared-memory image (K32 stage
//               after K32 stage, SWIZZLE_32B pattern applied on the host), so a stage of this CTA's half of the
//               rows is ONE bulk tensor copy into a 3-stage ring
//   warp 1      MMA issuer (leader CTA): tcgen05.mma.cta_group::2, A operand from TENSOR MEMORY,
//               B from shared memory; accumulators D[128 x Npad] fp32 in TMEM columns [0, Npad)
//   warps 2-5   converters (one per TMEM lane quarter): the raw DWI slab is staged through a 3-stage
//               shared-memory ring with cp.async (16-byte copies when the rows are 16-byte aligned), two chunks
//               ahead of the conversion; clamp, scale, hi/lo fp16 split, tcgen05.st into a 4-slot TMEM ring
//               (columns 384..511)
//   warps 6-17  epilogue (three groups of four): tcgen05.ld, un-scale, coalesced ODF store; every value is also
//               quantised to a 15-bit ORDER-PRESERVING key (fixed point relative to the voxel's mean ODF, which
//               the MMA delivers as one extra matrix row) and staged as a 128 x M key tile in shared memory.
//               The local-maximum scan of the folded mesh runs on the keys (4 voxels per thread, packed u16x2
//               max, one 32-bit subtraction tests two voxels, neighbour offsets from constant memory) and only
//               LISTS possible maxima; the few listed (voxel, vertex) pairs are then settled EXACTLY on the
//               fp32 values just written (L2 hits) and inserted into the voxel's sorted triple by chained
//               64-bit shared atomicMax on (value, ~index); QA, per-voxel mean -> atomicMax
// The full ODF never round-trips HBM: it is written once; the scan reads 2 bytes per value from shared memory.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math_constants.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include "common.cuh"

namespace fibers {

int launch_recon_simt_list(Plan* p, const ReconArgs& a, const int* d_list, const int* d_count, cudaStream_t st);

namespace {

constexpr int N_EPI = 12;            // epilogue warps: 3 groups of 4 (one warp per TMEM lane quarter); group k also writes peak k
constexpr int N_CONV = 4;            // converter warps (multiple of 4)
constexpr int W_MMA = 1, W_DWI = 2, W_CONV0 = 3, W_EPI0 = W_CONV0 + N_CONV;   // (every group of 4 consecutive warps covers the 4 TMEM lane quarters: quarter = warp & 3)
constexpr int TC_THREADS = (W_EPI0 + N_EPI) * 32;
constexpr int N_CPART = N_EPI / 4;   // column parts in the TMEM drain
constexpr int NSTAGE = 8;            // B ring, K16 sub-tiles (hi rows + lo rows): maximum depth; the launch picks what fits (p.nstage)
constexpr int DBOXW_SPLIT = 132;     // TMA mode, rows not 16-byte aligned: voxels per box row (128 + the <= 3 voxels the map base was rounded down by)
__host__ __device__ constexpr int dwi_box_bytes(int nmap) { return nmap > 1 ? ((16 / nmap * DBOXW_SPLIT * 4 + 127) / 128) * 128 : 16 * 128 * 4; }   // shared-memory slot of one box
__host__ __device__ constexpr int dwi_stage_bytes(int nmap) { return (nmap > 1 ? nmap : 1) * dwi_box_bytes(nmap); }                                   // one ring stage: 16 volumes x 128 voxels
constexpr int DSTAGE = 3;            // cp.async mode: raw DWI ring of the converters, K32 chunks of 128 voxels (16 KB each)
constexpr int TSTAGE = 8;            // TMA mode: raw DWI ring, K16 boxes of 128 voxels (8 KB each), filled by the DWI producer warp: maximum depth (p.tstage)
constexpr int OBOX_BYTES = 16 * 128 * 4;   // TMA mode: ODF staging box = 16 vertex rows x 128 voxels fp32, one TMA store each
constexpr int ASLOT = 4;             // A ring in TMEM: K32 chunks, 32 columns each
constexpr int TMEM_A_COL = 384;
constexpr int VOX_CTA = 128;
constexpr int EPI_THREADS = N_EPI * 32;
static_assert(N_EPI == 12, "the output stage maps warp group k to peak k");
constexpr float FP16_TARGET = 8192.f;   // the sampled maximum is scaled to <= 8192 (8x headroom to 65504)

constexpr int KEY_ROW = VOX_CTA * 2;    // bytes per vertex row of the key tile
constexpr int CAND_CAP = 1024;          // listed (voxel, vertex) pairs per 128-voxel tile (typically ~170); overflow -> SIMT fix-up
constexpr float KEY_WINDOW = 8.f;       // keys cover [0, 8 x mean(ODF of the voxel)) in 32767 steps (bit 15 is always set)

// Folded-mesh neighbour table as a KERNEL PARAMETER (constant bank 0, private to the launch): byte offsets
// (vertex * KEY_ROW) into the key tile, 8 per vertex (missing neighbours -> the all-zero sentinel row M).  The
// vertex index is warp-uniform, so the offsets arrive through the uniform datapath and each neighbour costs one
// LDS.64.  (A device-global __constant__ symbol here would be shared by every plan and stream of the device:
// two plans with different meshes running concurrently would read each other's table.)
constexpr int TC_MAX_VERT = 392;        // rows of the offset table: M + 1 sentinel rows up to M + 8 (prefetch overrun)
struct NbrOffTable { uint32_t off[TC_MAX_VERT * NBR_W]; };   // 12.5 KB of the 32 KB parameter space

struct TcParams {
    const float* dwi; int64_t dwi_pitch; const uint8_t* mask; int64_t nvox;
    int K, Kpad;                 // Kpad: multiple of 32
    int M, Npad, N1, N2;         // D columns: [0,N1) block 1, [N1, N1+N2) block 2; N1, N2 multiples of 16
    float* odf; int64_t out_pitch; float* peak[3]; float* qa[3]; int16_t* peak_idx; int32_t* stats;
    const uint16_t* nbr; const float* vert;
    const int* maxbits;          // device: bit pattern of the sampled max(s) (>= 0)
    int* fix_list; int* fix_count; int fix_cap;
    int ntiles;                  // 256-voxel tiles of the slab
    const int* tile_list; const int* tile_count;   // tiles that contain at least one mask voxel (built by tile_scan_kernel)
    int nbw;                     // max neighbour count of the folded mesh (<= 8)
    int nstage;                  // depth of the B ring (<= NSTAGE)
    int tstage;                  // TMA mode: depth of the raw DWI ring (<= TSTAGE, even)
    int obuf;                    // TMA mode: ODF staging boxes per epilogue warp (1 or 2); 0 = direct stores, no staging
    int abl;                     // ABLATION bits for timing experiments only (FIBERS_TC_ABLATE; results are then wrong by design)
    int l2pf;                    // TMA mode: the DWI producer prefetches the boxes of the tile l2pf rounds ahead into L2 (0 = off)
    int plain;                   // 1: rows are stored only (DSI pdf rows): no peak search, no statistics
    int cvol; float dscale;      // DSI: every output row is divided by den = dscale * max(s[cvol], 0)  (cvol < 0: none)
    int dwi_vec;                 // DWI staging copies: 2 = 16 bytes (base 16-byte aligned, pitch % 4 == 0), 1 = 8 bytes, 0 = 4 bytes
    int dshift[4];               // TMA mode, split maps: voxel coordinate of voxel 0 in DWI map r (its base is rounded down to 16 bytes)
    uint32_t conv_sleep, prod_sleep;   // back-off (ns) of the converters' a_empty wait and of the producers' ring waits
    int cand_cap;                // capacity of the candidate list (<= CAND_CAP; tests shrink it to force the fall-back)
    long long* trace;            // optional per-role clock trace of cluster 0 / CTA 0 (debug; FIBERS_TC_TRACE)
    uint32_t trace_skip;         // first traced tile iteration
};

// DWI slab maps of one launch.  TMA wants a 16-byte aligned base, row strides that are multiples of 16 bytes AND box
// coordinates whose byte offset along the row is a multiple of 16 (measured on B200: an fp32 box at voxel coordinate
// 4n + 1 or 4n + 2 raises cudaErrorIllegalInstruction, 4n + 4 is fine).  A slab whose frame pitch is the bare voxel
// count (what a device-resident caller of the reference layout has: rows 8- or 4-byte aligned) is therefore described
// by kTma = 2 or 4 maps: map r holds the volumes k = r (mod kTma), so its row stride kTma * pitch is a multiple of 16
// bytes again; its base is the address of volume r rounded DOWN to 16 bytes, and its boxes are 132 voxels wide and
// start at the tile's (aligned) voxel coordinate: the tile's 128 voxels sit TcParams::dshift[r] = 0..3 voxels into
// each row of the box.  The boxes of the kTma maps land in one ring stage as [r][16 / kTma][132]; the converters read
// the rows in volume order.
struct DwiMaps { CUtensorMap m[4]; };

// One launch of the kernel covers at most 384 matrix rows (TMEM columns 0..383; the A ring sits above).  GQI: a single pass (the ODF
// rows).  DSI: the ODF rows, then the pdf rows in passes of <= 336 (plain mode).
struct TcPass {
    __half* d_split = nullptr;   // [2 ranks][K32 chunks][shared-memory image of one stage]
    CUtensorMap tmap;
    int rows = 0, row0 = 0;      // matrix rows [row0, row0 + rows) of the plan's matrix
    int Npad = 0, N1 = 0, N2 = 0;
    int plain = 0;               // 1: pdf rows
};

struct TcState {
    std::vector<TcPass> pass;
    void* encode = nullptr;      // cuTensorMapEncodeTiled
    int Kpad = 0, nbw = 8;
    NbrOffTable nbr_off;                     // [M + 8][NBR_W] byte offsets (rows >= M: sentinel); passed by value with every launch
    int dev_smem = 0;            // opt-in shared memory per block of the device
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
// (relaxed: the data these arrivals publish lives in tensor memory and is ordered by tcgen05.wait /
//  tcgen05.fence, not by the generic-proxy memory model; a release.cluster costs a full MEMBAR)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
// kBackoff: non-critical waiters sleep between polls so that they do not steal issue slots from the working warps
// (measured: a polling loop without the sleep issues ~10 k instructions per tile and role)
template <int kSleepNs = 0>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (kSleepNs > 0 && !ok) __nanosleep(kSleepNs);
        if (spin > (1u << 26)) __trap();        // never hang the GPU: a lost signal becomes a launch error
    }
}
// run-time back-off (tuning experiments): ns == 0 -> hardware-suspended try_wait with a long time hint, no polling loop
__device__ __forceinline__ void mbar_wait_ns(uint64_t* bar, uint32_t parity, uint32_t ns) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        if (ns == 0)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(a), "r"(parity), "r"(20000u) : "memory");
        else {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(a), "r"(parity) : "memory");
            if (!ok) __nanosleep(ns);
        }
        if (spin > (1u << 26)) __trap();
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
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
    uint32_t r; asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory"); return r;
}
__device__ __forceinline__ uint32_t umax2(uint32_t a, uint32_t b) { uint32_t d; asm("max.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
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

// (16 consecutive tiles of CTA 0, starting at tile iteration p.trace_skip: FIBERS_TC_TRACE_SKIP, default 0)
#define TRACE(slot) do { if (kTrace && p.trace && blockIdx.x == 0 && lane == 0 && it - p.trace_skip < 16u) p.trace[(it - p.trace_skip) * 32 + (slot)] = clock64(); } while (0)
#define TRACE_ADD(slot, dt) do { if (kTrace && p.trace && blockIdx.x == 0 && lane == 0 && it - p.trace_skip < 16u) p.trace[(it - p.trace_skip) * 32 + (slot)] += (dt); } while (0)

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
// tile scan: one warp per 256-voxel tile.  Tiles without a single mask voxel are zero-filled here (the
// reference leaves them at the zero of its MRI constructor) and never reach the reconstruction kernel;
// the others are counted and, if any tile is empty, compacted into the work list.  On a brain-masked
// volume this removes most of the tiles.
// ---------------------------------------------------------------------------------------------
struct ScanOut { float* ptr[9]; int rows[9]; int n; int16_t* idx; };

// one WARP per 256-voxel tile (8 mask bytes per lane), 8 tiles per CTA
__global__ void __launch_bounds__(256) tile_scan_kernel(const uint8_t* __restrict__ mask, int64_t nvox, int64_t out_pitch, ScanOut o,
                                                         int32_t* stats, int* tile_flag, int ntiles, int* tile_count) {
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (tile >= ntiles) return;
    const int64_t v0 = (int64_t)tile * 256 + lane * 8;
    unsigned long long m = 0;
    if (v0 + 8 <= nvox && (reinterpret_cast<uintptr_t>(mask) & 7) == 0) m = *reinterpret_cast<const unsigned long long*>(mask + v0);
    else for (int j = 0; j < 8; ++j) if (v0 + j < nvox && mask[v0 + j]) m |= 1ull << (8 * j);
    const bool any = __any_sync(0xffffffffu, m != 0ull);
    if (lane == 0) { tile_flag[tile] = any ? 1 : 0; if (any) atomicAdd(tile_count, 1); }
    if (any) return;
    for (int a = 0; a < o.n; ++a)
        for (int r = 0; r < o.rows[a]; ++r)
            for (int j = lane; j < 256; j += 32) { const int64_t v = (int64_t)tile * 256 + j; if (v < nvox) o.ptr[a][(int64_t)r * out_pitch + v] = 0.f; }
    if (o.idx)
        for (int r = 0; r < 3; ++r)
            for (int j = lane; j < 256; j += 32) { const int64_t v = (int64_t)tile * 256 + j; if (v < nvox) o.idx[(int64_t)r * out_pitch + v] = (int16_t)-1; }
    if (lane == 0 && stats) atomicMax(stats, f2ord(0.f));      // skipped voxels count as mean(odf) = 0 in odfmax
}

// ordered compaction of the non-empty tiles (one block; ascending tile order keeps the DWI / ODF streams of
// neighbouring clusters adjacent in memory).  Nothing to do when no tile is empty: the kernel then walks the
// identity.
__global__ void __launch_bounds__(1024) tile_compact_kernel(const int* __restrict__ flag, int ntiles, int* list, const int* count) {
    if (*count >= ntiles) return;
    __shared__ int s_part[1024];
    const int per = (ntiles + 1023) / 1024;
    const int t0 = threadIdx.x * per, t1 = min(ntiles, t0 + per);
    int c = 0;
    for (int t = t0; t < t1; ++t) c += flag[t];
    s_part[threadIdx.x] = c;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                  // inclusive scan
        int v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    int pos = s_part[threadIdx.x] - c;
    for (int t = t0; t < t1; ++t) if (flag[t]) list[pos++] = t;
}

// ---------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {          // one lane of a converged warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// kTma: the DWI slab arrives by TMA (K16 x 128-voxel boxes, issued by the DWI producer warp) and the ODF tile leaves by
// TMA stores from a small staging ring, so that the accumulator drain is not paced by the SM's 32 B/clk store port
// (tools/store_probe.cu: 172 KB of ODF per tile = 5.4 k cycles of that port); needs 16-byte aligned slab / output
// rows.  !kTma: cp.async staging + direct stores (any alignment).
// kTrace: per-role clock trace of cluster 0 (FIBERS_TC_TRACE); compiled out of the production instantiations.
template <int kTma, bool kTrace>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
recon_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap tmapB, const __grid_constant__ NbrOffTable nbt,
                const __grid_constant__ DwiMaps tmapD, const __grid_constant__ CUtensorMap tmapO) {
    constexpr int NMAP = kTma ? kTma : 1;
    constexpr int DROWS = 16 / NMAP;                                      // rows of one DWI box (NMAP boxes per ring stage)
    constexpr int DBOXW = kTma > 1 ? DBOXW_SPLIT : VOX_CTA;               // voxels per box row
    constexpr uint32_t DSLOT = dwi_box_bytes(NMAP), DSTG = dwi_stage_bytes(NMAP);
    const uint32_t* const c_nbr_off = nbt.off;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // the warp index is rebuilt from warp votes so that the compiler can prove it warp-uniform (uniform registers,
    // uniform branches, constant-bank loads and [R + UR] addressing in the role loops)
    int warp = 0;
#pragma unroll
    for (int b = 0; b < 5; ++b) warp |= (__ballot_sync(0xffffffffu, (threadIdx.x >> (5 + b)) & 1u) & 1u) << b;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int cluster_id = blockIdx.x >> 1, ncluster = gridDim.x >> 1;
    const int Nh = (p.N1 + p.N2) >> 1, N1h = p.N1 >> 1;
    const uint32_t sub_bytes = (uint32_t)(2 * Nh * 32);          // one K16 sub-tile = one B stage: hi rows then lo rows (SWIZZLE_32B)
    const int nk32 = p.Kpad >> 5, nk16 = p.Kpad >> 4;
    const uint32_t stage_rows = sub_bytes >> 9;                  // 512-byte rows of the pre-tiled global image per stage
    const int ntl = min(*p.tile_count, p.ntiles);      // non-empty tiles; every role walks the same list
    const bool ident = ntl == p.ntiles;                // nothing skipped: the list is the identity

    // ---- shared memory carve-up -------------------------------------------------------------
    // (pointer arithmetic on the __shared__ array itself, so that the compiler keeps the shared
    //  address space and emits LDS/STS instead of generic LD/ST)
    //  The kernel has no static shared memory, so the 1024-byte alignment requested on the extern
    //  array holds for the dynamic window; a misaligned base would corrupt the swizzled tiles: trap.
    uint8_t* base = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
    uint8_t* sB = base;                                                   // nstage * sub_bytes
    uint16_t* keys = (uint16_t*)(sB + p.nstage * sub_bytes);                // [M + 1][128] 16-bit keys; row M = 0 (sentinel)
    const int Mk = p.plain ? 0 : p.M;                                     // plain passes stage no keys
    // (double-buffered: the settle step and the outputs of a tile run one loop iteration later, beside the next tile's drain)
    unsigned long long* s_top = (unsigned long long*)(keys + (size_t)(Mk + 8) * VOX_CTA);   // [2][3][128] best (value, ~index) per voxel
    uint32_t* s_cand = (uint32_t*)(s_top + 2 * 3 * VOX_CTA);              // [2][CAND_CAP] listed (vertex, voxel)
    float* s_min = (float*)(s_cand + 2 * CAND_CAP);                       // [2][N_CPART][128]
    float* s_mean = s_min + 2 * N_CPART * VOX_CTA;                        // [2][128] mean ODF per voxel (from the extra matrix row)
    // raw DWI ring (128-byte aligned: TMA destination), then (TMA mode) the ODF staging boxes
    float* s_dwi = (float*)(((uintptr_t)(s_mean + 2 * VOX_CTA) + 127) & ~(uintptr_t)127);   // cp.async: [DSTAGE][32][128]; TMA: [TSTAGE][16][128]
    uint8_t* s_obox = (uint8_t*)s_dwi + (kTma ? (size_t)p.tstage * DSTG : (size_t)DSTAGE * 32 * VOX_CTA * 4);   // [N_CPART][obuf][16][128] fp32
    uint4* s_nbr = (uint4*)(s_obox + (kTma ? (size_t)N_CPART * p.obuf * OBOX_BYTES : 0));  // [M] 8 x uint16 neighbour ids per vertex
    float* s_vert = (float*)(s_nbr + Mk);                                 // [M][3] first-half vertices (peak vectors); padded to 4 floats
    uint64_t* bars = (uint64_t*)(s_vert + ((3 * Mk + 3) & ~3));
    uint64_t* b_full = bars, *b_empty = bars + NSTAGE, *a_full = bars + 2 * NSTAGE, *a_empty = bars + 2 * NSTAGE + ASLOT;
    uint64_t* d_full = bars + 2 * NSTAGE + 2 * ASLOT, *d_empty = d_full + 1;
    uint64_t* w_full = d_empty + 1, *w_empty = w_full + TSTAGE;           // DWI ring (TMA mode)
    uint32_t* s_ncand = (uint32_t*)(w_empty + TSTAGE);                    // [2]
    uint32_t* tmem_ptr_s = s_ncand + 2;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < ASLOT; ++i) { mbar_init(&a_full[i], 8); mbar_init(&a_empty[i], 1); }   // 4 converter warps x 2 CTAs
        for (int i = 0; i < TSTAGE; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], N_CONV); }
        mbar_init(d_full, 1); mbar_init(d_empty, 2 * N_EPI);                                      // epilogue warps x 2 CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 8 * VOX_CTA; i += TC_THREADS) keys[(size_t)Mk * VOX_CTA + i] = 0x8000;   // key 0
    if (threadIdx.x == 0) { s_ncand[0] = 0u; s_ncand[1] = 0u; }
    for (int i = threadIdx.x; i < Mk; i += TC_THREADS) s_nbr[i] = __ldg(reinterpret_cast<const uint4*>(p.nbr) + i);
    for (int i = threadIdx.x; i < 3 * Mk; i += TC_THREADS) s_vert[i] = __ldg(p.vert + i);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;
    if (kTrace && p.trace && threadIdx.x == 0 && rank == 0) {           // per-cluster wall-clock span (debug trace only)
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[16 * 32 + 2 * cluster_id] = (long long)t;
        if (cluster_id == 0) p.trace[16 * 32 + 2 * 127] = clock64();            // SM cycles over the same span -> effective clock
    }

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
        // ===== TMA producer: this CTA's half of the split matrix rows, K32 per stage ===========
        // (the whole warp runs the loop so that control flow stays uniform; one elected lane issues)
        const uint32_t full0 = mapa(smem_u32(&b_full[0]), 0);
        uint32_t it = 0;
        int s = 0; uint32_t ph = 0;                         // ring stage and its phase bit (no run-time division in the loop)
        for (int ti_ = cluster_id; ti_ < ntl; ti_ += ncluster, ++it) {
            TRACE(13);
            for (int c = 0; c < nk16; ++c) {
                mbar_wait_ns(&b_empty[s], ph ^ 1, p.prod_sleep);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(&b_full[s], 2 * sub_bytes);         // both CTAs' bytes land on the leader's barrier
                    // one bulk tensor copy per stage: the global image is already in shared-memory order
                    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(smem_u32(sB + s * sub_bytes)), "l"(&tmapB), "r"(0), "r"((int)((rank * nk16 + c) * stage_rows)),
                                   "r"(full0 + s * 8) : "memory");
                }
                __syncwarp();
                if (++s == p.nstage) { s = 0; ph ^= 1u; }
            }
            TRACE(14);
        }
    } else if (warp == W_MMA) {
        // ===== MMA issuer (leader CTA; warp-uniform loop, one elected lane issues) ==============
        if (rank == 0) {
            const uint32_t idesc1 = make_idesc_f16(256, p.N1);
            const uint32_t idesc2 = p.N2 ? make_idesc_f16(256, p.N2) : 0u;
            uint32_t it = 0;
            int bs = 0; uint32_t bph = 0;                   // B ring stage / phase
            int slot = 0; uint32_t aph = 0;                 // A ring slot / phase
            // shared-memory descriptors of stage 0 (the start-address field is additive: + (stage * sub_bytes) >> 4)
            const uint32_t sB0 = smem_u32(sB);
            const uint64_t dhi1 = make_sdesc_sw32(sB0), dlo1 = make_sdesc_sw32(sB0 + Nh * 32);
            const uint64_t dhi2 = make_sdesc_sw32(sB0 + N1h * 32), dlo2 = make_sdesc_sw32(sB0 + Nh * 32 + N1h * 32);
            const uint32_t dstep = sub_bytes >> 4;
            const bool two = p.N2 != 0, one_product = (p.abl & 1) != 0;
            for (int ti_ = cluster_id; ti_ < ntl; ti_ += ncluster, ++it) {             // (this role only needs the tile count)
                mbar_wait<100>(d_empty, (it & 1) ^ 1);              // epilogue of the previous tile has drained TMEM
                tc_fence_after();
                TRACE(0);
                long long wait_a = 0, wait_b = 0;                       // (trace only)
                for (int c = 0; c < nk32; ++c) {
                    long long t1 = kTrace ? clock64() : 0;
                    mbar_wait(&a_full[slot], aph);
                    if (kTrace) wait_a += clock64() - t1;
                    const uint32_t a_base = tmem_base + TMEM_A_COL + slot * 32;
#pragma unroll
                    for (int sub = 0; sub < 2; ++sub) {
                        long long t0 = kTrace ? clock64() : 0;
                        mbar_wait(&b_full[bs], bph);
                        if (kTrace) wait_b += clock64() - t0;
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t a_hi = a_base + sub * 8, a_lo = a_hi + 16;
                            const uint32_t acc = (c | sub) ? 1u : 0u;
                            const uint64_t off = (uint64_t)((uint32_t)bs * dstep);
                            if (!one_product) {
                                mma_ts2(tmem_base, a_lo, dhi1 + off, idesc1, acc);       // small terms first
                                mma_ts2(tmem_base, a_hi, dlo1 + off, idesc1, 1u);
                                mma_ts2(tmem_base, a_hi, dhi1 + off, idesc1, 1u);
                                if (two) {
                                    mma_ts2(tmem_base + p.N1, a_lo, dhi2 + off, idesc2, acc);
                                    mma_ts2(tmem_base + p.N1, a_hi, dlo2 + off, idesc2, 1u);
                                    mma_ts2(tmem_base + p.N1, a_hi, dhi2 + off, idesc2, 1u);
                                }
                            } else {                                    // (timing ablation)
                                mma_ts2(tmem_base, a_hi, dhi1 + off, idesc1, acc);
                                if (two) mma_ts2(tmem_base + p.N1, a_hi, dhi2 + off, idesc2, acc);
                            }
                            mma_commit2(&b_empty[bs]);                     // B stage reusable once these MMAs retire
                            if (sub == 1) {
                                mma_commit2(&a_empty[slot]);               // ... and the A slot after its second half
                                if (c == nk32 - 1) mma_commit2(d_full);
                            }
                        }
                        __syncwarp();
                        if (++bs == p.nstage) { bs = 0; bph ^= 1u; }
                    }
                    if (++slot == ASLOT) { slot = 0; aph ^= 1u; }
                }
                TRACE(1);
                if (kTrace) { TRACE_ADD(15, wait_b); TRACE_ADD(16, wait_a); }
            }
        }
    } else if (warp == W_DWI) {
        // ===== DWI producer (TMA mode): one K16 x 128-voxel box of the raw slab per stage, TSTAGE stages ahead of the
        //       converters across tile boundaries.  Voxels past the end of the slab and volumes past K arrive as zeros. ====
        if (kTma) {
            int s = 0; uint32_t ph = 0;
            for (int ti_ = cluster_id; ti_ < ntl; ti_ += ncluster) {
                const int tile = ident ? ti_ : __ldg(p.tile_list + ti_);
                const int vox0 = tile * 256 + (int)rank * VOX_CTA;
                for (int c = 0; c < nk16; ++c) {
                    mbar_wait_ns(&w_empty[s], ph ^ 1, p.prod_sleep);
                    if (elect_one()) {
                        if (p.l2pf > 0 && ti_ + p.l2pf * ncluster < ntl) {     // the same box of a later tile of this cluster -> L2
                            const int t2 = ident ? ti_ + p.l2pf * ncluster : __ldg(p.tile_list + ti_ + p.l2pf * ncluster);
#pragma unroll
                            for (int r = 0; r < kTma; ++r)
                                asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
                                             ::"l"(&tmapD.m[r]), "r"(t2 * 256 + (int)rank * VOX_CTA), "r"(c * DROWS) : "memory");
                        }
                        mbar_expect_tx(&w_full[s], 16 * DBOXW * 4);
#pragma unroll
                        for (int r = 0; r < kTma; ++r)
                            asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                         ::"r"(smem_u32(s_dwi) + s * DSTG + r * DSLOT), "l"(&tmapD.m[r]), "r"(vox0), "r"(c * DROWS),
                                           "r"(smem_u32(&w_full[s])) : "memory");
                    }
                    __syncwarp();
                    if (++s == p.tstage) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp < W_EPI0) {
        // ===== converters: DWI fp32 -> clamp -> scale -> fp16 hi/lo -> TMEM ring ================
        // One warp per TMEM lane quarter; thread == voxel.  The raw samples are staged through a shared-memory ring
        // with cp.async, DSTAGE - 1 chunks (across tile boundaries) ahead of the conversion, so no registers are
        // held while the loads are in flight.  Aligned slabs (16-byte base, pitch % 4 == 0) move 4 voxels per
        // copy: warp w fetches volumes 8w .. 8w+7 of every chunk for all 128 voxels and a 128-thread barrier
        // publishes the chunk; otherwise every thread copies its own voxel 4 bytes at a time.
        const int cw = warp - W_CONV0, q = warp & 3;
        const int vl = q * 32 + lane;                                   // TMEM lane == voxel within the CTA's 128
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t afull0 = mapa(smem_u32(&a_full[0]), 0);
        const uint32_t sd0 = smem_u32(s_dwi);                           // [DSTAGE][32][128] floats
        constexpr int PF = DSTAGE - 1;
        constexpr uint32_t STAGE_B = 32 * VOX_CTA * 4;
        const int vec = p.dwi_vec;
        uint32_t vb[NMAP];                                              // TMA mode: this voxel's column in the box of map r
#pragma unroll
        for (int r = 0; r < NMAP; ++r) vb[r] = sd0 + r * DSLOT + (vl + (kTma > 1 ? p.dshift[r] : 0)) * 4;
        // prefetch cursor
        int p_ti = cluster_id, p_c = 0; uint32_t p_g = 0;
        int64_t p_vox0 = 0;
        auto prefetch = [&]() {
            if (p_ti < ntl) {
                if (p_c == 0) {
                    const int ptile = ident ? p_ti : __ldg(p.tile_list + p_ti);
                    p_vox0 = (int64_t)ptile * 256 + rank * VOX_CTA;
                }
                const uint32_t dst = sd0 + (p_g % DSTAGE) * STAGE_B;
                if (vec == 2) {
                    const int64_t v4 = p_vox0 + 4 * lane;
                    const int64_t left = p.nvox - v4;
                    const uint32_t full = left >= 4 ? 16u : (left > 0 ? (uint32_t)left * 4u : 0u);
                    const float* src = p.dwi + (full ? v4 : 0) + (int64_t)(p_c * 32 + cw * 8) * p.dwi_pitch;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int k = p_c * 32 + cw * 8 + j;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
                                     ::"r"(dst + (uint32_t)(cw * 8 + j) * (VOX_CTA * 4) + lane * 16), "l"(k < p.K ? src : p.dwi), "r"(k < p.K ? full : 0u) : "memory");
                        src += p.dwi_pitch;
                    }
                } else if (vec == 1) {                                  // rows 8-byte aligned: two voxels per copy, two copies per volume
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int64_t v2 = p_vox0 + 64 * half + 2 * lane;
                        const int64_t left = p.nvox - v2;
                        const uint32_t full = left >= 2 ? 8u : (left > 0 ? 4u : 0u);
                        const float* src = p.dwi + (full ? v2 : 0) + (int64_t)(p_c * 32 + cw * 8) * p.dwi_pitch;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int k = p_c * 32 + cw * 8 + j;
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;"
                                         ::"r"(dst + (uint32_t)(cw * 8 + j) * (VOX_CTA * 4) + half * 256 + lane * 8), "l"(k < p.K ? src : p.dwi), "r"(k < p.K ? full : 0u) : "memory");
                            src += p.dwi_pitch;
                        }
                    }
                } else {
                    const int64_t pvox = p_vox0 + vl;
                    const bool inb = pvox < p.nvox;
                    const float* src = p.dwi + (inb ? pvox : 0) + (int64_t)(p_c * 32) * p.dwi_pitch;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int k = p_c * 32 + j;
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;"
                                     ::"r"(dst + (uint32_t)j * (VOX_CTA * 4) + vl * 4), "l"(k < p.K ? src : p.dwi), "r"((inb && k < p.K) ? 4u : 0u) : "memory");
                        src += p.dwi_pitch;
                    }
                }
                if (++p_c == nk32) { p_c = 0; p_ti += ncluster; }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");         // (an empty group keeps the group count in step)
            ++p_g;
        };
        if (!kTma) {
#pragma unroll 1
            for (int i = 0; i < PF; ++i) prefetch();
        }
        uint32_t it = 0, g32 = 0;
        int ws = 0; uint32_t wph = 0;                       // DWI ring stage / phase (TMA mode; the depth is even)
        int slot = 0; uint32_t aph = 0;                     // A ring slot / phase
        for (int ti_ = cluster_id; ti_ < ntl; ti_ += ncluster, ++it) {
            const int tile = ident ? ti_ : __ldg(p.tile_list + ti_);
            const int64_t vox = (int64_t)tile * 256 + rank * VOX_CTA + vl;
            const bool inside = vox < p.nvox && p.mask[vox] != 0;
            const float vscale = inside ? scale : 0.f;
            if (warp == W_CONV0) TRACE(9);
#pragma unroll 1
            for (int c = 0; c < nk32; ++c, ++g32) {
                float x[32];
                if (p.abl & 8) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = 1.f;
                    if (kTma) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) mbar_wait<20>(&w_full[ws + h], wph);
                    } else {
                        asm volatile("cp.async.wait_group %0;" ::"n"(PF - 1) : "memory");
                        named_bar(2, N_CONV * 32);
                        prefetch();
                    }
                } else if (kTma) {
                    // two K16 boxes of the DWI ring per A slot; the stage goes back to the producer once this warp has its values
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        mbar_wait<20>(&w_full[ws + h], wph);
                        const uint32_t so = (ws + h) * DSTG;
#pragma unroll
                        for (int j = 0; j < 16; ++j)                    // volume j of the stage: row j / kTma of the box of map j % kTma
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[16 * h + j]) : "r"(vb[j % NMAP] + so + (j / NMAP) * (DBOXW * 4)));
                    }
                } else {
                    asm volatile("cp.async.wait_group %0;" ::"n"(PF - 1) : "memory");   // this thread's copies of chunk g32 have landed
                    named_bar(2, N_CONV * 32);                              // ... everybody's have; chunk g32 - 1 is no longer read
                    prefetch();                                             // refills the stage chunk g32 - 1 used
                    const uint32_t sbase = sd0 + (g32 % DSTAGE) * STAGE_B + vl * 4;
#pragma unroll
                    for (int j = 0; j < 32; ++j) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[j]) : "r"(sbase + j * (VOX_CTA * 4)));
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    // s[s<0] = 0; voxels outside the mask contribute zeros: their scale factor is 0 (fmaxf drops a NaN; a +Inf
                    // sample of a masked-out voxel gives a non-finite accumulator, which sends the tile to the exact fix-up)
                    const float v0 = fmaxf(x[2 * j], 0.f) * vscale, v1 = fmaxf(x[2 * j + 1], 0.f) * vscale;
                    const __half2 h = __floats2half2_rn(v0, v1);
                    const float2 hf = __half22float2(h);
                    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                    hi[j] = *reinterpret_cast<const uint32_t*>(&h);
                    lo[j] = *reinterpret_cast<const uint32_t*>(&l);
                }
                if (kTma) {                                             // (the conversions above consumed every x[]: the loads have completed)
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&w_empty[ws])) : "memory");
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&w_empty[ws + 1])) : "memory");
                    }
                    ws += 2; if (ws == p.tstage) { ws = 0; wph ^= 1u; }
                }
                if (warp == W_CONV0 && c == 0) TRACE(10);
                mbar_wait_ns(&a_empty[slot], aph ^ 1, p.conv_sleep);
                if (warp == W_CONV0 && c == 0) TRACE(11);
                tc_fence_after();
                const uint32_t col = lane_addr + TMEM_A_COL + slot * 32;
                tmem_st8(col, hi); tmem_st8(col + 8, hi + 8); tmem_st8(col + 16, lo); tmem_st8(col + 24, lo + 8);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(afull0 + slot * 8);
                if (++slot == ASLOT) { slot = 0; aph ^= 1u; }
            }
            if (warp == W_CONV0) TRACE(12);
        }
        if (!kTma) asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        // ===== epilogue ==========================================================================
        // phase 1 roles: warp quarter q owns TMEM lanes 32q..32q+31, `cpart` selects the column range
        // phase 2 roles: N_EPI vertex ranges (one per warp); lane owns voxels 4*lane .. 4*lane+3 (4 packed keys)
        const int ew = warp - W_EPI0, cpart = ew >> 2, q = warp & 3;
        const int et = ew * 32 + lane;                                  // 0..255 within the epilogue group
        const int vl = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t dempty0 = mapa(smem_u32(d_empty), 0);
        const int M = p.M;
        const int cper = ((p.Npad + N_CPART - 1) / N_CPART + 15) & ~15;
        const int c_begin = cpart * cper, c_end = min(min(c_begin + cper, p.Npad), (M + 15) & ~15);
        // TMA mode: this warp's ODF staging boxes ([16 rows][32 voxels] fp32 each) and its running box counter
        constexpr uint32_t WBOX = 16 * 32 * 4;
        const uint32_t obox0 = smem_u32(s_obox) + (uint32_t)(ew * p.obuf) * WBOX;
        const uint32_t obufm = (uint32_t)max(p.obuf, 1);
        uint32_t oc = 0;
        // The epilogue of a tile is SPLIT across two loop iterations so that the L2 round trips of the settle step hide
        // behind the next tile's accumulator drain:
        //   iteration `it`:  [settle-issue(it-1): candidate list of the previous tile -> exact fp32 loads into registers]
        //                    drain(it) (TMEM -> ODF stores + key tile)  -> TMEM released
        //                    [settle-finish(it-1): compare, sorted-triple inserts]  [outputs(it-1)]
        //                    scan(it) -> candidate list of tile `it`
        // Candidate list, sorted triples, per-voxel minimum / mean are double-buffered (index it & 1); the key tile is not:
        // the only reader outside the scan, the tie test of settle-issue, runs before the drain overwrites it.
        uint32_t it = 0;
        bool have_prev = false;
        int64_t vox0_prev = 0;
        for (int i = et; i < 2 * 3 * VOX_CTA; i += EPI_THREADS) s_top[i] = 0ull;
        named_bar(1, EPI_THREADS);
        auto insert_top = [&](unsigned long long* top, int cv, int cx, float c) {
            // Insert (value, ~index) into the voxel's sorted triple: atomicMax returns what it displaced, and the
            // smaller of the two moves down one level.  Larger value wins, equal values -> smaller index wins
            // (the reference's stable order).  Every level ends up with the right key whatever the interleaving.
            unsigned long long key = ((unsigned long long)__float_as_uint(c) << 32) | (0xFFFFFFFFu - (uint32_t)cv);
#pragma unroll
            for (int lvl = 0; lvl < 3; ++lvl) {
                const unsigned long long old = atomicMax(&top[lvl * VOX_CTA + cx], key);
                key = old < key ? old : key;
                if (key == 0ull) break;
            }
        };
        for (int ti_ = cluster_id; ; ti_ += ncluster, ++it) {
            const bool have_cur = ti_ < ntl;
            if (!have_cur && !have_prev) break;
            const int buf = (int)(it & 1u), pbuf = buf ^ 1;
            // ---- settle-issue (previous tile): thread e takes listed pair e.  A key that decided strictly needs the
            //      value only; a key tie needs the neighbours' fp32 values as well.  All loads are L2 hits (this CTA
            //      wrote the values one tile ago) and stay in flight during the drain below. ----
            uint32_t ncand_raw = 0u; int ncand = 0;
            bool act = false, tie = false;
            int cv = 0, cx = 0;
            float cval = 0.f, nv[8];
            if (have_prev) {
                if (kTma && p.obuf > 0 && lane == 0) {                      // staged ODF boxes of the previous tile have left
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                ncand_raw = s_ncand[pbuf];
                ncand = (int)min(ncand_raw, (uint32_t)p.cand_cap);
                if (et < ncand) {
                    act = true;
                    const uint32_t ent = s_cand[pbuf * CAND_CAP + et];
                    cv = (int)(ent >> 8); cx = (int)(ent & 0xFFu);
                    const float* col = p.odf + vox0_prev + cx;
                    cval = __ldcg(col + (int64_t)cv * p.out_pitch);
                    const uint4 n0 = s_nbr[cv];
                    const uint32_t nn[4] = {n0.x, n0.y, n0.z, n0.w};
                    const uint32_t kc = keys[cv * VOX_CTA + cx];
                    uint32_t kn = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t n = (nn[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
                        kn = max(kn, (uint32_t)keys[(n != NBR_NONE ? n : (uint32_t)M) * VOX_CTA + cx]);
                    }
                    tie = kc <= kn;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t n = (nn[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
                        nv[k] = (tie && n != NBR_NONE) ? __ldcg(col + (int64_t)n * p.out_pitch) : -CUDART_INF_F;
                    }
                }
                named_bar(1, EPI_THREADS);                                  // every warp has read the key tile / list size
                if (et == 0) s_ncand[pbuf] = 0u;
            }
            int64_t vox0 = 0;
            if (have_cur) {
                const int tile = ident ? ti_ : __ldg(p.tile_list + ti_);
                vox0 = (int64_t)tile * 256 + rank * VOX_CTA;               // first voxel of this CTA's half
                const int64_t vox = vox0 + vl;
                const bool vok = vox < p.nvox;
                // DSI: p = Re(FFT)/sum(p) with sum(p) = dscale * s+[cvol], folded into the un-scale factor.
                // A voxel whose q = 0 sample is <= 0 gets zeros: either it is skipped by the reference as well (all
                // samples <= 0) or the reference divides by zero there (undefined; excluded from parity).
                float scl = inv_scale;
                if (p.cvol >= 0) {
                    const float sc = vok ? __ldg(p.dwi + (int64_t)p.cvol * p.dwi_pitch + vox) : 0.f;
                    scl = sc > 0.f ? inv_scale * (1.f / (p.dscale * sc)) : 0.f;     // inv_scale is a power of two: exact
                }
                mbar_wait<50>(d_full, it & 1);
                tc_fence_after();
                if (warp == W_EPI0) TRACE(2);
                // ---- phase 1: TMEM -> registers -> un-scale -> global ODF (coalesced) + 16-bit key tile ----
                // key(val) = ceil(32767 * clamp(val * ks, 0, 1)), ks = 1 / (KEY_WINDOW * mean): monotone in val, key >= 1
                // <=> val > 0.  Stored as 0x8000 | key = the low mantissa bits of fma.rp(sat(val * ks), 32767, 2^23 + 2^15)
                // (no conversion instruction); the always-set bit 15 lets one 32-bit subtraction compare two packed keys.
                float mn = CUDART_INF_F, meanv = 0.f;
                {
                    float ks = 0.f;
                    if (!p.plain) {
                        meanv = __uint_as_float(tmem_ld1(lane_addr + M)) * scl;   // extra matrix row M: mean of the ODF rows
                        tmem_wait_ld();
                        ks = (meanv > 0.f && meanv < CUDART_INF_F) ? (1.f / KEY_WINDOW) / meanv : 1e-30f;
                    }
                    float* gp = p.odf + (int64_t)c_begin * p.out_pitch + (vok ? vox : 0);
                    uint16_t* kp = keys + c_begin * VOX_CTA + vl;
                    const int64_t pitch = p.out_pitch;
                    const bool plain = p.plain != 0;
                    // key = 0x8000 | ceil(32767 * sat(val * ks)): two instructions (FMUL.SAT, FFMA.RP), the low 16 bits of
                    // the second result are the stored key
                    // Row addresses: one 64-bit pointer advanced by an opaque add per row (IADD3 + IADD3.X).  Left to itself the
                    // compiler rebuilds every row address from the chunk base with IMAD.WIDE chains (3-4 instructions per store).
                    const int64_t pitch_b = pitch * 4;
                    auto next_row = [&](float*& g) { asm volatile("add.s64 %0, %0, %1;" : "+l"(g) : "l"(pitch_b)); };
                    auto st_row = [](float* g, float v) { asm volatile("st.global.f32 [%0], %1;" ::"l"(g), "f"(v)); };   // (the opaque add hides the address space)
                    auto process = [&](const uint32_t (&r)[16], int c0) {
                        const int nrow = min(16, M - c0);                   // warp-uniform
                        float* g = gp;
                        if (plain) {
    #pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                if (j < nrow && vok) st_row(g, __uint_as_float(r[j]) * scl);
                                next_row(g);
                            }
                        } else if (nrow == 16) {
                            float prev = CUDART_INF_F;
    #pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float val = __uint_as_float(r[j]) * scl;
                                if (vok) st_row(g, val);
                                next_row(g);
                                const float y = __fmaf_ru(__saturatef(val * ks), 32767.f, 8421376.f);
                                kp[j * VOX_CTA] = (uint16_t)__float_as_uint(y);
                                if (j & 1) asm("min.f32 %0, %0, %1, %2;" : "+f"(mn) : "f"(prev), "f"(val));   // FMNMX3: one instruction per two values
                                else prev = val;
                            }
                        } else {
    #pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                if (j < nrow) {
                                    const float val = __uint_as_float(r[j]) * scl;
                                    if (vok) st_row(g, val);
                                    const float y = __fmaf_ru(__saturatef(val * ks), 32767.f, 8421376.f);
                                    kp[j * VOX_CTA] = (uint16_t)__float_as_uint(y);
                                    mn = fminf(mn, val);
                                }
                                next_row(g);
                            }
                        }
                        gp = g; kp += 16 * VOX_CTA;
                    };
                    // TMA mode: the 16 x 32 values of a chunk go to this warp's staging box and leave with ONE bulk tensor store
                    // (rows >= M and voxels >= nvox are clipped by the tensor map); the box is reused two chunks later, once the
                    // TMA unit has read it.  The drain then runs at TMEM / issue speed instead of the store port's.
                    const int vcoord = (int)vox0 + q * 32;
                    long long box_wait = 0;                                // (trace only)
                    auto process_tma = [&](const uint32_t (&r)[16], int c0) {
                        const int nrow = min(16, M - c0);                   // warp-uniform
                        const uint32_t bx = obox0 + (oc % obufm) * WBOX + lane * 4;
                        const long long tw0 = (kTrace && p.trace && warp == W_EPI0) ? clock64() : 0;
                        if (lane == 0) {
                            if (p.obuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        }
                        __syncwarp();
                        if (kTrace && p.trace && warp == W_EPI0) { box_wait += clock64() - tw0; TRACE(17 + ((c0 - c_begin) >> 4)); }
                        if (plain) {
    #pragma unroll
                            for (int j = 0; j < 16; ++j)
                                asm volatile("st.shared.f32 [%0], %1;" ::"r"(bx + j * 128), "f"(__uint_as_float(r[j]) * scl) : "memory");
                        } else if (nrow == 16) {
    #pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float val = __uint_as_float(r[j]) * scl;
                                asm volatile("st.shared.f32 [%0], %1;" ::"r"(bx + j * 128), "f"(val) : "memory");
                                const float y = __fmaf_ru(__saturatef(val * ks), 32767.f, 8421376.f);
                                kp[j * VOX_CTA] = (uint16_t)__float_as_uint(y);
                                mn = fminf(mn, val);
                            }
                        } else {
    #pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float val = __uint_as_float(r[j]) * scl;
                                asm volatile("st.shared.f32 [%0], %1;" ::"r"(bx + j * 128), "f"(val) : "memory");
                                if (j < nrow) {
                                    const float y = __fmaf_ru(__saturatef(val * ks), 32767.f, 8421376.f);
                                    kp[j * VOX_CTA] = (uint16_t)__float_as_uint(y);
                                    mn = fminf(mn, val);
                                }
                            }
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                                         ::"l"(&tmapO), "r"(vcoord), "r"(c0), "r"(bx) : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        kp += 16 * VOX_CTA; ++oc;
                    };
                    const bool staged = kTma && p.obuf > 0;
                    auto step = [&](const uint32_t (&r)[16], int c0) { if (p.abl & 4) { mn = fminf(mn, __uint_as_float(r[0])); return; } if (staged) process_tma(r, c0); else process(r, c0); };
                    uint32_t ra[16], rb[16];                               // double buffer: the next chunk's load is in flight
                    if (c_begin < c_end) tmem_ld16(lane_addr + c_begin, ra);
                    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                        tmem_wait_ld();
                        if (c0 + 16 < c_end) tmem_ld16(lane_addr + c0 + 16, rb);
                        step(ra, c0);
                        if (c0 + 16 < c_end) {
                            tmem_wait_ld();
                            if (c0 + 32 < c_end) tmem_ld16(lane_addr + c0 + 32, ra);
                            step(rb, c0 + 16);
                        }
                    }
                    if (kTma && kTrace && p.trace && warp == W_EPI0) TRACE_ADD(24, box_wait);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(dempty0);               // TMEM may be overwritten by the next tile
                if (warp == W_EPI0) TRACE(3);
                if (!p.plain) {
                    s_min[(buf * N_CPART + cpart) * VOX_CTA + vl] = mn;
                    if (cpart == 0) s_mean[buf * VOX_CTA + vl] = meanv;
                }
            }
            if (p.plain) continue;                                          // rows only (DSI pdf): nothing is staged
            // ---- settle-finish (previous tile) ----
            if (have_prev) {
                unsigned long long* top = s_top + pbuf * 3 * VOX_CTA;
                if (act) {
                    bool ok = cval > 0.f;
                    if (tie) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) ok = ok && (cval > nv[k]);
                    }
                    if (ok) insert_top(top, cv, cx, cval);
                }
                // more listed pairs than threads (rare): the rest is settled on fp32 values alone (the key tile is gone)
                for (int e = et + EPI_THREADS; e < ncand; e += EPI_THREADS) {
                    const uint32_t ent = s_cand[pbuf * CAND_CAP + e];
                    const int ev = (int)(ent >> 8), ex = (int)(ent & 0xFFu);
                    const float* col = p.odf + vox0_prev + ex;
                    const float c = __ldcg(col + (int64_t)ev * p.out_pitch);
                    const uint4 n0 = s_nbr[ev];
                    const uint32_t nn[4] = {n0.x, n0.y, n0.z, n0.w};
                    float w[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t n = (nn[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
                        w[k] = (n != NBR_NONE) ? __ldcg(col + (int64_t)n * p.out_pitch) : -CUDART_INF_F;
                    }
                    bool ok = c > 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) ok = ok && (c > w[k]);
                    if (ok) insert_top(top, ev, ex, c);
                }
            }
            named_bar(1, EPI_THREADS);          // sorted triples of the previous tile complete; key tile + ODF stores of this tile visible to the group
            if (warp == W_EPI0) TRACE(4);
            // ---- outputs (previous tile): warp group k (4 warps = 128 voxels) writes peak k; group 0 also owns the statistics ----
            if (have_prev) {
                unsigned long long* top = s_top + pbuf * 3 * VOX_CTA;
                const int k = ew >> 2;                                  // 0..2 (N_EPI == 12)
                const int ov = (ew & 3) * 32 + lane;                    // voxel within the CTA
                const int64_t ovox = vox0_prev + ov;
                const bool ook = ovox < p.nvox;
                float omin = s_min[pbuf * N_CPART * VOX_CTA + ov];
#pragma unroll
                for (int cp = 1; cp < N_CPART; ++cp) omin = fminf(omin, s_min[(pbuf * N_CPART + cp) * VOX_CTA + ov]);
                {
                    const unsigned long long key = top[k * VOX_CTA + ov];
                    top[k * VOX_CTA + ov] = 0ull;                       // (this thread is the entry's only reader) ready for the tile after next
                    if (ook) {
                        const bool ok = key != 0ull;
                        const int id = ok ? (int)(0xFFFFFFFFu - (uint32_t)key) : 0;
                        const float val = __uint_as_float((uint32_t)(key >> 32));
                        p.peak[k][ovox]                   = ok ? s_vert[id * 3 + 0] : 0.f;
                        p.peak[k][ovox + p.out_pitch]     = ok ? s_vert[id * 3 + 1] : 0.f;
                        p.peak[k][ovox + 2 * p.out_pitch] = ok ? s_vert[id * 3 + 2] : 0.f;
                        p.qa[k][ovox] = ok ? val - omin : 0.f;
                        if (p.peak_idx) p.peak_idx[ovox + k * p.out_pitch] = ok ? (int16_t)id : (int16_t)-1;
                    }
                }
                if (k == 0) {
                    float mean = s_mean[pbuf * VOX_CTA + ov];
                    // fp16 overflow of the scaled signal (every accumulator column of the voxel, the mean row included, is
                    // then non-finite), or more listed pairs than the list holds
                    const bool bad = ook && (!(fabsf(mean) < CUDART_INF_F) || ncand_raw > (uint32_t)p.cand_cap);
                    if (!ook || bad) mean = -CUDART_INF_F;
                    const unsigned anybad = __ballot_sync(0xffffffffu, bad);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) mean = fmaxf(mean, __shfl_xor_sync(0xffffffffu, mean, o));
                    if (lane == 0) {
                        if (mean > -CUDART_INF_F) atomicMax(p.stats, f2ord(mean));
                        if (anybad) {                                   // recompute this 64-voxel tile with the SIMT kernel
                            const int slot = atomicAdd(p.fix_count, 1);
                            if (slot < p.fix_cap) p.fix_list[slot] = (int)((vox0_prev + (ew & 3) * 32) >> 6);
                        }
                    }
                }
                if (warp == W_EPI0) TRACE(7);
            }
            if (have_cur) {
                // ---- phase 2: scan.  A vertex can only be a local maximum (value > 0, > every mesh neighbour) if its
                //      key is >= max(neighbour keys, 1); equality means "cannot tell from the keys".  Both kinds are
                //      listed.  Two vertices per iteration; software-pipelined: the keys of pair i+1 and the offsets of
                //      pair i+2 are in flight while pair i is tested.  (Prefetches past the warp's range touch rows
                //      < M + 8 of the offset table / key tile, which exist; their values are never tested.) ----
                if (!(p.abl & 2)) {
                    uint32_t kb = smem_u32(keys) + lane * 8;
                    asm volatile("mov.u32 %0, %0;" : "+r"(kb));             // pin: keeps the compiler from re-deriving the base in every iteration
                    const uint32_t ncand32 = smem_u32(s_ncand + buf);
                    auto ld = [&](uint32_t off) {
                        uint2 r; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(kb + off)); return r;
                    };
                    // max of the six neighbour keys and of the threshold key 1 (two packed voxels): three 3-input packed max
                    auto max6 = [](uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f) {
                        return umax2(umax2(umax2(umax2(a, b), c), umax2(umax2(d, e), f)), 0x80018001u);
                    };
                    const bool wide = p.nbw > 6;                            // warp-uniform (meshes with degree 7-8)
                    // Vertex pairs are dealt round-robin to the warps (pair ew, ew + N_EPI, ...): ODF peaks are spatially
                    // compact, so contiguous ranges would give all the listing work of a tile to one or two warps.
                    const int npair = (M + 1) >> 1;
                    auto pair_vertex = [&](int pi) { return 2 * min(pi, npair); };      // past the end -> sentinel rows
                    int pi = ew;
                    int v = pair_vertex(pi), vn = pair_vertex(pi + N_EPI);
                    const uint32_t* op = c_nbr_off + v * NBR_W;
                    uint4 oA0 = *reinterpret_cast<const uint4*>(op), oB0 = *reinterpret_cast<const uint4*>(op + 8);
                    uint2 oA1 = *reinterpret_cast<const uint2*>(op + 4), oB1 = *reinterpret_cast<const uint2*>(op + 12);
                    uint2 cA = ld(v * KEY_ROW), cB = ld((v + 1) * KEY_ROW);
                    uint2 a0 = ld(oA0.x), a1 = ld(oA0.y), a2 = ld(oA0.z), a3 = ld(oA0.w), a4 = ld(oA1.x), a5 = ld(oA1.y);
                    uint2 b0 = ld(oB0.x), b1 = ld(oB0.y), b2 = ld(oB0.z), b3 = ld(oB0.w), b4 = ld(oB1.x), b5 = ld(oB1.y);
                    op = c_nbr_off + vn * NBR_W;
                    oA0 = *reinterpret_cast<const uint4*>(op); oB0 = *reinterpret_cast<const uint4*>(op + 8);
                    oA1 = *reinterpret_cast<const uint2*>(op + 4); oB1 = *reinterpret_cast<const uint2*>(op + 12);
                    uint32_t fw0 = 0u, fw1 = 0u, fw2 = 0u, fw3 = 0u;       // result bits: one byte per iteration (<= 16 iterations: M <= 383)
                    int iter = 0;
                    for (; pi < npair; pi += N_EPI) {
                        // reduce pair i
                        uint32_t mAx = max6(a0.x, a1.x, a2.x, a3.x, a4.x, a5.x), mAy = max6(a0.y, a1.y, a2.y, a3.y, a4.y, a5.y);
                        uint32_t mBx = max6(b0.x, b1.x, b2.x, b3.x, b4.x, b5.x), mBy = max6(b0.y, b1.y, b2.y, b3.y, b4.y, b5.y);
                        if (wide) {
                            const uint2 wA = *reinterpret_cast<const uint2*>(c_nbr_off + v * NBR_W + 6), wB = *reinterpret_cast<const uint2*>(c_nbr_off + v * NBR_W + 14);
                            const uint2 a6 = ld(wA.x), a7 = ld(wA.y), b6 = ld(wB.x), b7 = ld(wB.y);
                            mAx = umax2(mAx, umax2(a6.x, a7.x)); mAy = umax2(mAy, umax2(a6.y, a7.y));
                            mBx = umax2(mBx, umax2(b6.x, b7.x)); mBy = umax2(mBy, umax2(b6.y, b7.y));
                        }
                        const uint2 kA = cA, kB = cB;
                        // issue pair i+1 (offsets arrived during the previous iteration) and fetch the offsets of pair i+2
                        v = vn; vn = pair_vertex(pi + 2 * N_EPI);
                        cA = ld(v * KEY_ROW); cB = ld((v + 1) * KEY_ROW);
                        a0 = ld(oA0.x); a1 = ld(oA0.y); a2 = ld(oA0.z); a3 = ld(oA0.w); a4 = ld(oA1.x); a5 = ld(oA1.y);
                        b0 = ld(oB0.x); b1 = ld(oB0.y); b2 = ld(oB0.z); b3 = ld(oB0.w); b4 = ld(oB1.x); b5 = ld(oB1.y);
                        op = c_nbr_off + vn * NBR_W;
                        oA0 = *reinterpret_cast<const uint4*>(op); oB0 = *reinterpret_cast<const uint4*>(op + 8);
                        oA1 = *reinterpret_cast<const uint2*>(op + 4); oB1 = *reinterpret_cast<const uint2*>(op + 12);
                        // stored keys have bit 15 set, so per 16-bit half  (k - t + 0x8000) has bit 15 set  <=>  k >= t,
                        // and the two halves cannot borrow from each other: one 32-bit subtraction tests two voxels.
                        // threshold t = max(neighbour keys, 1)
                        const uint32_t hAx = kA.x - mAx + 0x80008000u, hAy = kA.y - mAy + 0x80008000u;
                        const uint32_t hBx = kB.x - mBx + 0x80008000u, hBy = kB.y - mBy + 0x80008000u;
                        // No branch in the loop: the result bits of the pair are bit 15 / 31 of the four words.  Two byte
                        // permutes collect the bytes that hold them (byte b of pA = vertex v, voxel 4*lane + b; same for pB and
                        // vertex v + 1), one mask + shift + or interleaves them (A at bit 7, B at bit 6 of every byte) and the
                        // pair of iteration j lands 2 j bits lower: four iterations fill a 32-bit word.
                        uint32_t pA, pB;
                        asm("prmt.b32 %0, %1, %2, 0x7531;" : "=r"(pA) : "r"(hAx), "r"(hAy));
                        asm("prmt.b32 %0, %1, %2, 0x7531;" : "=r"(pB) : "r"(hBx), "r"(hBy));
                        const uint32_t fl = (pA & 0x80808080u) | ((pB >> 1) & 0x40404040u);
                        const uint32_t sh = (uint32_t)(iter & 3) * 2u;
                        switch (iter >> 2) {                                // (warp-uniform)
                            case 0: fw0 |= fl >> sh; break;
                            case 1: fw1 |= fl >> sh; break;
                            case 2: fw2 |= fl >> sh; break;
                            default: fw3 |= fl >> sh; break;
                        }
                        ++iter;
                    }
                    // List the hits (typically a dozen lanes per warp and tile have one or two).  The slots come from ONE shared-memory
                    // atomic per warp (warp prefix sum of the per-lane counts): per-lane atomics on the single counter serialise.
                    const uint32_t nhit = (p.abl & 32) ? 0u : __popc(fw0) + __popc(fw1) + __popc(fw2) + __popc(fw3);   // (ablation 32: nothing is listed)
                    if (__any_sync(0xffffffffu, nhit != 0u)) {
                        uint32_t incl = nhit;
    #pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                        uint32_t base = 0u;
                        if (lane == 31) asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(base) : "r"(ncand32), "r"(incl) : "memory");
                        base = __shfl_sync(0xffffffffu, base, 31);
                        uint32_t slot = base + incl - nhit;
                        const uint32_t fws[4] = {fw0, fw1, fw2, fw3};
    #pragma unroll
                        for (int wi = 0; wi < 4; ++wi) {
                            uint32_t bits = fws[wi];
                            while (bits) {
                                const int h = __ffs(bits) - 1;
                                bits &= bits - 1;
                                const int r = 7 - (h & 7);                  // 2 * (iteration within the word) + (vertex v + 1 ?)
                                const int vv = pair_vertex(ew + (wi * 4 + (r >> 1)) * N_EPI) + (r & 1);
                                if (slot < (uint32_t)p.cand_cap) s_cand[buf * CAND_CAP + slot] = ((uint32_t)vv << 8) | (uint32_t)(4 * lane + (h >> 3));
                                ++slot;
                            }
                        }
                    }
                }
                named_bar(1, EPI_THREADS);                                  // candidate list of this tile complete
                if (warp == W_EPI0) TRACE(5);
            }
            have_prev = have_cur;
            vox0_prev = vox0;
        }
    }

    // ---- teardown --------------------------------------------------------------------------
    if (kTma && warp >= W_EPI0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // staging boxes still being read
    if (kTrace && p.trace && warp == W_EPI0 && lane == 0 && rank == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[16 * 32 + 2 * cluster_id + 1] = (long long)t;
        if (cluster_id == 0) p.trace[16 * 32 + 2 * 127 + 1] = clock64();
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == W_MMA)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// dynamic shared memory of one instantiation (carve-up of recon_tc_kernel; M = 0 for plain passes)
size_t tc_smem_bytes(int tma, int M, int Nh, int nstage, int tstage, int obuf) {     // tma: 0 = cp.async staging, else the number of DWI maps
    size_t b = (size_t)nstage * 2 * Nh * 32 + (size_t)(M + 8) * KEY_ROW + 2 * (3 * VOX_CTA * 8 + (size_t)CAND_CAP * 4 + (N_CPART + 1) * VOX_CTA * 4);
    b = (b + 127) & ~(size_t)127;
    b += tma ? (size_t)tstage * dwi_stage_bytes(tma) + (size_t)N_CPART * obuf * OBOX_BYTES : (size_t)DSTAGE * 32 * VOX_CTA * 4;
    b += (size_t)M * 16 + (size_t)((3 * M + 3) & ~3) * 4 + (2 * NSTAGE + 2 * ASLOT + 2 + 2 * TSTAGE) * 8 + 16;
    return b + 1024 + 64;
}

// The dynamic shared-memory limit is an attribute of (function, device), shared by all plans: it is only ever
// raised (a plan with a smaller tile must not lower it under a live plan with a larger one).
std::mutex g_smem_mu;
size_t g_smem_limit[2][64] = {};
int raise_smem_limit(int device, bool tma, size_t smem);

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// Decide whether the tensor-core kernel can take this plan; build the split fp16 operands and their
// tensor maps.  Returns 0 when usable.
static void split_dims(int rows, int& Npad, int& N1, int& N2) {
    Npad = (rows + 15) / 16 * 16; N1 = Npad; N2 = 0;
    if (Npad > 256) { N1 = (Npad / 2 + 15) / 16 * 16; N2 = Npad - N1; }
}

static void tc_state_free(TcState* st) {
    if (!st) return;
    for (auto& ps : st->pass) cudaFree(ps.d_split);
    cudaFree(st->d_scratch);
    delete st;
}

int tc_plan_init(Plan* p) {
    if (p->kind != PLAN_GQI && p->kind != PLAN_DSI) { set_error("tensor-core path: GQI / DSI plans only"); return 1; }
    const int M = p->nvert, K = p->nvol;
    if ((M + 1 + 15) / 16 * 16 > 384 || M + 8 > TC_MAX_VERT) { set_error("tensor-core path: more than 383 half-sphere vertices"); return 1; }
    if (p->kind == PLAN_DSI && p->cvol < 0) { set_error("tensor-core path: DSI without a q-space origin sample"); return 1; }
    int cc_major = 0;
    cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, p->device);
    if (cc_major != 10) { set_error("tensor-core path needs sm_100"); return 1; }
    void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
        set_error("tensor-core path: cuTensorMapEncodeTiled unavailable"); cudaGetLastError(); return 1;
    }
    const int Kpad = (K + 31) / 32 * 32;
    TcState* st = new TcState();
    st->Kpad = Kpad; st->nbw = p->nbr_width; st->encode = fn;
    for (int i = 0; i < TC_MAX_VERT * NBR_W; ++i) st->nbr_off.off[i] = (uint32_t)M * KEY_ROW;      // sentinel row M everywhere ...
    for (int v = 0; v < M; ++v)
        for (int k = 0; k < NBR_W; ++k) {
            const uint16_t n = p->h_nbr[(size_t)v * NBR_W + k];
            if (n != NBR_NONE) st->nbr_off.off[(size_t)v * NBR_W + k] = (uint32_t)n * KEY_ROW;
        }
    // passes: ODF rows, then (DSI) the pdf rows in blocks of <= 336
    std::vector<std::pair<int, int>> ranges = {{0, M}};
    if (p->kind == PLAN_DSI) for (int r0 = 0; r0 < K; r0 += 336) ranges.push_back({M + r0, std::min(336, K - r0)});
    size_t smem = 0;
    int dev_smem = 0;
    if (cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, p->device) != cudaSuccess) dev_smem = 0;
    for (size_t i = 0; i < ranges.size(); ++i) {
        TcPass ps;
        ps.row0 = ranges[i].first; ps.rows = ranges[i].second; ps.plain = i > 0;
        const int img_rows_n = ps.rows + (ps.plain ? 0 : 1);          // ODF pass: one extra row = mean of the ODF rows
        split_dims(img_rows_n, ps.Npad, ps.N1, ps.N2);
        const int Nh = (ps.N1 + ps.N2) / 2, N1h = ps.N1 / 2, N2h = ps.N2 / 2;
        // smallest configuration of either instantiation must fit (the launch picks the ring depths)
        const int Mk = ps.plain ? 0 : ps.rows;
        smem = std::max(smem, std::max(tc_smem_bytes(4, Mk, Nh, 2, 4, 0), tc_smem_bytes(0, Mk, Nh, 2, 0, 0)));
        // Split operand as a ready-made shared-memory image: [rank][K32 chunk][K16 sub-tile][hi | lo][row][16 halves],
        // rows in the order (blk1 rows 0..N1h, blk2 rows 0..N2h) of that rank, with the SWIZZLE_32B pattern the
        // tcgen05 descriptors expect already applied (16-byte chunk ^= bit 2 of the row).  A stage is then ONE
        // contiguous 128*Nh-byte block: the producer moves it with a single TMA copy of 512-byte rows instead of
        // 4*Nh separate 32-byte rows.
        const int nk32 = Kpad / 32;
        const size_t stage_halves = (size_t)4 * Nh * 16;
        std::vector<__half> split((size_t)2 * nk32 * stage_halves, __float2half(0.f));
        std::vector<float> mean_row;
        if (!ps.plain) {                                   // column means of the ODF rows (fp64 accumulation)
            mean_row.resize(K);
            for (int k = 0; k < K; ++k) {
                double acc = 0.0;
                for (int n = 0; n < ps.rows; ++n) acc += p->h_matrix[(size_t)(ps.row0 + n) * K + k];
                mean_row[k] = (float)(acc / ps.rows);
            }
        }
        for (int n = 0; n < img_rows_n; ++n) {
            int rank, local;
            if (n < ps.N1) { rank = n / N1h; local = n % N1h; }
            else { const int m = n - ps.N1; rank = m / N2h; local = N1h + m % N2h; }
            const float* src = n < ps.rows ? p->h_matrix.data() + (size_t)(ps.row0 + n) * K : mean_row.data();
            for (int k = 0; k < K; ++k) {
                const __half h = __float2half_rn(src[k]);
                const __half l = __float2half_rn(src[k] - __half2float(h));
                const int c = k / 32, sub = (k % 32) / 16, j = k % 16;
                const size_t stage = ((size_t)rank * nk32 + c) * stage_halves;
                const size_t inrow = (size_t)(((j / 8) ^ ((local >> 2) & 1)) * 8 + j % 8);
                split[stage + ((size_t)(sub * 2 + 0) * Nh + local) * 16 + inrow] = h;
                split[stage + ((size_t)(sub * 2 + 1) * Nh + local) * 16 + inrow] = l;
            }
        }
        if (cudaMalloc(&ps.d_split, split.size() * sizeof(__half)) != cudaSuccess ||
            cudaMemcpy(ps.d_split, split.data(), split.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("tensor-core path: device allocation failed"); cudaGetLastError(); tc_state_free(st); return 1;
        }
        st->pass.push_back(ps);
        const cuuint64_t img_rows = (cuuint64_t)2 * nk32 * (stage_halves / 256);
        cuuint64_t gdim[2] = {256, img_rows};
        cuuint64_t gstr[1] = {512};
        cuuint32_t box[2] = {256, (cuuint32_t)(stage_halves / 512)};          // one K16 sub-tile (hi rows + lo rows) per stage
        cuuint32_t estr[2] = {1, 1};
        if (((EncodeFn)fn)(&st->pass.back().tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ps.d_split, gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            set_error("tensor-core path: cuTensorMapEncodeTiled failed"); tc_state_free(st); return 1;
        }
    }
    if (smem > (size_t)dev_smem) { set_error("tensor-core path: tile does not fit in shared memory"); tc_state_free(st); return 1; }
    st->dev_smem = dev_smem;
    if (raise_smem_limit(p->device, true, (size_t)dev_smem) || raise_smem_limit(p->device, false, (size_t)dev_smem)) { tc_state_free(st); return 1; }
    p->tc = st;
    return 0;
}

namespace {
int raise_smem_limit(int device, bool tma, size_t smem) {
    std::lock_guard<std::mutex> lk(g_smem_mu);
    size_t& cur = g_smem_limit[tma ? 1 : 0][device & 63];
    if (smem <= cur) return 0;
    const int sm = (int)smem;
    cudaError_t e = cudaSuccess;
    auto raise = [&](const void* fn) { if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, sm); };
    if (tma) {
        raise((const void*)recon_tc_kernel<1, false>); raise((const void*)recon_tc_kernel<1, true>);
        raise((const void*)recon_tc_kernel<2, false>); raise((const void*)recon_tc_kernel<4, false>);
    } else {
        raise((const void*)recon_tc_kernel<0, false>); raise((const void*)recon_tc_kernel<0, true>);
    }
    if (e != cudaSuccess) { set_error("tensor-core path: cannot raise the shared-memory limit"); cudaGetLastError(); return 1; }
    cur = smem;
    return 0;
}
}  // namespace

void tc_plan_free(Plan* p) {
    tc_state_free(reinterpret_cast<TcState*>(p->tc));
    p->tc = nullptr;
}

int launch_recon_tc(Plan* p, const ReconArgs& a, cudaStream_t stream) {
    TcState* st = reinterpret_cast<TcState*>(p->tc);
    if (!st) return fail(FIBERS_ERR_ARG, "plan has no tensor-core state");
    if (a.nvox <= 0) return 0;
    if (a.nvox > 0x7FFFFFFFLL * 32 || a.out_pitch >= (1ll << 30)) return fail(FIBERS_ERR_ARG, "slab too large");
    const int64_t ntile64 = (a.nvox + 63) / 64, ntile256 = (a.nvox + 255) / 256;
    const int64_t need = 4 + 2 * ntile64 + 2 * ntile256 + 8;   // [0] maxbits [1] fix_count [2] tile_count | fix list | tile list | tile flags
    if (st->scratch_cap < need) {
        if (st->d_scratch) cudaFree(st->d_scratch);
        st->d_scratch = nullptr; st->scratch_cap = 0;
        FB_CUDA(cudaMalloc(&st->d_scratch, sizeof(int) * (size_t)need));
        st->scratch_cap = need;
    }
    if (int rc = plan_enter(p, stream)) return rc;                  // the scratch below belongs to the plan
    FB_CUDA(cudaMemsetAsync(st->d_scratch, 0, 4 * sizeof(int), stream));
    int* d_fix_list = st->d_scratch + 4; int* d_tile_list = st->d_scratch + 4 + 2 * ntile64;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, p->device);
    sample_max_kernel<<<nsm * 8, 256, 0, stream>>>(a.dwi, a.dwi_pitch, a.nvox, p->nvol, st->d_scratch);
    {
        ScanOut so{}; int n = 0;
        so.ptr[n] = a.odf; so.rows[n++] = p->nvert;
        if (p->kind == PLAN_DSI && a.pdf) { so.ptr[n] = a.pdf; so.rows[n++] = p->nvol; }
        for (int k = 0; k < 3; ++k) { so.ptr[n] = a.peak[k]; so.rows[n++] = 3; }
        for (int k = 0; k < 3; ++k) { so.ptr[n] = a.qa[k]; so.rows[n++] = 1; }
        so.n = n; so.idx = a.peak_idx;
        int* d_tile_flag = d_tile_list + ntile256;
        tile_scan_kernel<<<(unsigned)((ntile256 + 7) / 8), 256, 0, stream>>>(a.mask, a.nvox, a.out_pitch, so, a.stats, d_tile_flag, (int)ntile256, st->d_scratch + 2);
        tile_compact_kernel<<<1, 1024, 0, stream>>>(d_tile_flag, (int)ntile256, d_tile_list, st->d_scratch + 2);
    }
    count_launch(3);
    const char* trace_path = getenv("FIBERS_TC_TRACE");
    for (size_t ip = 0; ip < st->pass.size(); ++ip) {
        const TcPass& ps = st->pass[ip];
        float* out = ps.plain ? a.pdf + (int64_t)(ps.row0 - p->nvert) * a.out_pitch : a.odf;
        if (ps.plain && !a.pdf) continue;
        TcParams tp{};
        tp.dwi = a.dwi; tp.dwi_pitch = a.dwi_pitch; tp.mask = a.mask; tp.nvox = a.nvox;
        tp.dwi_vec = ((uintptr_t)a.dwi % 16 == 0 && a.dwi_pitch % 4 == 0) ? 2 : ((uintptr_t)a.dwi % 8 == 0 && a.dwi_pitch % 2 == 0) ? 1 : 0;
        tp.cand_cap = CAND_CAP;
        tp.conv_sleep = 100; tp.prod_sleep = 200;
        if (const char* e = getenv("FIBERS_TC_CONV_SLEEP")) tp.conv_sleep = (uint32_t)atoi(e);
        if (const char* e = getenv("FIBERS_TC_PROD_SLEEP")) tp.prod_sleep = (uint32_t)atoi(e);
        if (const char* cap = getenv("FIBERS_TC_CAND_CAP")) tp.cand_cap = std::max(0, std::min(CAND_CAP, atoi(cap)));
        tp.K = p->nvol; tp.Kpad = st->Kpad; tp.M = ps.rows; tp.Npad = ps.Npad; tp.N1 = ps.N1; tp.N2 = ps.N2;
        tp.odf = out; tp.out_pitch = a.out_pitch;
        for (int k = 0; k < 3; ++k) { tp.peak[k] = a.peak[k]; tp.qa[k] = a.qa[k]; }
        tp.peak_idx = a.peak_idx; tp.stats = a.stats; tp.nbr = p->d_nbr; tp.vert = p->d_vert;
        tp.maxbits = st->d_scratch; tp.fix_count = st->d_scratch + 1; tp.fix_list = d_fix_list;
        tp.tile_list = d_tile_list; tp.tile_count = st->d_scratch + 2;
        tp.fix_cap = (int)(2 * ntile64);
        tp.ntiles = (int)((a.nvox + 255) / 256);
        tp.nbw = st->nbw;
        // TMA instantiation: the DWI slab through 1, 2 or 4 maps (16-, 8-, 4-byte aligned rows: DwiMaps); the output map (staged
        // ODF stores, FIBERS_TC_OBUF) needs 16-byte aligned output rows, the default direct stores do not
        int nmap = ((uintptr_t)a.dwi % 16 == 0 && a.dwi_pitch % 4 == 0) ? 1 : a.dwi_pitch % 2 == 0 ? 2 : 4;
        if (const char* e = getenv("FIBERS_TC_DBG_NMAP")) nmap = std::max(nmap, atoi(e));      // (experiments: more maps than the pitch needs)
        const bool out_al = (uintptr_t)out % 16 == 0 && a.out_pitch % 4 == 0;
        const bool tma = getenv("FIBERS_TC_NO_TMA") == nullptr && (uintptr_t)a.dwi % 4 == 0 && p->nvol >= nmap && a.nvox < (1ll << 31) - 512 &&
                         (nmap == 1 || getenv("FIBERS_TC_NO_SPLIT_TMA") == nullptr);
        // ring depths: the deepest B ring (L2 latency), a DWI ring that covers L2 latency (HBM latency is covered by the L2
        // prefetch one tile ahead) and one ODF staging box per epilogue warp, shrunk until the tile fits in shared memory
        const int Mk = ps.plain ? 0 : ps.rows, Nhh = (ps.N1 + ps.N2) / 2;
        auto envi = [](const char* n, int d) { const char* e = getenv(n); return e ? atoi(e) : d; };
        tp.nstage = std::max(2, std::min(NSTAGE, envi("FIBERS_TC_BSTAGES", 8)));
        // (plain passes -- DSI pdf rows -- keep no key tile in shared memory and are paced by the DWI feed: 8 stages there, - 4.4 % on cfg3)
        tp.tstage = std::max(2, std::min(TSTAGE, envi("FIBERS_TC_DSTAGES", ps.plain ? 8 : 4))) & ~1;
        tp.obuf = std::max(0, std::min(2, envi("FIBERS_TC_OBUF", 0)));
        tp.l2pf = std::max(0, envi("FIBERS_TC_L2PF", 0));
        tp.abl = envi("FIBERS_TC_ABLATE", 0);
        if (!tma) { tp.tstage = 0; tp.obuf = 0; tp.l2pf = 0; }
        if (!out_al) tp.obuf = 0;
        while (tc_smem_bytes(tma ? nmap : 0, Mk, Nhh, tp.nstage, tp.tstage, tp.obuf) > (size_t)st->dev_smem) {
            if (tp.nstage > 4) --tp.nstage; else if (tp.obuf > 0) --tp.obuf; else if (tp.tstage > 4) tp.tstage -= 2; else if (tp.nstage > 2) --tp.nstage;
            else return fail(FIBERS_ERR_ARG, "tensor-core path: tile does not fit in shared memory");
        }
        const size_t smem_launch = tc_smem_bytes(tma ? nmap : 0, Mk, Nhh, tp.nstage, tp.tstage, tp.obuf);
        if (getenv("FIBERS_TC_VERBOSE")) {
            static std::atomic<int> once{0};
            if (once.fetch_add(1) < 4) fprintf(stderr, "[fibers tc] pass %zu: tma %d, B stages %d, DWI stages %d, ODF boxes %d, L2 prefetch %d, smem %zu\n",
                                               ip, (int)tma, tp.nstage, tp.tstage, tp.obuf, tp.l2pf, smem_launch);
        }
        tp.plain = ps.plain;
        tp.cvol = p->kind == PLAN_DSI ? p->cvol : -1; tp.dscale = p->dscale;
        const int nclusters = std::max(1, std::min(nsm / 2, tp.ntiles));
        long long* d_trace = nullptr;
        if (trace_path && *trace_path && ip == 0) {
            FB_CUDA(cudaMalloc(&d_trace, (16 * 32 + 2 * 128) * sizeof(long long)));
            FB_CUDA(cudaMemsetAsync(d_trace, 0, (16 * 32 + 2 * 128) * sizeof(long long), stream));
            tp.trace = d_trace;
            if (const char* sk = getenv("FIBERS_TC_TRACE_SKIP")) tp.trace_skip = (uint32_t)atoi(sk);
        }
        if (tma) {
            // per-launch tensor maps: the DWI slab [K][nvox] fp32 (row pitch dwi_pitch) in K16 x 128-voxel boxes, the output
            // rows [rows][nvox] fp32 (row pitch out_pitch) in 16 x 32 boxes; out-of-range voxels / rows are zero-filled / clipped
            DwiMaps mD; CUtensorMap mO;
            cuuint32_t estr[2] = {1, 1};
            cuuint64_t sD[1] = {(cuuint64_t)a.dwi_pitch * 4 * nmap};
            cuuint32_t bD[2] = {(cuuint32_t)(nmap > 1 ? DBOXW_SPLIT : VOX_CTA), (cuuint32_t)(16 / nmap)};
            cuuint64_t dO[2] = {(cuuint64_t)a.nvox, (cuuint64_t)ps.rows}, sO[1] = {(cuuint64_t)a.out_pitch * 4};
            cuuint32_t bO[2] = {32, 16};
            for (int r = 0; r < 4; ++r) {
                const int rr = r < nmap ? r : 0;                            // (unused slots repeat map 0)
                const uintptr_t row = (uintptr_t)(a.dwi + (int64_t)rr * a.dwi_pitch);
                const int shift = (int)(row % 16) / 4;
                tp.dshift[r] = shift;
                cuuint64_t dD[2] = {(cuuint64_t)a.nvox + shift, (cuuint64_t)((p->nvol - rr + nmap - 1) / nmap)};
                if (((EncodeFn)st->encode)(&mD.m[r], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)(row - shift * 4), dD, sD, bD, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                    return fail(FIBERS_ERR_CUDA, "tensor-core path: cuTensorMapEncodeTiled failed for the slab map");
            }
            if (!out_al) mO = mD.m[0];                                      // never dereferenced: no staged stores
            else if (((EncodeFn)st->encode)(&mO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)out, dO, sO, bO, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return fail(FIBERS_ERR_CUDA, "tensor-core path: cuTensorMapEncodeTiled failed for the output map");
            if (tp.trace && nmap == 1) recon_tc_kernel<1, true><<<2 * nclusters, TC_THREADS, smem_launch, stream>>>(tp, ps.tmap, st->nbr_off, mD, mO);
            else if (nmap == 1) recon_tc_kernel<1, false><<<2 * nclusters, TC_THREADS, smem_launch, stream>>>(tp, ps.tmap, st->nbr_off, mD, mO);
            else if (nmap == 2) recon_tc_kernel<2, false><<<2 * nclusters, TC_THREADS, smem_launch, stream>>>(tp, ps.tmap, st->nbr_off, mD, mO);
            else recon_tc_kernel<4, false><<<2 * nclusters, TC_THREADS, smem_launch, stream>>>(tp, ps.tmap, st->nbr_off, mD, mO);
        } else {
            DwiMaps mD;
            for (int r = 0; r < 4; ++r) mD.m[r] = ps.tmap;                  // never dereferenced
            if (tp.trace) recon_tc_kernel<0, true><<<2 * nclusters, TC_THREADS, smem_launch, stream>>>(tp, ps.tmap, st->nbr_off, mD, ps.tmap);
            else recon_tc_kernel<0, false><<<2 * nclusters, TC_THREADS, smem_launch, stream>>>(tp, ps.tmap, st->nbr_off, mD, ps.tmap);
        }
        count_launch(1);
        FB_CUDA(cudaGetLastError());
        if (d_trace) {
            std::vector<long long> h(16 * 32 + 2 * 128);               // tile trace + per-cluster (start, end) in ns
            FB_CUDA(cudaStreamSynchronize(stream));
            FB_CUDA(cudaMemcpy(h.data(), d_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            cudaFree(d_trace);
            if (FILE* f = fopen(trace_path, "wb")) { fwrite(h.data(), sizeof(long long), h.size(), f); fclose(f); }
        }
    }
    // voxels whose scaled signal overflowed fp16 (rare): recompute their 64-voxel tiles in fp32
    if (int rc = launch_recon_simt_list(p, a, d_fix_list, st->d_scratch + 1, stream)) return rc;
    return plan_leave(p, stream);
}

}  // namespace fibers
