// umma_bench: micro-benchmarks that calibrate the design of recon_tc.cu on the actual part.
//   A. tcgen05.ld (TMEM -> registers) throughput for several shapes / warp counts
//   B. tcgen05.mma issue-to-completion rate for N in {64..256}, A from TMEM (TS) or SMEM (SS),
//      B tile SWIZZLE_32B (K16 chunks) or SWIZZLE_128B (K64 chunks), cta_group 1 or 2
// Results are cycles (clock64) on one SM (pair); data values are irrelevant (uninitialised operands).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t ph) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(ph) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t ph) { for (int i = 0; i < (1 << 24); ++i) if (mbar_try(bar, ph)) return; __trap(); }
template <int CG> __device__ __forceinline__ void tmem_alloc(uint32_t* dst) {
    if (CG == 1) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(512u) : "memory");
                   asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
    else { asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(512u) : "memory");
           asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
}
template <int CG> __device__ __forceinline__ void tmem_dealloc(uint32_t a) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(a), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(a), "r"(512u) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
#define LD_X16(r, a) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
  : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]) : "r"(a) : "memory")
#define LD_X32(r, a) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
  : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]), \
    "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]) : "r"(a) : "memory")

// ---- A: LDTM throughput.  mode 0: x16 + wait each; 1: x32 + wait each; 2: x32, wait once per 64 columns
__global__ void __launch_bounds__(256, 1) ldtm_kernel(int nwarps, int mode, int ncols, int reps, long long* out, float* sink) {
    __shared__ uint32_t tb;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<1>(&tb);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps) {
        const int c_begin = (warp >> 2) * (ncols / ((nwarps + 3) / 4)), c_end = c_begin + ncols / ((nwarps + 3) / 4);
        for (int rep = 0; rep < reps; ++rep) {
            if (mode == 0) {
                for (int c = c_begin; c < c_end; c += 16) { uint32_t r[16]; LD_X16(r, base + c); wait_ld();
                    #pragma unroll
                    for (int j = 0; j < 16; ++j) acc += __uint_as_float(r[j]); }
            } else if (mode == 1) {
                for (int c = c_begin; c + 32 <= c_end; c += 32) { uint32_t r[32]; LD_X32(r, base + c); wait_ld();
                    #pragma unroll
                    for (int j = 0; j < 32; ++j) acc += __uint_as_float(r[j]); }
            } else {
                for (int c = c_begin; c + 64 <= c_end; c += 64) { uint32_t r[32], s[32]; LD_X32(r, base + c); LD_X32(s, base + c + 32); wait_ld();
                    #pragma unroll
                    for (int j = 0; j < 32; ++j) acc += __uint_as_float(r[j]) + __uint_as_float(s[j]); }
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc<1>(tb);
}

// ---- B: MMA rate
__host__ __device__ inline uint32_t make_idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ inline uint64_t make_sdesc(uint32_t saddr, uint32_t sbo, uint32_t lt) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)lt << 61);
}
template <int CG>
__global__ void __launch_bounds__(128, 1) mma_kernel(int N, int ts, int sw128, int nmma, int ndst, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tb;
    const int warp = threadIdx.x >> 5;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc<CG>(&tb);
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;   // halves = 1.0
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); else __syncthreads();
    tc_fence_after();
    if (rank == 0 && warp == 0) {                      // warp-uniform issue loop, one elected lane issues
        const uint32_t idesc = make_idesc(128 * CG, N);
        const uint32_t sb = smem_u32(smem);
        const uint64_t bdesc = sw128 ? make_sdesc(sb, 1024, 2) : make_sdesc(sb, 256, 6);
        const uint64_t adesc = sw128 ? make_sdesc(sb + 32768, 1024, 2) : make_sdesc(sb + 32768, 256, 6);
        const uint32_t d = tb, a = tb + 480u;
        long long t0 = clock64();
        for (int i = 0; i < nmma; ++i) {
            uint32_t el;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(el));
            if (el) {
                if (ndst == 1) {          // plain form (no disable-output-lane vector)
                    if (ts) {
                        if (CG == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
                        else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
                    } else {
                        if (CG == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
                        else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
                    }
                } else {                  // form with the disable-output-lane mask vector
                    if (ts) {
                        if (CG == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5,%5,%5,%5}, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u) : "memory");
                        else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5,%5,%5,%5,%5,%5,%5,%5}, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u) : "memory");
                    } else {
                        if (CG == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5,%5,%5,%5}, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u) : "memory");
                        else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5,%5,%5,%5,%5,%5,%5,%5}, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u) : "memory");
                    }
                }
            }
            __syncwarp();
        }
        long long t1 = clock64();
        uint32_t el;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(el));
        if (el) {
            if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    } else mbar_wait(&bar, 0);
    tc_fence_before();
    if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); else __syncthreads();
    if (warp == 0) tmem_dealloc<CG>(tb);
}

// ---- C: does accumulating into the SAME D back to back serialise?  6 MMAs per elected issue block (as in the
//         production kernel), destinations cycling over `nacc` accumulators of N columns each.
template <int NACC, int ORDER>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma_alt_kernel(int N, int nblk, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tb;
    const int warp = threadIdx.x >> 5;
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc<2>(&tb);
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    tc_fence_after();
    if (rank == 0 && warp == 0) {
        const uint32_t idesc = make_idesc(256, N);
        const uint32_t sb = smem_u32(smem);
        const uint64_t bdesc = make_sdesc(sb, 256, 6);
        const uint32_t d = tb, a = tb + 480u;
        long long t0 = clock64();
        for (int i = 0; i < nblk; ++i) {
            uint32_t el;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(el));
            if (el) {
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    // order 0: d0 d0 d0 d1 d1 d1 (grouped)   order 1: d0 d1 d2 d0 d1 d2 (interleaved)
                    constexpr int dummy = 0; (void)dummy; const int k = ORDER == 0 ? (j / (6 / NACC)) : (j % NACC);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(d + (uint32_t)(k * N)), "r"(a), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
                }
            }
            __syncwarp();
        }
        long long t1 = clock64();
        uint32_t el;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(el));
        if (el) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    } else mbar_wait(&bar, 0);
    tc_fence_before();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) tmem_dealloc<2>(tb);
}

int main() {
    long long* dout; float* sink; CK(cudaMalloc(&dout, 64)); CK(cudaMalloc(&sink, 4));
    long long h[2];
    printf("== A. tcgen05.ld: 128 lanes x 336 columns (172 KB) per pass, cycles per pass ==\n");
    for (int nw : {4, 8}) for (int mode : {0, 1, 2}) {
        int ncols = 320 + (mode == 0 ? 16 : 0);   // 336 for x16; 320 for the x32/x64 variants
        ldtm_kernel<<<1, 256>>>(nw, mode, ncols, 50, dout, sink); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, dout, 8, cudaMemcpyDeviceToHost));
        printf("  warps=%d mode=%s cols=%d : %.0f cycles/pass  (%.1f B/cycle)\n", nw, mode == 0 ? "x16+wait" : mode == 1 ? "x32+wait" : "2*x32+wait",
               ncols, h[0] / 50.0, 128.0 * ncols * 4 / (h[0] / 50.0));
    }
    printf("== B. tcgen05.mma: cycles per MMA (issue loop / until commit completes), 256 MMAs ==\n");
    const size_t smem = 64 * 1024 + 1024;
    CK(cudaFuncSetAttribute(mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int cg : {1, 2}) for (int ts : {1, 0}) for (int sw128 : {0, 1}) for (int N : {64, 160, 176, 256}) for (int ndst : {1, 2}) {
        const int nmma = 256;
        if (cg == 1) mma_kernel<1><<<1, 128, smem>>>(N, ts, sw128, nmma, ndst, dout);
        else {
            cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, mma_kernel<2>, N, ts, sw128, nmma, ndst, dout));
        }
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost));
        const double ideal = 128.0 * N / 256.0;      // per-SM 4096 MAC/cycle
        printf("  cta_group=%d A=%s B=%s N=%3d form=%s : issue %.0f, total %.0f cycles/MMA (ideal %.0f)\n", cg, ts ? "TMEM" : "SMEM",
               sw128 ? "SW128" : "SW32 ", N, ndst == 1 ? "plain" : "mask ", h[0] / (double)nmma, h[1] / (double)nmma, ideal);
    }
    printf("== C. cta_group::2 TS-mode, 6 MMAs per issue block, destinations over nacc accumulators ==\n");
    for (int N : {64, 112, 160}) for (int nacc : {1, 2, 3}) for (int order : {0, 1}) {
        const int nblk = 64;
        void (*kf)(int, int, long long*) = nacc == 1 ? (order ? mma_alt_kernel<1, 1> : mma_alt_kernel<1, 0>)
                                         : nacc == 2 ? (order ? mma_alt_kernel<2, 1> : mma_alt_kernel<2, 0>)
                                                     : (order ? mma_alt_kernel<3, 1> : mma_alt_kernel<3, 0>);
        CK(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kf<<<2, 128, smem>>>(N, nblk, dout);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost));
        printf("  N=%3d nacc=%d %s : issue %.0f, total %.0f cycles/MMA (ideal %.0f)\n", N, nacc, order ? "interleaved" : "grouped    ",
               h[0] / (6.0 * nblk), h[1] / (6.0 * nblk), 128.0 * N / 256.0);
    }
    return 0;
}
