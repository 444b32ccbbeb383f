// umma_probe: stand-alone validation of the tcgen05 building blocks used by recon_tc.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu && ./umma_probe
// Each test computes D[m][n] = sum_k A[m][k] * B[n][k] (fp16 inputs, fp32 accumulate) with
//   A (the voxel operand) written to TENSOR MEMORY with tcgen05.st   (TS-mode MMA)
//   B (the reconstruction matrix) in shared memory, K-major, in one of the canonical layouts
// and compares against a host reference.  Tests:
//   mode 0: SWIZZLE_NONE (interleaved core matrices), K16 chunks, manual fill
//   mode 1: SWIZZLE_32B, K16 chunks, manual fill
//   mode 2: SWIZZLE_64B, K32 chunks, manual fill
//   mode 3: SWIZZLE_128B, K64 chunks, manual fill
//   mode 4: SWIZZLE_32B filled by TMA (cp.async.bulk.tensor.2d)
//   mode 5: cta_group::2 (CTA pair, M = 256), SWIZZLE_32B by TMA, B split across the pair
// All waits are bounded; a timeout sets an error flag instead of hanging the GPU.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t phase, int* err) {
    for (int i = 0; i < 2000000; ++i) if (mbar_try_wait(bar, phase)) return true;
    atomicExch(err, 1);
    return false;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) {
    if (CG == 1) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols) : "memory");
    else asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols) : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_relinquish() {
    if (CG == 1) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    else asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t r[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t r[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]
template <int CG>
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
template <int CG>
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// instruction descriptor: kind::f16, A/B = f16, D = f32, A K-major (TMEM), B K-major
__host__ __device__ inline uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// shared-memory matrix descriptor
__device__ inline uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                 // version = 1 (sm_100)
    d |= (uint64_t)layout_type << 61;
    return d;
}

struct ProbeParams {
    const __half* A;   // [Mtot][K]
    const __half* B;   // [N][K]
    float* D;          // [Mtot][N]
    int K, N, mode;
    int* err;
};

// byte offset of element (row, k) inside a K-chunk tile of `rows` rows, chunk width CW elements
__device__ inline uint32_t b_offset(int mode, int rows, int row, int k) {
    if (mode == 0) {            // interleaved: [k/8][row][8 elems]
        return (uint32_t)((k >> 3) * rows * 16 + row * 16 + (k & 7) * 2);
    }
    int rowbytes = mode == 1 ? 32 : (mode == 2 ? 64 : 128);
    uint32_t lin = (uint32_t)(row * rowbytes + k * 2);
    uint32_t mask = mode == 1 ? 1u : (mode == 2 ? 3u : 7u);
    return lin ^ (((lin >> 7) & mask) << 4);
}

template <int CG>
__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeParams p, const __grid_constant__ CUtensorMap tmapB) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_mma, bar_tma;
    __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int K = p.K, N = p.N;
    const int mode = p.mode >= 4 ? 1 : p.mode;                 // TMA tests use SWIZZLE_32B
    const int CW = mode <= 1 ? 16 : (mode == 2 ? 32 : 64);     // K elements per smem chunk
    const int nrows = N / CG;                                  // B rows held by this CTA
    const int chunk_bytes = nrows * CW * 2;
    const int nchunk = K / CW;
    uint8_t* sB = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);

    if (t == 0) { mbar_init(&bar_mma, 1); mbar_init(&bar_tma, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) { tmem_alloc<CG>(&tmem_base_s, 512); tmem_relinquish<CG>(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t tmem_d = tmem_base, tmem_a = tmem_base + 256;

    // ---- A: thread t owns row (lane) t of this CTA; pack K halves into K/2 columns ------------
    {
        const __half* arow = p.A + (size_t)(rank * 128 + t) * K;
        for (int k0 = 0; k0 < K; k0 += 16) {
            uint32_t r[8];
            for (int j = 0; j < 8; ++j) {
                __half2 h = __halves2half2(arow[k0 + 2 * j], arow[k0 + 2 * j + 1]);    // low half = even k
                r[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            tmem_st8(tmem_a + ((uint32_t)(warp * 32) << 16) + k0 / 2, r);
        }
        tmem_wait_st();
    }
    // ---- B ----------------------------------------------------------------------------------
    if (p.mode < 4) {
        for (int i = t; i < nrows * K; i += 128) {
            int row = i / K, k = i % K;
            int c = k / CW, kk = k % CW;
            *(__half*)(sB + c * chunk_bytes + b_offset(mode, nrows, row, kk)) = p.B[(size_t)(rank * nrows + row) * K + k];
        }
        fence_async_smem();
    } else if (t == 0) {
        mbar_expect_tx(&bar_tma, (uint32_t)(nchunk * chunk_bytes));
        for (int c = 0; c < nchunk; ++c) {
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(sB + c * chunk_bytes)), "l"(&tmapB), "r"(c * CW), "r"((int)(rank * nrows)), "r"(smem_u32(&bar_tma)) : "memory");
        }
    }
    if (p.mode >= 4) mbar_wait_bounded(&bar_tma, 0, p.err);
    tc_fence_before();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    else __syncthreads();
    tc_fence_after();

    // ---- MMA (one thread of the leader CTA) ---------------------------------------------------
    if (rank == 0 && t == 0) {
        const uint32_t idesc = make_idesc(128 * CG, N);
        const uint32_t ltype = mode == 0 ? 0u : (mode == 1 ? 6u : (mode == 2 ? 4u : 2u));
        for (int k0 = 0; k0 < K; k0 += 16) {
            int c = k0 / CW, kk = k0 % CW;
            uint32_t base = smem_u32(sB + c * chunk_bytes);
            uint64_t desc;
            if (mode == 0) desc = make_sdesc(base, (uint32_t)(nrows * 16), 128, 0);
            else desc = make_sdesc(base + kk * 2, 16, (uint32_t)(8 * CW * 2), ltype);
            mma_ts<CG>(tmem_d, tmem_a + k0 / 2, desc, idesc, k0 > 0 ? 1u : 0u);
        }
        mma_commit<CG>(&bar_mma);
    }
    mbar_wait_bounded(&bar_mma, 0, p.err);
    tc_fence_after();
    // ---- D: thread t reads row t, N columns ---------------------------------------------------
    for (int n0 = 0; n0 < N; n0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + n0, r);
        tmem_wait_ld();
        for (int j = 0; j < 16; ++j) p.D[(size_t)(rank * 128 + t) * N + n0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    else __syncthreads();
    if (warp == 0) tmem_dealloc<CG>(tmem_base, 512);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int K = 64, N = 176;
    EncodeFn encode = nullptr;
    {
        cudaDriverEntryPointQueryResult qres;
        void* fn = nullptr;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        encode = (EncodeFn)fn;
    }
    int fails = 0;
    for (int mode = 0; mode <= 5; ++mode) {
        const int CG = mode == 5 ? 2 : 1;
        const int Mtot = 128 * CG;
        std::vector<__half> hA((size_t)Mtot * K), hB((size_t)N * K);
        std::vector<float> ref((size_t)Mtot * N, 0.f), out((size_t)Mtot * N, -1.f);
        srand(1234 + mode);
        for (auto& x : hA) x = __float2half((float)(rand() % 2001 - 1000) / 256.f);
        for (auto& x : hB) x = __float2half((float)(rand() % 2001 - 1000) / 1024.f);
        for (int m = 0; m < Mtot; ++m)
            for (int n = 0; n < N; ++n) {
                double a = 0;
                for (int k = 0; k < K; ++k) a += (double)__half2float(hA[(size_t)m * K + k]) * (double)__half2float(hB[(size_t)n * K + k]);
                ref[(size_t)m * N + n] = (float)a;
            }
        __half *dA, *dB; float* dD; int* dErr;
        CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, out.size() * 4)); CK(cudaMalloc(&dErr, 4));
        CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemset(dD, 0xFF, out.size() * 4)); CK(cudaMemset(dErr, 0, 4));
        CUtensorMap tmap; memset(&tmap, 0, sizeof(tmap));
        if (mode >= 4) {
            cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
            cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
            cuuint32_t box[2] = {16, (cuuint32_t)(N / CG)};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("mode %d: cuTensorMapEncodeTiled failed %d\n", mode, (int)r); ++fails; continue; }
        }
        ProbeParams p{dA, dB, dD, K, N, mode, dErr};
        size_t smem = (size_t)N * K * 2 + 2048;
        cudaError_t le;
        if (CG == 1) {
            CK(cudaFuncSetAttribute(probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            probe_kernel<1><<<1, 128, smem>>>(p, tmap);
            le = cudaGetLastError();
        } else {
            CK(cudaFuncSetAttribute(probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            le = cudaLaunchKernelEx(&cfg, probe_kernel<2>, p, tmap);
        }
        cudaError_t se = cudaDeviceSynchronize();
        int herr = -1;
        if (le != cudaSuccess || se != cudaSuccess) {
            printf("mode %d: launch/sync error: %s / %s\n", mode, cudaGetErrorString(le), cudaGetErrorString(se));
            ++fails; fflush(stdout);
            return 2;      // context is dead after a device fault
        }
        CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost));
        double maxerr = 0, maxref = 0; size_t nbad = 0;
        for (size_t i = 0; i < out.size(); ++i) {
            double e = fabs((double)out[i] - ref[i]);
            if (!(e <= 1e-3 * (1 + fabs(ref[i])))) ++nbad;
            if (e > maxerr || e != e) maxerr = e; if (fabs(ref[i]) > maxref) maxref = fabs(ref[i]);
        }
        printf("mode %d (cta_group %d): timeout_flag=%d  max|err|=%.4g (max|ref|=%.4g)  mismatches=%zu/%zu  -> %s\n", mode, CG, herr,
               maxerr, maxref, nbad, out.size(), (nbad == 0 && herr == 0) ? "PASS" : "FAIL");
        if (nbad) {
            printf("   sample: D[0][0..3] = %g %g %g %g ; ref = %g %g %g %g ; D[127][%d] = %g ref %g\n", out[0], out[1], out[2], out[3],
                   ref[0], ref[1], ref[2], ref[3], N - 1, out[(size_t)127 * N + N - 1], ref[(size_t)127 * N + N - 1]);
            ++fails;
        }
        fflush(stdout);
        cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dErr);
    }
    printf("umma_probe: %d failing mode(s)\n", fails);
    return fails ? 1 : 0;
}
