// Per-SM global-store bandwidth probe (B200): how fast can ONE SM push an ODF-shaped tile (rows of 128 voxels x 4 B,
// row pitch = the volume) to L2 / HBM, with plain STG.32 from 12 warps, with STG.128, and with bulk (TMA) stores
// from shared memory -- alone on the chip and with all 148 SMs doing the same.  Explains the 6.4 k-cycle TMEM drain
// of recon_tc_kernel (172 KB per tile and SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/store_probe tools/store_probe.cu && tools/_bin/store_probe
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ROWS = 336, VOX = 128;

// mode 0: STG.32, thread = voxel (warp w of 12: rows w*28 ..), one row per instruction  (what the kernel does)
// mode 1: STG.128, thread = 4 voxels, a warp covers one 512-byte row
// mode 2: cp.async.bulk shared -> global, one 512-byte row per copy, issued by one lane per warp
// mode 3: bulk TENSOR stores, box 32 voxels x 16 rows (2 KB) per warp from a 2-box ring, exactly what recon_tc_kernel<true> does
// mode 4: bulk TENSOR stores, box 128 voxels x 16 rows (8 KB), one issuing lane per group of 4 warps
__global__ void __launch_bounds__(384, 1) store_probe(float* out, int64_t pitch, int tiles, int mode, long long* cyc,
                                                      const __grid_constant__ CUtensorMap m32, const __grid_constant__ CUtensorMap m128) {
    __shared__ __align__(128) float stage[12][VOX];
    __shared__ __align__(128) float box[12][2][16 * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 12 * VOX; i += blockDim.x) (&stage[0][0])[i] = (float)i;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
        float* base = out + ((int64_t)blockIdx.x * tiles + t) * VOX;
        if (mode == 0) {
            const int q = warp & 3, part = warp >> 2;                   // lane quarter, column part (112 rows each)
            float* g = base + (int64_t)(part * 112) * pitch + q * 32 + lane;
#pragma unroll 16
            for (int r = 0; r < 112; ++r) { *g = (float)(r + t); g += pitch; }
        } else if (mode == 1) {
            float4* g = reinterpret_cast<float4*>(base + (int64_t)(warp * 28) * pitch) + lane;
#pragma unroll 14
            for (int r = 0; r < 28; ++r) { *g = make_float4((float)r, 1.f, 2.f, (float)t); g = reinterpret_cast<float4*>(reinterpret_cast<float*>(g) + pitch); }
        } else if (mode == 3) {
            const int q = warp & 3, part = warp >> 2;
            for (int ch = 0; ch < 7; ++ch) {
                const int b = (t * 7 + ch) & 1;
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                for (int j = 0; j < 16; ++j) box[warp][b][j * 32 + lane] = (float)(j + t);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                                 ::"l"(&m32), "r"((int)((blockIdx.x * tiles + t) * VOX + q * 32)), "r"(part * 112 + ch * 16),
                                   "r"((uint32_t)__cvta_generic_to_shared(&box[warp][b][0])) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        } else if (mode == 4) {
            const int q = warp & 3, part = warp >> 2;
            float* gb = &box[part * 4][0][0];                            // 4 warps x 2 x 2 KB = two 8 KB boxes per group
            for (int ch = 0; ch < 7; ++ch) {
                const int b = (t * 7 + ch) & 1;
                if (q == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                for (int j = 0; j < 16; ++j) gb[b * 2048 + j * 128 + q * 32 + lane] = (float)(j + t);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(1 + part) : "memory");
                if (q == 0 && lane == 0) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                                 ::"l"(&m128), "r"((int)((blockIdx.x * tiles + t) * VOX)), "r"(part * 112 + ch * 16),
                                   "r"((uint32_t)__cvta_generic_to_shared(gb + b * 2048)) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        } else {
            if (lane == 0) {
                const uint32_t s = (uint32_t)__cvta_generic_to_shared(&stage[warp][0]);
                for (int r = 0; r < 28; ++r) {
                    float* g = base + (int64_t)(warp * 28 + r) * pitch;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 512;" ::"l"(g), "r"(s) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
            }
            __syncwarp();
        }
    }
    if (mode >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    const int tiles = 96;
    const int64_t pitch = 148ll * tiles * VOX;                 // every CTA owns its own voxel range of every row
    float* out; long long* cyc;
    CK(cudaMalloc(&out, sizeof(float) * pitch * ROWS));
    CK(cudaMalloc(&cyc, sizeof(long long) * 148));
    long long h[148];
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    CUtensorMap m32, m128;
    cuuint64_t dims[2] = {(cuuint64_t)pitch, ROWS}, strd[1] = {(cuuint64_t)pitch * 4};
    cuuint32_t b32[2] = {32, 16}, b128[2] = {128, 16}, es[2] = {1, 1};
    if (((EncodeFn)fn)(&m32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strd, b32, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        ((EncodeFn)fn)(&m128, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strd, b128, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("tensor map encode failed\n"); return 1; }
    const char* names[5] = {"STG.32 (thread = voxel, 12 warps)", "STG.128 (warp = one 512 B row)", "cp.async.bulk 512 B rows from smem",
                            "tensor store 32 x 16 boxes, per warp", "tensor store 128 x 16 boxes, per 4 warps"};
    for (int grid : {1, 148})
        for (int mode = 0; mode < 5; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                store_probe<<<grid, 384>>>(out, pitch, tiles, mode, cyc, m32, m128);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            CK(cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            const double bytes = (double)tiles * ROWS * VOX * 4;
            printf("grid %3d  %-40s  %7.0f cycles per 172 KB tile  %5.1f B/clk/SM  chip %.0f GB/s (%.3f ms)\n", grid, names[mode],
                   (double)mx / tiles, bytes / (double)mx, bytes * grid / (ms * 1e-3) / 1e9, ms);
        }
    return 0;
}
