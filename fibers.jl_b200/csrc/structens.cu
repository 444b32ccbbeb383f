// Structure-tensor reconstruction: st_eigen and st_recon of the reference
// (src/structens.jl:13-34, :40-88).
//   st_eigen : per-voxel closed-form eigen-decomposition of the symmetric tensor (Sxx .. Szz), eigenvalues ascending,
//              eigvec[x,y,z,:,k] = k-th eigenvector (StaticArrays `eigen(Symmetric(S, :L))`): the same device function as the
//              DTI fit; one thread per voxel, 6 coalesced loads and 12 coalesced stores (24 + 48 B / voxel: HBM-bound).
//   st_recon : Gaussian pre-smoothing (sigma), Scharr gradients, outer products, Gaussian smoothing of the six products
//              (rho), st_eigen.  The filters are ImageFiltering.jl's separable factors applied with imfilter(..., "reflect"):
//              KernelFactors.gaussian(s) = exp(-x^2 / 2 s^2) on 4 ceil(s) + 1 taps, normalised; KernelFactors.scharr = derivative
//              [-1, 0, 1] / 2 along the gradient axis and [3, 10, 3] / 16 along the other two; correlation; "reflect" mirrors
//              about the edge sample without repeating it (restated from the package's published behaviour: it is not vendored).
//              Every 1-D pass is one thread per voxel reading its taps along the axis (neighbouring threads share lines).
// This translation unit is compiled with -fmad=false (eig3.cuh).
#include <algorithm>
#include <cmath>
#include <vector>
#include "common.cuh"
#include "eig3.cuh"

namespace fibers {
namespace {

__global__ void st_eigen_kernel(const float* __restrict__ sxx, const float* __restrict__ sxy, const float* __restrict__ sxz,
                                const float* __restrict__ syy, const float* __restrict__ syz, const float* __restrict__ szz, int64_t nvox,
                                float* __restrict__ evec /*[3 comp][3 k][nvox] = [nx,ny,nz,3,3] column-major*/, float* __restrict__ eval /*[3][nvox]*/) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    float w[3], q[3][3];
    eig3_sym(sxx[v], sxy[v], sxz[v], syy[v], syz[v], szz[v], w, q);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        eval[(int64_t)k * nvox + v] = w[k];
#pragma unroll
        for (int i = 0; i < 3; ++i) evec[((int64_t)k * 3 + i) * nvox + v] = q[i][k];       // frame index = i + 3 k
    }
}

constexpr int MAXTAP = 129;
struct Taps { float w[MAXTAP]; int n; };

__device__ __forceinline__ int reflect(int i, int n) {            // dcb|abcd|cba ; repeated for kernels longer than the axis
    if (n == 1) return 0;
    const int period = 2 * n - 2;
    i %= period; if (i < 0) i += period;
    return i < n ? i : period - i;
}

// out[p] = sum_t w[t] * in[p + (t - n/2) along axis]  (correlation, mirrored borders)
__global__ void st_filter1d_kernel(const float* __restrict__ in, float* __restrict__ out, int nx, int ny, int nz, int axis, const __grid_constant__ Taps tp) {
    const int64_t nvox = (int64_t)nx * ny * nz;
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    const int x = (int)(v % nx), y = (int)((v / nx) % ny), z = (int)(v / ((int64_t)nx * ny));
    const int len = axis == 0 ? nx : axis == 1 ? ny : nz, pos = axis == 0 ? x : axis == 1 ? y : z;
    const int64_t stride = axis == 0 ? 1 : axis == 1 ? nx : (int64_t)nx * ny;
    const int64_t base = v - (int64_t)pos * stride;
    const int h = tp.n / 2;
    float acc = 0.f;
    for (int t = 0; t < tp.n; ++t) acc += tp.w[t] * in[base + (int64_t)reflect(pos + t - h, len) * stride];
    out[v] = acc;
}

__global__ void st_products_kernel(const float* __restrict__ gx, const float* __restrict__ gy, const float* __restrict__ gz, int64_t nvox,
                                   float* __restrict__ xx, float* __restrict__ xy, float* __restrict__ xz, float* __restrict__ yy,
                                   float* __restrict__ yz, float* __restrict__ zz) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    const float a = gx[v], b = gy[v], c = gz[v];
    xx[v] = a * a; xy[v] = a * b; xz[v] = a * c; yy[v] = b * b; yz[v] = b * c; zz[v] = c * c;
}

Taps gaussian_taps(double sigma) {                 // KernelFactors.gaussian(sigma): l = 4 ceil(sigma) + 1
    Taps t{}; t.n = 4 * (int)std::ceil(sigma) + 1;
    const int h = t.n / 2; double s = 0;
    std::vector<double> w(t.n);
    for (int i = 0; i < t.n; ++i) { const double x = i - h; w[i] = std::exp(-x * x / (2 * sigma * sigma)); s += w[i]; }
    for (int i = 0; i < t.n; ++i) t.w[i] = (float)(w[i] / s);
    return t;
}

struct Dev { std::vector<void*> p; ~Dev() { for (void* q : p) cudaFree(q); }
             cudaError_t alloc(float** o, size_t n) { void* q = nullptr; cudaError_t e = cudaMalloc(&q, n * sizeof(float)); if (e == cudaSuccess) p.push_back(q); *o = (float*)q; return e; } };

}  // namespace
}  // namespace fibers

using namespace fibers;
#define S_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return fail(_e == cudaErrorMemoryAllocation ? FIBERS_ERR_NOMEM : FIBERS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)

static int st_check_device(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return fail(FIBERS_ERR_NODEV, "no CUDA device available (libfibers_cuda has no CPU fallback)"); }
    if (device < 0 || device >= n) return fail(FIBERS_ERR_ARG, "device ordinal out of range");
    return 0;
}

extern "C" int fibers_st_eigen_device(const float* d_sxx, const float* d_sxy, const float* d_sxz, const float* d_syy, const float* d_syz,
                                      const float* d_szz, int64_t nvox, float* d_eigvec, float* d_eigval, void* stream) {
    if (!d_sxx || !d_sxy || !d_sxz || !d_syy || !d_syz || !d_szz || !d_eigvec || !d_eigval) return fail(FIBERS_ERR_ARG, "NULL device pointer");
    if (nvox <= 0) return 0;
    st_eigen_kernel<<<(unsigned)((nvox + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_sxx, d_sxy, d_sxz, d_syy, d_syz, d_szz, nvox, d_eigvec, d_eigval);
    count_launch(1);
    S_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int fibers_st_eigen(const float* sxx, const float* sxy, const float* sxz, const float* syy, const float* syz, const float* szz,
                               int nx, int ny, int nz, float* eigvec, float* eigval, int device) {
    if (!sxx || !sxy || !sxz || !syy || !syz || !szz || !eigvec || !eigval) return fail(FIBERS_ERR_ARG, "NULL pointer");
    if (nx <= 0 || ny <= 0 || nz <= 0) return fail(FIBERS_ERR_ARG, "volume dimensions must be positive");
    if (int rc = st_check_device(device)) return rc;
    S_CUDA(cudaSetDevice(device));
    const int64_t nvox = (int64_t)nx * ny * nz;
    Dev D; float* in[6]; float *dv, *dw;
    const float* h[6] = {sxx, sxy, sxz, syy, syz, szz};
    for (int i = 0; i < 6; ++i) { S_CUDA(D.alloc(&in[i], (size_t)nvox)); S_CUDA(cudaMemcpy(in[i], h[i], sizeof(float) * nvox, cudaMemcpyHostToDevice)); }
    S_CUDA(D.alloc(&dv, (size_t)9 * nvox)); S_CUDA(D.alloc(&dw, (size_t)3 * nvox));
    if (int rc = fibers_st_eigen_device(in[0], in[1], in[2], in[3], in[4], in[5], nvox, dv, dw, nullptr)) return rc;
    S_CUDA(cudaMemcpy(eigvec, dv, sizeof(float) * 9 * nvox, cudaMemcpyDeviceToHost));
    S_CUDA(cudaMemcpy(eigval, dw, sizeof(float) * 3 * nvox, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int fibers_st_recon(const float* vol, int nx, int ny, int nz, float sigma, float rho, float* eigvec, float* eigval, int device) {
    if (!vol || !eigvec || !eigval) return fail(FIBERS_ERR_ARG, "NULL pointer");
    if (nx <= 0 || ny <= 0 || nz <= 0) return fail(FIBERS_ERR_ARG, "volume dimensions must be positive");
    if (4 * (int)std::ceil(std::max(sigma, rho)) + 1 > MAXTAP) return fail(FIBERS_ERR_ARG, "sigma / rho too large (more than 129 filter taps)");
    if (int rc = st_check_device(device)) return rc;
    S_CUDA(cudaSetDevice(device));
    const int64_t nvox = (int64_t)nx * ny * nz;
    const unsigned grid = (unsigned)((nvox + 255) / 256);
    Dev D; float *img, *t0, *t1, *g[3], *s[6], *dv, *dw;
    S_CUDA(D.alloc(&img, (size_t)nvox)); S_CUDA(D.alloc(&t0, (size_t)nvox)); S_CUDA(D.alloc(&t1, (size_t)nvox));
    for (auto& q : g) S_CUDA(D.alloc(&q, (size_t)nvox));
    for (auto& q : s) S_CUDA(D.alloc(&q, (size_t)nvox));
    S_CUDA(cudaMemcpy(img, vol, sizeof(float) * nvox, cudaMemcpyHostToDevice));
    auto pass = [&](const float* in, float* out, int axis, const Taps& tp) { st_filter1d_kernel<<<grid, 256>>>(in, out, nx, ny, nz, axis, tp); count_launch(1); };
    // separable filter = the three factors applied along x, y, z in turn; result in `dst` (may equal src)
    auto sep = [&](float* src, float* dst, const Taps& fx, const Taps& fy, const Taps& fz) { pass(src, t0, 0, fx); pass(t0, t1, 1, fy); pass(t1, dst, 2, fz); };
    if (sigma > 0) { const Taps gs = gaussian_taps(sigma); sep(img, img, gs, gs, gs); }                 // (:44-49)
    Taps der{}; der.n = 3; der.w[0] = -0.5f; der.w[1] = 0.f; der.w[2] = 0.5f;
    Taps smo{}; smo.n = 3; smo.w[0] = 3.f / 16.f; smo.w[1] = 10.f / 16.f; smo.w[2] = 3.f / 16.f;
    sep(img, g[0], der, smo, smo); sep(img, g[1], smo, der, smo); sep(img, g[2], smo, smo, der);        // (:52-57)
    st_products_kernel<<<grid, 256>>>(g[0], g[1], g[2], nvox, s[0], s[1], s[2], s[3], s[4], s[5]);          // (:60-65)
    count_launch(1);
    if (rho > 0) { const Taps gr = gaussian_taps(rho); for (auto& q : s) sep(q, q, gr, gr, gr); }        // (:69-83)
    S_CUDA(D.alloc(&dv, (size_t)9 * nvox)); S_CUDA(D.alloc(&dw, (size_t)3 * nvox));
    if (int rc = fibers_st_eigen_device(s[0], s[1], s[2], s[3], s[4], s[5], nvox, dv, dw, nullptr)) return rc;
    S_CUDA(cudaGetLastError());
    S_CUDA(cudaMemcpy(eigvec, dv, sizeof(float) * 9 * nvox, cudaMemcpyDeviceToHost));
    S_CUDA(cudaMemcpy(eigval, dw, sizeof(float) * 3 * nvox, cudaMemcpyDeviceToHost));
    return 0;
}
