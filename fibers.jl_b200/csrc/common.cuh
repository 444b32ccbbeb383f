// Shared declarations of libfibers_cuda: plan layout, error plumbing, small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/fibers_cuda.h"

#define FIBERS_VERSION 100

namespace fibers {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const std::string& msg);
int  fail(int code, const std::string& msg);
void count_launch(int n = 1);

#define FB_CUDA(expr)                                                                       \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return ::fibers::fail(_e == cudaErrorMemoryAllocation ? FIBERS_ERR_NOMEM        \
                                                                   : FIBERS_ERR_CUDA,       \
                                  std::string(#expr) + ": " + cudaGetErrorString(_e));      \
    } while (0)

enum PlanKind { PLAN_DTI = 1, PLAN_ADC = 2, PLAN_GQI = 3, PLAN_DSI = 4 };

// Neighbour table width (max folded degree is 6 on sphere_362/642, 7 on sphere_724).
constexpr int NBR_W = 8;
constexpr uint16_t NBR_NONE = 0xFFFF;

// Per-protocol constants resident on one device (GPU analogue of GQIwork/DSIwork/DTIwork).
struct Plan {
    int kind = 0;
    int device = 0;
    int nvol = 0;        // K: number of DWI volumes
    int nvert = 0;       // M: ODF vertices on the half sphere (GQI/DSI)
    int rows = 0;        // rows of the reconstruction matrix: M (GQI), M+nvol (DSI), 7 (DTI), 2 (ADC)
    int rows_pad = 0;    // rows rounded up to a multiple of 16
    int kernel = FIBERS_KERNEL_SIMT;
    // DSI: odf/pdf are divided by den = dscale * max(s[cvol],0)
    int   cvol = -1;
    float dscale = 0.f;
    std::vector<float> h_matrix;   // row-major [rows][nvol] (host copy, for tests / introspection)
    std::vector<uint16_t> h_nbr;   // host copy of the neighbour table [nvert][NBR_W]
    // device buffers
    float*    d_mt = nullptr;      // transposed, zero padded: [nvol][rows_pad]  (SIMT path operand)
    float*    d_pinv = nullptr;    // DTI/ADC: [rows][nvol]
    float*    d_design = nullptr;  // DTI/ADC: design matrix A [nvol][rows] (partial-sample path)
    uint8_t*  d_ib0 = nullptr;     // DTI/ADC: [nvol]
    int       nb0 = 0;             // DTI/ADC: number of minimum-b volumes (ib0 flags set)
    uint16_t* d_nbr = nullptr;     // [nvert][NBR_W]
    int       nbr_width = NBR_W;   // max folded-mesh degree actually present
    float*    d_vert = nullptr;    // first-half vertices [nvert][3] row-major (peak vectors)
    int*      d_list = nullptr;    // DTI partial-path voxel list (grown on demand)
    int64_t   list_cap = 0;
    int*      d_count = nullptr;   // DTI partial-path counters (two, used alternately: see launch_fit)
    unsigned  count_flip = 0;
    void*     tc = nullptr;        // tensor-core path state (recon_tc.cu), or null
    cudaEvent_t ev_done = nullptr; // end of the plan's most recent launch (see plan_enter / plan_leave)
};

// Launches of ONE plan share its device scratch (work lists, counters, scale): their kernel sections are
// serialised across streams with an event chain, while the copies queued on those streams still overlap.
inline int plan_enter(Plan* p, cudaStream_t st) {
    if (p->ev_done) FB_CUDA(cudaStreamWaitEvent(st, p->ev_done, 0));
    return 0;
}
inline int plan_leave(Plan* p, cudaStream_t st) {
    if (!p->ev_done) FB_CUDA(cudaEventCreateWithFlags(&p->ev_done, cudaEventDisableTiming));
    FB_CUDA(cudaEventRecord(p->ev_done, st));
    return 0;
}

// ---- host-side set-up (setup.cpp) -------------------------------------------------------
// All return "" on success or an error message.
std::string build_dti_design(int nvol, const float* bval, const float* bvec,
                             std::vector<float>& A /*[nvol][7]*/, std::vector<float>& pA /*[7][nvol]*/,
                             std::vector<uint8_t>& ib0);
std::string build_adc_design(int nvol, const float* bval, std::vector<float>& A, std::vector<float>& pA,
                             std::vector<uint8_t>& ib0);
std::string build_gqi_matrix(int nvol, const float* bval, const float* bvec, const float* vertices,
                             int nvert2, float sigma, std::vector<float>& A /*[M][nvol]*/);
std::string build_dsi_matrix(int nvol, const float* bval, const float* bvec, const float* vertices,
                             int nvert2, int hann_width, std::vector<float>& MoMp /*[M+nvol][nvol]*/,
                             int& cvol, float& dscale);
std::string build_neighbours(const int32_t* faces, int nface, int nvert,
                             std::vector<uint16_t>& nbr /*[nvert][NBR_W]*/);
// Moore-Penrose pseudo-inverse of a row-major [m][n] matrix (m >= n), float64, via the
// eigen-decomposition of A'A (cyclic Jacobi) with LAPACK-pinv style truncation.
void pinv_rowmajor(const double* A, int m, int n, double rtol, double* pA /*[n][m]*/);

// ---- kernels launchers ----------------------------------------------------------------
int launch_dti(Plan* p, const float* d_dwi, int64_t dwi_pitch, const uint8_t* d_mask, int64_t nvox,
               int64_t out_pitch, float* const out[10], uint8_t* d_valid, cudaStream_t st);
int launch_adc(Plan* p, const float* d_dwi, int64_t dwi_pitch, const uint8_t* d_mask, int64_t nvox,
               float* d_adc, float* d_s0, cudaStream_t st);

struct ReconArgs {
    const float* dwi; int64_t dwi_pitch; const uint8_t* mask; int64_t nvox; int64_t out_pitch;
    float* pdf; float* odf; float* peak[3]; float* qa[3]; int16_t* peak_idx; int32_t* stats;
};
int launch_recon_simt(Plan* p, const ReconArgs& a, cudaStream_t st);
int launch_recon_tc(Plan* p, const ReconArgs& a, cudaStream_t st);     // recon_tc.cu
int tc_plan_init(Plan* p);                                              // returns 0 if TC path usable
void tc_plan_free(Plan* p);
int launch_stats_init(int32_t* d_stats, cudaStream_t st);
int launch_qa_scale(float* q1, float* q2, float* q3, int64_t nvox, const int32_t* d_stats, float odfmax,
                    cudaStream_t st);
int launch_convert(const void* src, int dtype, float* dst, int64_t n, cudaStream_t st);

// ---- order-preserving float <-> int encoding for atomicMax on floats -------------------
__host__ __device__ inline int32_t f2ord(float f) {
    int32_t i; memcpy(&i, &f, 4);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__host__ __device__ inline float ord2f(int32_t i) {
    i = i >= 0 ? i : i ^ 0x7FFFFFFF;
    float f; memcpy(&f, &i, 4); return f;
}
constexpr int32_t ORD_NEG_INF = (int32_t)0x807FFFFF;   // f2ord(-inf) = 0xFF800000 ^ 0x7FFFFFFF

}  // namespace fibers
