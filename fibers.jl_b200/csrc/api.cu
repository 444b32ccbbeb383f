// C ABI of libfibers_cuda (include/fibers_cuda.h): plans, device-resident entry points, and the
// host-pointer entry points with the z-slab partitioner (one host worker thread + stream ring per
// GPU, pitched H2D / D2H copies overlapped with the kernels, no inter-GPU collective).
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <algorithm>
#include <cmath>
#include "common.cuh"
#include "host_pipeline.h"

namespace fibers {

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};
static std::mutex g_cfg_mu;
static std::vector<int> g_devices;
static bool g_devices_init = false;
static int g_kernel = FIBERS_KERNEL_AUTO;

void set_error(const std::string& msg) { g_err = msg; }
const std::string& last_error() { return g_err; }
int fail(int code, const std::string& msg) { g_err = msg; return code; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int device_count_raw() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

std::vector<int> device_list() {
    std::lock_guard<std::mutex> lk(g_cfg_mu);
    if (!g_devices_init) {
        g_devices_init = true;
        const char* env = getenv("FIBERS_CUDA_DEVICES");
        if (env && *env) {
            const char* s = env;
            while (*s) {
                char* end;
                long v = strtol(s, &end, 10);
                if (end == s) break;
                g_devices.push_back((int)v);
                s = (*end == ',') ? end + 1 : end;
            }
        }
        if (g_devices.empty()) {
            int n = device_count_raw();
            for (int i = 0; i < n; ++i) g_devices.push_back(i);
        }
    }
    return g_devices;
}

static int kernel_choice() {
    std::lock_guard<std::mutex> lk(g_cfg_mu);
    int k = g_kernel;
    if (k == FIBERS_KERNEL_AUTO) {
        const char* env = getenv("FIBERS_CUDA_KERNEL");
        if (env && !strcmp(env, "simt")) k = FIBERS_KERNEL_SIMT;
        else if (env && !strcmp(env, "tc")) k = FIBERS_KERNEL_TC;
    }
    return k;
}

template <typename T>
static int upload(T** dptr, const std::vector<T>& h) {
    *dptr = nullptr;
    if (h.empty()) return 0;
    FB_CUDA(cudaMalloc(dptr, sizeof(T) * h.size()));
    FB_CUDA(cudaMemcpy(*dptr, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
    return 0;
}

void plan_free(Plan* p) {
    if (!p) return;
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(p->device);
    tc_plan_free(p);
    cudaFree(p->d_mt); cudaFree(p->d_pinv); cudaFree(p->d_design); cudaFree(p->d_ib0);
    cudaFree(p->d_nbr); cudaFree(p->d_vert); cudaFree(p->d_list); cudaFree(p->d_count);
    if (p->ev_done) cudaEventDestroy(p->ev_done);
    cudaSetDevice(cur);
    delete p;
}

static int plan_begin(Plan** out, int device, int kind, int nvol) {
    if (!out) return fail(FIBERS_ERR_ARG, "plan output pointer is NULL");
    *out = nullptr;
    int n = device_count_raw();
    if (n <= 0) return fail(FIBERS_ERR_NODEV, "no CUDA device available (libfibers_cuda has no CPU fallback)");
    if (device < 0 || device >= n) return fail(FIBERS_ERR_ARG, "device ordinal out of range");
    if (nvol <= 0) return fail(FIBERS_ERR_TABLE, "Missing b-value table from input DWI structure");
    FB_CUDA(cudaSetDevice(device));
    Plan* p = new Plan();
    p->kind = kind; p->device = device; p->nvol = nvol;
    *out = p;
    return 0;
}

// Builds the transposed, zero-padded operand [nvol][rows_pad] from row-major blocks.
// For DSI the pdf block starts at row round_up(M,16) so that row panels stay 16-aligned.
static int plan_upload_recon(Plan* p, const std::vector<float>& mat, const float* vertices, int nvert2,
                             const int32_t* faces, int nface) {
    const int M = p->nvert, N = p->nvol;
    const int mpad = (M + 15) / 16 * 16;
    const int extra = p->kind == PLAN_DSI ? (N + 15) / 16 * 16 : 0;
    p->rows_pad = mpad + extra;
    std::vector<float> mt((size_t)N * p->rows_pad, 0.f);
    for (int r = 0; r < M; ++r)
        for (int k = 0; k < N; ++k) mt[(size_t)k * p->rows_pad + r] = mat[(size_t)r * N + k];
    if (p->kind == PLAN_DSI)
        for (int r = 0; r < N; ++r)
            for (int k = 0; k < N; ++k) mt[(size_t)k * p->rows_pad + mpad + r] = mat[(size_t)(M + r) * N + k];
    int rc = upload(&p->d_mt, mt);
    if (rc) return rc;
    std::vector<uint16_t> nbr;
    std::string e = build_neighbours(faces, nface, M, nbr);
    if (!e.empty()) return fail(FIBERS_ERR_ARG, e);
    if ((rc = upload(&p->d_nbr, nbr))) return rc;
    p->h_nbr = nbr;
    p->nbr_width = 0;
    for (int v = 0; v < M; ++v) {
        int d = 0;
        while (d < NBR_W && nbr[(size_t)v * NBR_W + d] != NBR_NONE) ++d;
        p->nbr_width = std::max(p->nbr_width, d);
    }
    std::vector<float> vert((size_t)M * 3);
    for (int i = 0; i < M; ++i)
        for (int c = 0; c < 3; ++c) vert[(size_t)i * 3 + c] = vertices[(size_t)c * nvert2 + i];   // first half rows
    if ((rc = upload(&p->d_vert, vert))) return rc;
    p->kernel = FIBERS_KERNEL_SIMT;
    int want = kernel_choice();
    if (want != FIBERS_KERNEL_SIMT) {
        if (tc_plan_init(p) == 0) p->kernel = FIBERS_KERNEL_TC;
        else if (want == FIBERS_KERNEL_TC) return fail(FIBERS_ERR_ARG, "tensor-core kernel requested but not usable for this shape: " + g_err);
    }
    return 0;
}

static uint64_t fnv1a(uint64_t h, const void* data, size_t n) {
    const unsigned char* p = (const unsigned char*)data;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

static int to_rc(const std::string& e) {
    if (e.empty()) return 0;
    bool table = e.find("Missing") == 0;
    return fail(table ? FIBERS_ERR_TABLE : FIBERS_ERR_ARG, e);
}

}  // namespace fibers

using namespace fibers;

extern "C" {

int fibers_cuda_version(void) { return FIBERS_VERSION; }
int fibers_cuda_device_count(void) { return device_count_raw(); }
const char* fibers_cuda_last_error(void) { return g_err.c_str(); }
int64_t fibers_cuda_launch_count(void) { return g_launches.load(); }
float fibers_stats_decode_max(int32_t encoded) { return ord2f(encoded); }

int fibers_cuda_set_devices(const int* devices, int n) {
    if (n < 0 || (n > 0 && !devices)) return fail(FIBERS_ERR_ARG, "bad device list");
    int have = device_count_raw();
    for (int i = 0; i < n; ++i)
        if (devices[i] < 0 || devices[i] >= have) return fail(FIBERS_ERR_ARG, "device ordinal out of range");
    std::lock_guard<std::mutex> lk(g_cfg_mu);
    g_devices.assign(devices, devices + n);
    g_devices_init = n > 0;
    return 0;
}

int fibers_cuda_set_kernel(int kernel) {
    if (kernel < FIBERS_KERNEL_AUTO || kernel > FIBERS_KERNEL_TC) return fail(FIBERS_ERR_ARG, "unknown kernel id");
    std::lock_guard<std::mutex> lk(g_cfg_mu);
    g_kernel = kernel;
    return 0;
}

int fibers_dti_plan_create(fibers_plan** plan, int device, int nvol, const float* bval, const float* bvec) {
    Plan* p = nullptr;
    int rc = plan_begin(&p, device, PLAN_DTI, bval ? nvol : 0);
    if (rc) return rc;
    std::vector<float> A, pA; std::vector<uint8_t> ib0;
    rc = to_rc(build_dti_design(nvol, bval, bvec, A, pA, ib0));
    if (!rc) { p->rows = 7; p->h_matrix = pA; rc = upload(&p->d_pinv, pA); }
    if (!rc) rc = upload(&p->d_design, A);
    if (!rc) rc = upload(&p->d_ib0, ib0);
    if (!rc) for (uint8_t f : ib0) p->nb0 += f != 0;
    if (rc) { plan_free(p); return rc; }
    *plan = reinterpret_cast<fibers_plan*>(p);
    return 0;
}

int fibers_adc_plan_create(fibers_plan** plan, int device, int nvol, const float* bval) {
    Plan* p = nullptr;
    int rc = plan_begin(&p, device, PLAN_ADC, bval ? nvol : 0);
    if (rc) return rc;
    std::vector<float> A, pA; std::vector<uint8_t> ib0;
    rc = to_rc(build_adc_design(nvol, bval, A, pA, ib0));
    if (!rc) { p->rows = 2; p->h_matrix = pA; rc = upload(&p->d_pinv, pA); }
    if (!rc) rc = upload(&p->d_design, A);
    if (!rc) rc = upload(&p->d_ib0, ib0);
    if (!rc) for (uint8_t f : ib0) p->nb0 += f != 0;
    if (rc) { plan_free(p); return rc; }
    *plan = reinterpret_cast<fibers_plan*>(p);
    return 0;
}

int fibers_gqi_plan_create(fibers_plan** plan, int device, int nvol, const float* bval, const float* bvec,
                           const float* vertices, int nvert2, const int32_t* faces, int nface, float sigma) {
    Plan* p = nullptr;
    int rc = plan_begin(&p, device, PLAN_GQI, bval ? nvol : 0);
    if (rc) return rc;
    std::vector<float> A;
    rc = to_rc(build_gqi_matrix(nvol, bval, bvec, vertices, nvert2, sigma, A));
    if (!rc) {
        p->nvert = nvert2 / 2; p->rows = p->nvert; p->h_matrix = A;
        rc = plan_upload_recon(p, A, vertices, nvert2, faces, nface);
    }
    if (rc) { plan_free(p); return rc; }
    *plan = reinterpret_cast<fibers_plan*>(p);
    return 0;
}

int fibers_dsi_plan_create(fibers_plan** plan, int device, int nvol, const float* bval, const float* bvec,
                           const float* vertices, int nvert2, const int32_t* faces, int nface, int hann_width) {
    Plan* p = nullptr;
    int rc = plan_begin(&p, device, PLAN_DSI, bval ? nvol : 0);
    if (rc) return rc;
    std::vector<float> MM;
    rc = to_rc(build_dsi_matrix(nvol, bval, bvec, vertices, nvert2, hann_width, MM, p->cvol, p->dscale));
    if (!rc) {
        p->nvert = nvert2 / 2; p->rows = p->nvert + nvol; p->h_matrix = MM;
        rc = plan_upload_recon(p, MM, vertices, nvert2, faces, nface);
    }
    if (rc) { plan_free(p); return rc; }
    *plan = reinterpret_cast<fibers_plan*>(p);
    return 0;
}

void fibers_plan_destroy(fibers_plan* plan) { plan_free(reinterpret_cast<Plan*>(plan)); }

int fibers_plan_matrix(const fibers_plan* plan, float* out, int64_t capacity) {
    const Plan* p = reinterpret_cast<const Plan*>(plan);
    if (!p) return -FIBERS_ERR_ARG;
    if (out) {
        if (capacity < (int64_t)p->h_matrix.size()) return -FIBERS_ERR_ARG;
        memcpy(out, p->h_matrix.data(), sizeof(float) * p->h_matrix.size());
    }
    return p->rows;
}

int fibers_plan_kernel(const fibers_plan* plan) {
    const Plan* p = reinterpret_cast<const Plan*>(plan);
    return p ? p->kernel : -1;
}

int fibers_dti_fit_device(fibers_plan* plan, const float* d_dwi, int64_t dwi_pitch, const uint8_t* d_mask,
                          int64_t nvox, int64_t out_pitch, float* d_s0, float* d_eval1, float* d_eval2,
                          float* d_eval3, float* d_evec1, float* d_evec2, float* d_evec3, float* d_rd,
                          float* d_md, float* d_fa, uint8_t* d_valid, void* stream) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    if (!p || p->kind != PLAN_DTI) return fail(FIBERS_ERR_ARG, "not a DTI plan");
    float* outp[10] = {d_s0, d_eval1, d_eval2, d_eval3, d_evec1, d_evec2, d_evec3, d_rd, d_md, d_fa};
    for (float* o : outp) if (!o) return fail(FIBERS_ERR_ARG, "NULL output pointer");
    if (!d_dwi || !d_mask || dwi_pitch < nvox || out_pitch < nvox) return fail(FIBERS_ERR_ARG, "bad dwi/mask/pitch");
    FB_CUDA(cudaSetDevice(p->device));
    return launch_dti(p, d_dwi, dwi_pitch, d_mask, nvox, out_pitch, outp, d_valid, (cudaStream_t)stream);
}

int fibers_adc_fit_device(fibers_plan* plan, const float* d_dwi, int64_t dwi_pitch, const uint8_t* d_mask,
                          int64_t nvox, float* d_adc, float* d_s0, void* stream) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    if (!p || p->kind != PLAN_ADC) return fail(FIBERS_ERR_ARG, "not an ADC plan");
    if (!d_dwi || !d_mask || !d_adc || !d_s0 || dwi_pitch < nvox) return fail(FIBERS_ERR_ARG, "bad argument");
    FB_CUDA(cudaSetDevice(p->device));
    return launch_adc(p, d_dwi, dwi_pitch, d_mask, nvox, d_adc, d_s0, (cudaStream_t)stream);
}

int fibers_stats_init_device(int32_t* d_stats, void* stream) {
    if (!d_stats) return fail(FIBERS_ERR_ARG, "d_stats is NULL");
    return launch_stats_init(d_stats, (cudaStream_t)stream);
}

int fibers_qa_scale_device(float* q1, float* q2, float* q3, int64_t nvox, const int32_t* d_stats, float odfmax,
                           void* stream) {
    if (!q1 || !q2 || !q3) return fail(FIBERS_ERR_ARG, "NULL qa pointer");
    return launch_qa_scale(q1, q2, q3, nvox, d_stats, odfmax, (cudaStream_t)stream);
}

int fibers_recon_device(fibers_plan* plan, const float* d_dwi, int64_t dwi_pitch, const uint8_t* d_mask,
                        int64_t nvox, int64_t out_pitch, float* d_pdf, float* d_odf, float* d_peak1,
                        float* d_peak2, float* d_peak3, float* d_qa1, float* d_qa2, float* d_qa3,
                        int16_t* d_peak_idx, int32_t* d_stats, int finalize, void* stream) {
    Plan* p = reinterpret_cast<Plan*>(plan);
    if (!p || (p->kind != PLAN_GQI && p->kind != PLAN_DSI)) return fail(FIBERS_ERR_ARG, "not a GQI/DSI plan");
    if (!d_dwi || !d_mask || !d_odf || !d_peak1 || !d_peak2 || !d_peak3 || !d_qa1 || !d_qa2 || !d_qa3 || !d_stats)
        return fail(FIBERS_ERR_ARG, "NULL device pointer");
    if (p->kind == PLAN_DSI && !d_pdf) return fail(FIBERS_ERR_ARG, "DSI needs a pdf output");
    if (dwi_pitch < nvox || out_pitch < nvox) return fail(FIBERS_ERR_ARG, "pitch smaller than nvox");
    FB_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = 0;
    if (finalize && (rc = launch_stats_init(d_stats, st))) return rc;
    ReconArgs a{d_dwi, dwi_pitch, d_mask, nvox, out_pitch, d_pdf, d_odf, {d_peak1, d_peak2, d_peak3},
                {d_qa1, d_qa2, d_qa3}, d_peak_idx, d_stats};
    rc = p->kernel == FIBERS_KERNEL_TC ? launch_recon_tc(p, a, st) : launch_recon_simt(p, a, st);
    if (rc) return rc;
    if (finalize) rc = launch_qa_scale(d_qa1, d_qa2, d_qa3, nvox, d_stats, 0.f, st);
    return rc;
}

}  // extern "C"

extern "C" {

void fibers_cuda_release_cache(void) { release_host_caches(); }

int fibers_host_build_matrix(int kind, int nvol, const float* bval, const float* bvec, const float* vertices,
                             int nvert2, float sigma, int hann_width, float* out, int64_t capacity, int* cvol,
                             float* dscale) {
    std::vector<float> A, M; std::vector<uint8_t> ib0;
    std::string e; int rows = 0; int cv = -1; float ds = 0.f;
    switch (kind) {
        case PLAN_DTI: e = build_dti_design(nvol, bval, bvec, A, M, ib0); rows = 7; break;
        case PLAN_ADC: e = build_adc_design(nvol, bval, A, M, ib0); rows = 2; break;
        case PLAN_GQI: e = build_gqi_matrix(nvol, bval, bvec, vertices, nvert2, sigma, M); rows = nvert2 / 2; break;
        case PLAN_DSI: e = build_dsi_matrix(nvol, bval, bvec, vertices, nvert2, hann_width, M, cv, ds);
                       rows = nvert2 / 2 + nvol; break;
        default: fail(FIBERS_ERR_ARG, "unknown matrix kind"); return -FIBERS_ERR_ARG;
    }
    if (!e.empty()) return -to_rc(e);
    if (capacity < (int64_t)M.size() || !out) { fail(FIBERS_ERR_ARG, "output buffer too small"); return -FIBERS_ERR_ARG; }
    memcpy(out, M.data(), sizeof(float) * M.size());
    if (cvol) *cvol = cv;
    if (dscale) *dscale = ds;
    return rows;
}

int fibers_host_build_neighbours(const int32_t* faces, int nface, int nvert, uint16_t* out) {
    std::vector<uint16_t> nbr;
    std::string e = build_neighbours(faces, nface, nvert, nbr);
    if (!e.empty()) return to_rc(e);
    memcpy(out, nbr.data(), sizeof(uint16_t) * nbr.size());
    return 0;
}

int fibers_host_partition_slabs(const uint8_t* mask, int64_t nxny, int nz, int ngpu, int64_t* out) {
    if (!mask || !out || nxny <= 0 || nz <= 0 || ngpu < 1) return fail(FIBERS_ERR_ARG, "bad argument");
    ngpu = std::min(ngpu, nz);
    std::vector<Shard> sh = partition_slabs(mask, nxny, nz, ngpu);
    for (int g = 0; g < ngpu; ++g) { out[2 * g] = sh[g].v0; out[2 * g + 1] = sh[g].v1; }
    return 0;
}

int fibers_dti_fit(const float* dwi, const uint8_t* mask, int nx, int ny, int nz, int nvol, const float* bval,
                   const float* bvec, float* s0, float* eval1, float* eval2, float* eval3, float* evec1,
                   float* evec2, float* evec3, float* rd, float* md, float* fa, uint8_t* valid, int ngpu) {
    if (!bval || nvol <= 0) return fail(FIBERS_ERR_TABLE, "Missing b-value table from input DWI structure");
    if (!bvec) return fail(FIBERS_ERR_TABLE, "Missing gradient table from input DWI structure");
    if (nx <= 0 || ny <= 0 || nz <= 0) return fail(FIBERS_ERR_ARG, "volume dimensions must be positive");
    float* outs[10] = {s0, eval1, eval2, eval3, evec1, evec2, evec3, rd, md, fa};
    const int nfr[10] = {1, 1, 1, 1, 3, 3, 3, 1, 1, 1};
    HostJob job;
    job.kind = PLAN_DTI; job.nvol = nvol; job.dtype = FIBERS_F32;
    job.nxny = (int64_t)nx * ny; job.nz = nz; job.nvox = job.nxny * nz;
    job.dwi = dwi; job.mask = mask; job.valid = valid;
    for (int i = 0; i < 10; ++i) {
        if (!outs[i]) return fail(FIBERS_ERR_ARG, "NULL output pointer");
        job.out_f32.push_back({outs[i], nfr[i]});
    }
    job.make_plan = [=](Plan** p, int dev) {
        return fibers_dti_plan_create(reinterpret_cast<fibers_plan**>(p), dev, nvol, bval, bvec);
    };
    job.plan_key = fnv1a(fnv1a(fnv1a(1469598103934665603ull, "dti", 3), bval, sizeof(float) * nvol), bvec, sizeof(float) * 3 * nvol);
    return run_host_jobs({job}, ngpu);
}

int fibers_adc_fit(const float* dwi, const uint8_t* mask, int nx, int ny, int nz, int nvol, const float* bval,
                   float* adc, float* s0, int ngpu) {
    if (!bval || nvol <= 0) return fail(FIBERS_ERR_TABLE, "Missing b-value table from input DWI structure");
    if (nx <= 0 || ny <= 0 || nz <= 0) return fail(FIBERS_ERR_ARG, "volume dimensions must be positive");
    if (!adc || !s0) return fail(FIBERS_ERR_ARG, "NULL output pointer");
    HostJob job;
    job.kind = PLAN_ADC; job.nvol = nvol; job.dtype = FIBERS_F32;
    job.nxny = (int64_t)nx * ny; job.nz = nz; job.nvox = job.nxny * nz;
    job.dwi = dwi; job.mask = mask;
    job.out_f32 = {{adc, 1}, {s0, 1}};
    job.make_plan = [=](Plan** p, int dev) {
        return fibers_adc_plan_create(reinterpret_cast<fibers_plan**>(p), dev, nvol, bval);
    };
    job.plan_key = fnv1a(fnv1a(1469598103934665603ull, "adc", 3), bval, sizeof(float) * nvol);
    return run_host_jobs({job}, ngpu);
}

// Builds one subject's GQI / DSI request (optionally with the companion DTI fit of the fused entry point).
// odf may be NULL: the ODF is still formed on the device (it feeds the peak search) but is not copied back.
static int recon_job(HostJob& job, int kind, const void* dwi, int dwi_dtype, const uint8_t* mask, int nx, int ny, int nz, int nvol,
                     const float* bval, const float* bvec, const float* vertices, int nvert2, const int32_t* faces,
                     int nface, float sigma, int hann_width, float* pdf, float* odf, float* peak1, float* peak2,
                     float* peak3, float* qa1, float* qa2, float* qa3, int16_t* peak_idx, float* const* dti_out = nullptr) {
    if (!bval || nvol <= 0) return fail(FIBERS_ERR_TABLE, "Missing b-value table from input DWI structure");
    if (!bvec) return fail(FIBERS_ERR_TABLE, "Missing gradient table from input DWI structure");
    if (nx <= 0 || ny <= 0 || nz <= 0) return fail(FIBERS_ERR_ARG, "volume dimensions must be positive");
    if (dwi_dtype < FIBERS_F32 || dwi_dtype > FIBERS_U8) return fail(FIBERS_ERR_ARG, "unsupported dwi element type");
    if (!peak1 || !peak2 || !peak3 || !qa1 || !qa2 || !qa3 || (kind == PLAN_DSI && !pdf))
        return fail(FIBERS_ERR_ARG, "NULL output pointer");
    job.kind = kind; job.nvol = nvol; job.dtype = dwi_dtype;
    job.nxny = (int64_t)nx * ny; job.nz = nz; job.nvox = job.nxny * nz;
    job.dwi = dwi; job.mask = mask; job.peak_idx = peak_idx;
    if (kind == PLAN_DSI) job.out_f32.push_back({pdf, nvol});
    job.out_f32.push_back({odf, nvert2 / 2});
    job.out_f32.push_back({peak1, 3}); job.out_f32.push_back({peak2, 3}); job.out_f32.push_back({peak3, 3});
    job.qa[0] = qa1; job.qa[1] = qa2; job.qa[2] = qa3;
    job.make_plan = [=](Plan** p, int dev) {
        return kind == PLAN_GQI
            ? fibers_gqi_plan_create(reinterpret_cast<fibers_plan**>(p), dev, nvol, bval, bvec, vertices, nvert2, faces, nface, sigma)
            : fibers_dsi_plan_create(reinterpret_cast<fibers_plan**>(p), dev, nvol, bval, bvec, vertices, nvert2, faces, nface, hann_width);
    };
    if (vertices && faces && nvert2 > 0 && nface > 0) {
        uint64_t h = fnv1a(1469598103934665603ull, kind == PLAN_GQI ? "gqi" : "dsi", 3);
        h = fnv1a(h, bval, sizeof(float) * nvol); h = fnv1a(h, bvec, sizeof(float) * 3 * nvol);
        h = fnv1a(h, vertices, sizeof(float) * 3 * (size_t)nvert2); h = fnv1a(h, faces, sizeof(int32_t) * 3 * (size_t)nface);
        h = fnv1a(h, &sigma, sizeof(sigma)); h = fnv1a(h, &hann_width, sizeof(hann_width));
        const int kc = kernel_choice(); h = fnv1a(h, &kc, sizeof(kc));
        job.plan_key = h;
    }
    if (dti_out) {
        const int nfr[10] = {1, 1, 1, 1, 3, 3, 3, 1, 1, 1};
        for (int i = 0; i < 10; ++i) {
            if (!dti_out[i]) return fail(FIBERS_ERR_ARG, "NULL output pointer");
            job.out2_f32.push_back({dti_out[i], nfr[i]});
        }
        job.make_plan2 = [=](Plan** p, int dev) {
            return fibers_dti_plan_create(reinterpret_cast<fibers_plan**>(p), dev, nvol, bval, bvec);
        };
        job.plan2_key = fnv1a(fnv1a(fnv1a(1469598103934665603ull, "dti", 3), bval, sizeof(float) * nvol), bvec, sizeof(float) * 3 * nvol);
    }
    return 0;
}

int fibers_gqi_rec(const void* dwi, int dwi_dtype, const uint8_t* mask, int nx, int ny, int nz, int nvol,
                   const float* bval, const float* bvec, const float* vertices, int nvert2, const int32_t* faces,
                   int nface, float sigma, float* odf, float* peak1, float* peak2, float* peak3, float* qa1,
                   float* qa2, float* qa3, int16_t* peak_idx, int ngpu) {
    HostJob job;
    if (int rc = recon_job(job, PLAN_GQI, dwi, dwi_dtype, mask, nx, ny, nz, nvol, bval, bvec, vertices, nvert2, faces, nface,
                           sigma, 0, nullptr, odf, peak1, peak2, peak3, qa1, qa2, qa3, peak_idx)) return rc;
    return run_host_jobs({job}, ngpu);
}

int fibers_dti_gqi_fit(const float* dwi, const uint8_t* mask, int nx, int ny, int nz, int nvol, const float* bval,
                       const float* bvec, float* s0, float* eval1, float* eval2, float* eval3, float* evec1,
                       float* evec2, float* evec3, float* rd, float* md, float* fa, const float* vertices,
                       int nvert2, const int32_t* faces, int nface, float sigma, float* odf, float* peak1,
                       float* peak2, float* peak3, float* qa1, float* qa2, float* qa3, int ngpu) {
    float* const dti_out[10] = {s0, eval1, eval2, eval3, evec1, evec2, evec3, rd, md, fa};
    HostJob job;
    if (int rc = recon_job(job, PLAN_GQI, dwi, FIBERS_F32, mask, nx, ny, nz, nvol, bval, bvec, vertices, nvert2, faces, nface,
                           sigma, 0, nullptr, odf, peak1, peak2, peak3, qa1, qa2, qa3, nullptr, dti_out)) return rc;
    return run_host_jobs({job}, ngpu);
}

int fibers_dti_gqi_fit_batch(int nsub, const float* const* dwi, const uint8_t* const* mask, int nx, int ny, int nz, int nvol,
                             const float* bval, const float* bvec, float* const* dti_out, const float* vertices,
                             int nvert2, const int32_t* faces, int nface, float sigma, float* const* gqi_out, int ngpu) {
    if (nsub < 0 || (nsub > 0 && (!dwi || !mask || !gqi_out))) return fail(FIBERS_ERR_ARG, "NULL subject table");
    std::vector<HostJob> jobs((size_t)nsub);
    for (int i = 0; i < nsub; ++i) {
        float* const* g = gqi_out + (size_t)i * 7;
        if (int rc = recon_job(jobs[i], PLAN_GQI, dwi[i], FIBERS_F32, mask[i], nx, ny, nz, nvol, bval, bvec, vertices, nvert2, faces,
                               nface, sigma, 0, nullptr, g[0], g[1], g[2], g[3], g[4], g[5], g[6], nullptr,
                               dti_out ? dti_out + (size_t)i * 10 : nullptr)) return rc;
    }
    return run_host_jobs(jobs, ngpu);
}

int fibers_dsi_rec(const void* dwi, int dwi_dtype, const uint8_t* mask, int nx, int ny, int nz, int nvol,
                   const float* bval, const float* bvec, const float* vertices, int nvert2, const int32_t* faces,
                   int nface, int hann_width, float* pdf, float* odf, float* peak1, float* peak2, float* peak3,
                   float* qa1, float* qa2, float* qa3, int16_t* peak_idx, int ngpu) {
    HostJob job;
    if (int rc = recon_job(job, PLAN_DSI, dwi, dwi_dtype, mask, nx, ny, nz, nvol, bval, bvec, vertices, nvert2, faces, nface,
                           0.f, hann_width, pdf, odf, peak1, peak2, peak3, qa1, qa2, qa3, peak_idx)) return rc;
    return run_host_jobs({job}, ngpu);
}

int fibers_cuda_host_register(void* ptr, size_t bytes) {
    if (!ptr || !bytes) return fail(FIBERS_ERR_ARG, "NULL pointer / zero size");
    if (device_count_raw() <= 0) return fail(FIBERS_ERR_NODEV, "no CUDA device available (libfibers_cuda has no CPU fallback)");
    FB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return 0;
}

int fibers_cuda_host_unregister(void* ptr) {
    if (!ptr) return fail(FIBERS_ERR_ARG, "NULL pointer");
    FB_CUDA(cudaHostUnregister(ptr));
    return 0;
}

}  // extern "C"
