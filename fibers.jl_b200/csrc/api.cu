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

namespace fibers {

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};
static std::mutex g_cfg_mu;
static std::vector<int> g_devices;
static bool g_devices_init = false;
static int g_kernel = FIBERS_KERNEL_AUTO;

void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) { g_err = msg; return code; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int device_count_raw() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static std::vector<int> device_list() {
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

static void plan_free(Plan* p) {
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

// ============================================================================================
// Host-pointer entry points: z-slab partitioner + per-GPU pipelines
// ============================================================================================
namespace fibers {

struct Shard { int64_t v0, v1; };   // voxel range [v0, v1), aligned to z-slab boundaries

// Split nz slices into ngpu contiguous z-slabs balanced by masked-voxel count.
static std::vector<Shard> partition_slabs(const uint8_t* mask, int64_t nxny, int nz, int ngpu) {
    std::vector<int64_t> cnt(nz);
    int64_t total = 0;
    for (int z = 0; z < nz; ++z) {
        int64_t c = 0;
        const uint8_t* m = mask + (int64_t)z * nxny;
        for (int64_t i = 0; i < nxny; ++i) c += m[i] != 0;
        cnt[z] = c + 1;           // +1: empty slices still cost a little
        total += cnt[z];
    }
    std::vector<Shard> out;
    int z0 = 0; int64_t acc = 0;
    for (int g = 0; g < ngpu; ++g) {
        int64_t target = total * (g + 1) / ngpu;
        int z1 = z0;
        while (z1 < nz && (acc + cnt[z1] <= target || z1 == z0) && (nz - z1) > (ngpu - 1 - g)) { acc += cnt[z1]; ++z1; }
        if (g == ngpu - 1) z1 = nz;
        out.push_back({(int64_t)z0 * nxny, (int64_t)z1 * nxny});
        z0 = z1;
    }
    return out;
}

struct Frames {            // one host array with `nframes` frames of nvox elements each
    void* host; int nframes; int elem; bool input;
    char* dev[3];          // per pipeline slot
};

struct HostJob {
    int kind;
    int nvol, dtype;
    int64_t nvox, nxny; int nz;
    const void* dwi; const uint8_t* mask;
    std::function<int(Plan**, int)> make_plan;
    uint64_t plan_key = 0;                          // hash of everything the plan depends on (context cache)
    // outputs (host)
    std::vector<std::pair<void*, int>> out_f32;     // (ptr, nframes) float outputs in kernel order
    float* qa[3] = {nullptr, nullptr, nullptr};
    int16_t* peak_idx = nullptr; uint8_t* valid = nullptr;
    // optional companion DTI fit on the same resident slab (fibers_dti_gqi_fit): one H2D of the DWI feeds both
    std::function<int(Plan**, int)> make_plan2;
    uint64_t plan2_key = 0;
    std::vector<std::pair<void*, int>> out2_f32;    // the 10 DTI outputs in kernel order
};

struct Rendezvous {       // cross-shard reduction of odfmax (host side; no device collective)
    std::mutex mu; std::condition_variable cv;
    int arrived = 0, n = 0; float maxv = -INFINITY; bool failed = false;
    float wait_max(float mine, bool ok) {
        std::unique_lock<std::mutex> lk(mu);
        if (!ok) failed = true;
        maxv = std::max(maxv, mine);
        if (++arrived == n) cv.notify_all();
        else cv.wait(lk, [&] { return arrived == n; });
        return maxv;
    }
};

// Per-device context cache: streams, slab ring, QA scratch and the last plan survive between host
// calls (a batch of subjects with one protocol pays for set-up once).  One call at a time may own a
// device's cache; concurrent calls on the same device fall back to private allocations.
struct DeviceCache {
    std::mutex mu; bool busy = false;
    cudaStream_t st[3] = {nullptr, nullptr, nullptr};
    char* slab[3] = {nullptr, nullptr, nullptr}; size_t slab_bytes = 0;
    float* qa = nullptr; size_t qa_bytes = 0; int32_t* stats = nullptr;
    Plan* plan = nullptr; uint64_t plan_key = 0;
    Plan* plan2 = nullptr; uint64_t plan2_key = 0;     // companion DTI plan of the fused entry point
};
static DeviceCache g_cache[64];

static uint64_t fnv1a(uint64_t h, const void* data, size_t n) {
    const unsigned char* p = (const unsigned char*)data;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

#define W_CUDA(expr)                                                                          \
    do { cudaError_t _e = (expr); if (_e != cudaSuccess) {                                    \
        err = std::string(#expr) + ": " + cudaGetErrorString(_e);                             \
        code = _e == cudaErrorMemoryAllocation ? FIBERS_ERR_NOMEM : FIBERS_ERR_CUDA; goto done; } } while (0)

static void shard_worker(const HostJob& job, int device, Shard sh, Rendezvous* rv, int* out_code, std::string* out_err) {
    int code = 0; std::string err;
    Plan* plan = nullptr; Plan* plan2 = nullptr;
    const bool fused = (bool)job.make_plan2;
    constexpr int NSLOT = 3;
    cudaStream_t st[NSLOT] = {nullptr, nullptr, nullptr};
    char* slab[NSLOT] = {nullptr, nullptr, nullptr};
    float* d_qa_all = nullptr; int32_t* d_stats = nullptr;
    DeviceCache* dc = nullptr;
    if (device >= 0 && device < 64) {
        std::lock_guard<std::mutex> lk(g_cache[device].mu);
        if (!g_cache[device].busy) { g_cache[device].busy = true; dc = &g_cache[device]; }
    }
    const int64_t n = sh.v1 - sh.v0;
    const bool recon = job.kind == PLAN_GQI || job.kind == PLAN_DSI;
    bool reached_rv = false;
    const int esz = job.dtype == FIBERS_F32 ? 4 : job.dtype == FIBERS_F64 ? 8 : job.dtype == FIBERS_I32 ? 4
                  : job.dtype == FIBERS_U8 ? 1 : 2;
    {
        W_CUDA(cudaSetDevice(device));
        if (n > 0) {
            if (dc && dc->plan && dc->plan_key == job.plan_key && job.plan_key != 0) plan = dc->plan;
            else {
                if (dc && dc->plan) { plan_free(dc->plan); dc->plan = nullptr; }
                code = job.make_plan(&plan, device);
                if (code) { err = g_err; goto done; }
                if (dc) { dc->plan = plan; dc->plan_key = job.plan_key; }
            }
            if (fused) {
                if (dc && dc->plan2 && dc->plan2_key == job.plan2_key && job.plan2_key != 0) plan2 = dc->plan2;
                else {
                    if (dc && dc->plan2) { plan_free(dc->plan2); dc->plan2 = nullptr; }
                    code = job.make_plan2(&plan2, device);
                    if (code) { err = g_err; goto done; }
                    if (dc) { dc->plan2 = plan2; dc->plan2_key = job.plan2_key; }
                }
            }
            // per-voxel device bytes of one pipeline slot
            int out_frames = 0;
            for (auto& o : job.out_f32) out_frames += o.second;
            for (auto& o : job.out2_f32) out_frames += o.second;
            const int64_t per_vox = (int64_t)job.nvol * 4 + (job.dtype != FIBERS_F32 ? (int64_t)job.nvol * esz : 0)
                                  + 1 + (int64_t)out_frames * 4 + 6 + 1;
            size_t free_b = 0, total_b = 0;
            W_CUDA(cudaMemGetInfo(&free_b, &total_b));
            int64_t chunk_max = 1 << 18;
            if (const char* cv = getenv("FIBERS_CUDA_CHUNK_VOXELS")) { long v = atol(cv); if (v >= 4096) chunk_max = v; }
            int64_t chunk = std::min<int64_t>(n, chunk_max);
            while (chunk > 4096 && (double)chunk * per_vox * NSLOT > 0.6 * (double)free_b) chunk /= 2;
            chunk = (chunk + 63) / 64 * 64;
            const int64_t cp = chunk;                  // device pitch (elements) inside a slot
            const size_t slab_need = (size_t)(per_vox * cp + 1024);
            if (dc) {
                for (int s = 0; s < NSLOT; ++s) if (!dc->st[s]) W_CUDA(cudaStreamCreateWithFlags(&dc->st[s], cudaStreamNonBlocking));
                if (dc->slab_bytes < slab_need) {
                    for (int s = 0; s < NSLOT; ++s) { if (dc->slab[s]) cudaFree(dc->slab[s]); dc->slab[s] = nullptr; }
                    dc->slab_bytes = 0;
                    for (int s = 0; s < NSLOT; ++s) W_CUDA(cudaMalloc(&dc->slab[s], slab_need));
                    dc->slab_bytes = slab_need;
                }
                for (int s = 0; s < NSLOT; ++s) { st[s] = dc->st[s]; slab[s] = dc->slab[s]; }
            } else {
                for (int s = 0; s < NSLOT; ++s) {
                    W_CUDA(cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking));
                    W_CUDA(cudaMalloc(&slab[s], slab_need));
                }
            }
            if (recon) {
                const size_t qa_need = sizeof(float) * 3 * (size_t)n;
                if (dc) {
                    if (dc->qa_bytes < qa_need) { if (dc->qa) cudaFree(dc->qa); dc->qa = nullptr; dc->qa_bytes = 0;
                                                   W_CUDA(cudaMalloc(&dc->qa, qa_need)); dc->qa_bytes = qa_need; }
                    if (!dc->stats) W_CUDA(cudaMalloc(&dc->stats, 2 * sizeof(int32_t)));
                    d_qa_all = dc->qa; d_stats = dc->stats;
                } else {
                    W_CUDA(cudaMalloc(&d_qa_all, qa_need));
                    W_CUDA(cudaMalloc(&d_stats, 2 * sizeof(int32_t)));
                }
                code = launch_stats_init(d_stats, st[0]);
                if (code) { err = g_err; goto done; }
                W_CUDA(cudaStreamSynchronize(st[0]));
            }
            int ci = 0;
            for (int64_t c0 = 0; c0 < n; c0 += chunk, ++ci) {
                const int s = ci % NSLOT;
                const int64_t cn = std::min(chunk, n - c0);
                const int64_t g0 = sh.v0 + c0;          // global voxel offset
                char* base = slab[s];
                // slot layout: [dwi f32 nvol*cp][raw (non-f32 input) nvol*cp*esz][outputs...][idx i16 3cp][mask cp][valid cp]
                float* d_dwi = (float*)base; base += sizeof(float) * job.nvol * cp;
                char* d_raw = nullptr;
                if (job.dtype != FIBERS_F32) { d_raw = base; base += (size_t)esz * job.nvol * cp; base = (char*)(((uintptr_t)base + 255) & ~(uintptr_t)255); }
                std::vector<float*> d_out;
                for (auto& o : job.out_f32) { d_out.push_back((float*)base); base += sizeof(float) * o.second * cp; }
                std::vector<float*> d_out2;
                for (auto& o : job.out2_f32) { d_out2.push_back((float*)base); base += sizeof(float) * o.second * cp; }
                int16_t* d_idx = (int16_t*)base; base += 6 * cp;
                uint8_t* d_mask = (uint8_t*)base; base += cp;
                uint8_t* d_valid = (uint8_t*)base;
                // H2D: every volume contributes one contiguous run of cn voxels (pitch = full volume)
                if (job.dtype == FIBERS_F32)
                    W_CUDA(cudaMemcpy2DAsync(d_dwi, cp * 4, (const char*)job.dwi + g0 * 4, job.nvox * 4, cn * 4, job.nvol,
                                             cudaMemcpyHostToDevice, st[s]));
                else {
                    W_CUDA(cudaMemcpy2DAsync(d_raw, cp * esz, (const char*)job.dwi + g0 * esz, job.nvox * esz, cn * esz,
                                             job.nvol, cudaMemcpyHostToDevice, st[s]));
                    code = launch_convert(d_raw, job.dtype, d_dwi, (int64_t)job.nvol * cp, st[s]);
                    if (code) { err = g_err; goto done; }
                }
                W_CUDA(cudaMemcpyAsync(d_mask, job.mask + g0, cn, cudaMemcpyHostToDevice, st[s]));
                if (job.kind == PLAN_DTI) {
                    code = launch_dti(plan, d_dwi, cp, d_mask, cn, cp, d_out.data(), job.valid ? d_valid : nullptr, st[s]);
                } else if (job.kind == PLAN_ADC) {
                    code = launch_adc(plan, d_dwi, cp, d_mask, cn, d_out[0], d_out[1], st[s]);
                } else {
                    ReconArgs a{};
                    a.dwi = d_dwi; a.dwi_pitch = cp; a.mask = d_mask; a.nvox = cn; a.out_pitch = cp;
                    int oi = 0;
                    if (job.kind == PLAN_DSI) a.pdf = d_out[oi++];
                    a.odf = d_out[oi++];
                    for (int k = 0; k < 3; ++k) a.peak[k] = d_out[oi++];
                    for (int k = 0; k < 3; ++k) a.qa[k] = d_qa_all + (size_t)k * n + c0;   // QA stays on device until odfmax is known
                    a.peak_idx = job.peak_idx ? d_idx : nullptr; a.stats = d_stats;
                    code = plan->kernel == FIBERS_KERNEL_TC ? launch_recon_tc(plan, a, st[s]) : launch_recon_simt(plan, a, st[s]);
                }
                if (code) { err = g_err; goto done; }
                if (fused) {                              // companion DTI fit on the slab that is already resident
                    code = launch_dti(plan2, d_dwi, cp, d_mask, cn, cp, d_out2.data(), nullptr, st[s]);
                    if (code) { err = g_err; goto done; }
                    for (size_t i = 0; i < job.out2_f32.size(); ++i)
                        W_CUDA(cudaMemcpy2DAsync((char*)job.out2_f32[i].first + g0 * 4, job.nvox * 4, d_out2[i], cp * 4, cn * 4,
                                                 job.out2_f32[i].second, cudaMemcpyDeviceToHost, st[s]));
                }
                // D2H gathers into the caller's arrays at the slab offset
                for (size_t i = 0; i < job.out_f32.size(); ++i)
                    W_CUDA(cudaMemcpy2DAsync((char*)job.out_f32[i].first + g0 * 4, job.nvox * 4, d_out[i], cp * 4, cn * 4,
                                             job.out_f32[i].second, cudaMemcpyDeviceToHost, st[s]));
                if (job.peak_idx)
                    W_CUDA(cudaMemcpy2DAsync((char*)job.peak_idx + g0 * 2, job.nvox * 2, d_idx, cp * 2, cn * 2, 3,
                                             cudaMemcpyDeviceToHost, st[s]));
                if (job.valid)
                    W_CUDA(cudaMemcpyAsync(job.valid + g0, d_valid, cn, cudaMemcpyDeviceToHost, st[s]));
            }
            for (int s = 0; s < NSLOT; ++s) W_CUDA(cudaStreamSynchronize(st[s]));
        }
        if (recon) {
            // the one cross-slab datum: odfmax = max over ALL voxels of mean(odf) (src/gqi.jl:164)
            float mine = -INFINITY;
            if (n > 0) {
                int32_t h[2];
                W_CUDA(cudaMemcpy(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost));
                mine = ord2f(h[0]);
            }
            reached_rv = true;
            float odfmax = rv->wait_max(mine, true);
            if (n > 0) {
                code = launch_qa_scale(d_qa_all, d_qa_all + n, d_qa_all + 2 * n, n, nullptr, odfmax, st[0]);
                if (code) { err = g_err; goto done; }
                for (int k = 0; k < 3; ++k)
                    W_CUDA(cudaMemcpyAsync(job.qa[k] + sh.v0, d_qa_all + (size_t)k * n, sizeof(float) * n,
                                           cudaMemcpyDeviceToHost, st[0]));
                W_CUDA(cudaStreamSynchronize(st[0]));
            }
        }
    }
done:
    if (recon && !reached_rv) rv->wait_max(-INFINITY, false);
    if (dc) {                                            // everything stays in the cache for the next call
        if (code) for (int s = 0; s < NSLOT; ++s) if (dc->st[s]) cudaStreamSynchronize(dc->st[s]);
        std::lock_guard<std::mutex> lk(dc->mu);
        dc->busy = false;
    } else {
        for (int s = 0; s < NSLOT; ++s) { if (slab[s]) cudaFree(slab[s]); if (st[s]) cudaStreamDestroy(st[s]); }
        if (d_qa_all) cudaFree(d_qa_all);
        if (d_stats) cudaFree(d_stats);
        if (plan) plan_free(plan);
        if (plan2) plan_free(plan2);
    }
    *out_code = code; *out_err = err;
}

static int run_host_job(const HostJob& job, int ngpu) {
    if (job.nvox <= 0) return fail(FIBERS_ERR_ARG, "empty volume");
    if (!job.dwi || !job.mask) return fail(FIBERS_ERR_ARG, "dwi / mask pointer is NULL");
    std::vector<int> devs = device_list();
    if (devs.empty()) return fail(FIBERS_ERR_NODEV, "no CUDA device available (libfibers_cuda has no CPU fallback)");
    if (ngpu < 1) return fail(FIBERS_ERR_ARG, "ngpu must be >= 1");
    ngpu = std::min<int>(ngpu, (int)devs.size());
    ngpu = std::min<int>(ngpu, job.nz);
    std::vector<Shard> shards = partition_slabs(job.mask, job.nxny, job.nz, ngpu);
    Rendezvous rv; rv.n = ngpu;
    std::vector<int> codes(ngpu, 0); std::vector<std::string> errs(ngpu);
    if (ngpu == 1) shard_worker(job, devs[0], shards[0], &rv, &codes[0], &errs[0]);
    else {
        std::vector<std::thread> th;
        for (int g = 0; g < ngpu; ++g)
            th.emplace_back(shard_worker, std::cref(job), devs[g], shards[g], &rv, &codes[g], &errs[g]);
        for (auto& t : th) t.join();
    }
    for (int g = 0; g < ngpu; ++g) if (codes[g]) return fail(codes[g], errs[g]);
    return 0;
}

}  // namespace fibers

extern "C" {

void fibers_cuda_release_cache(void) {
    for (int d = 0; d < 64; ++d) {
        DeviceCache& c = g_cache[d];
        std::lock_guard<std::mutex> lk(c.mu);
        if (c.busy) continue;
        bool any = c.plan || c.plan2 || c.qa || c.stats || c.slab[0] || c.st[0];
        if (!any) continue;
        int cur = 0; cudaGetDevice(&cur); cudaSetDevice(d);
        for (int s = 0; s < 3; ++s) { if (c.slab[s]) cudaFree(c.slab[s]); if (c.st[s]) cudaStreamDestroy(c.st[s]); c.slab[s] = nullptr; c.st[s] = nullptr; }
        if (c.qa) cudaFree(c.qa); if (c.stats) cudaFree(c.stats);
        if (c.plan) plan_free(c.plan);
        if (c.plan2) plan_free(c.plan2);
        c.qa = nullptr; c.stats = nullptr; c.plan = nullptr; c.plan2 = nullptr; c.slab_bytes = c.qa_bytes = 0; c.plan_key = c.plan2_key = 0;
        cudaSetDevice(cur);
    }
}

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
    return run_host_job(job, ngpu);
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
    return run_host_job(job, ngpu);
}

static int recon_host(int kind, const void* dwi, int dwi_dtype, const uint8_t* mask, int nx, int ny, int nz, int nvol,
                      const float* bval, const float* bvec, const float* vertices, int nvert2, const int32_t* faces,
                      int nface, float sigma, int hann_width, float* pdf, float* odf, float* peak1, float* peak2,
                      float* peak3, float* qa1, float* qa2, float* qa3, int16_t* peak_idx, int ngpu,
                      float* const* dti_out = nullptr) {
    if (!bval || nvol <= 0) return fail(FIBERS_ERR_TABLE, "Missing b-value table from input DWI structure");
    if (!bvec) return fail(FIBERS_ERR_TABLE, "Missing gradient table from input DWI structure");
    if (nx <= 0 || ny <= 0 || nz <= 0) return fail(FIBERS_ERR_ARG, "volume dimensions must be positive");
    if (dwi_dtype < FIBERS_F32 || dwi_dtype > FIBERS_U8) return fail(FIBERS_ERR_ARG, "unsupported dwi element type");
    if (!odf || !peak1 || !peak2 || !peak3 || !qa1 || !qa2 || !qa3 || (kind == PLAN_DSI && !pdf))
        return fail(FIBERS_ERR_ARG, "NULL output pointer");
    HostJob job;
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
    return run_host_job(job, ngpu);
}

int fibers_gqi_rec(const void* dwi, int dwi_dtype, const uint8_t* mask, int nx, int ny, int nz, int nvol,
                   const float* bval, const float* bvec, const float* vertices, int nvert2, const int32_t* faces,
                   int nface, float sigma, float* odf, float* peak1, float* peak2, float* peak3, float* qa1,
                   float* qa2, float* qa3, int16_t* peak_idx, int ngpu) {
    return recon_host(PLAN_GQI, dwi, dwi_dtype, mask, nx, ny, nz, nvol, bval, bvec, vertices, nvert2, faces, nface,
                      sigma, 0, nullptr, odf, peak1, peak2, peak3, qa1, qa2, qa3, peak_idx, ngpu);
}

int fibers_dti_gqi_fit(const float* dwi, const uint8_t* mask, int nx, int ny, int nz, int nvol, const float* bval,
                       const float* bvec, float* s0, float* eval1, float* eval2, float* eval3, float* evec1,
                       float* evec2, float* evec3, float* rd, float* md, float* fa, const float* vertices,
                       int nvert2, const int32_t* faces, int nface, float sigma, float* odf, float* peak1,
                       float* peak2, float* peak3, float* qa1, float* qa2, float* qa3, int ngpu) {
    float* const dti_out[10] = {s0, eval1, eval2, eval3, evec1, evec2, evec3, rd, md, fa};
    return recon_host(PLAN_GQI, dwi, FIBERS_F32, mask, nx, ny, nz, nvol, bval, bvec, vertices, nvert2, faces, nface,
                      sigma, 0, nullptr, odf, peak1, peak2, peak3, qa1, qa2, qa3, nullptr, ngpu, dti_out);
}

int fibers_dsi_rec(const void* dwi, int dwi_dtype, const uint8_t* mask, int nx, int ny, int nz, int nvol,
                   const float* bval, const float* bvec, const float* vertices, int nvert2, const int32_t* faces,
                   int nface, int hann_width, float* pdf, float* odf, float* peak1, float* peak2, float* peak3,
                   float* qa1, float* qa2, float* qa3, int16_t* peak_idx, int ngpu) {
    return recon_host(PLAN_DSI, dwi, dwi_dtype, mask, nx, ny, nz, nvol, bval, bvec, vertices, nvert2, faces, nface,
                      0.f, hann_width, pdf, odf, peak1, peak2, peak3, qa1, qa2, qa3, peak_idx, ngpu);
}

}  // extern "C"
