// Host-pointer side of libfibers_cuda: the z-slab partitioner, the subjects x slabs work queue and the per-GPU
// transfer pipeline behind fibers_dti_fit / fibers_adc_fit / fibers_gqi_rec / fibers_dsi_rec / fibers_dti_gqi_fit and
// their batch variant (reference: the `Threads.@threads for iz` voxel nests of src/dti.jl:258, src/gqi.jl:132,
// src/dsi.jl:197 and the serial odfmax post-pass src/gqi.jl:164-168).
//
//   * one host worker thread per GPU, a 3-slot ring of device slabs + streams; a slot carries one z-slab chunk
//     (<= 2^18 voxels): H2D -> kernels -> D2H, chunks of different slots overlap;
//   * caller arrays in PINNED / registered host memory are copied directly (pitched 2-D DMA);
//   * caller arrays in PAGEABLE memory (what a Julia `Array` is, src/mri.jl:249-255) go through a pinned bounce
//     ring: a small pool of host threads gathers the chunk's rows into a staging buffer laid out like the device
//     slab (one contiguous DMA each way) and scatters the results back while the GPU works on the next chunks;
//   * the worker and its copy threads are bound to the CPUs local to the GPU's PCIe root
//     (/sys/bus/pci/devices/<id>/local_cpulist) before the staging buffers are allocated (first touch = local node);
//   * QA planes stay on the device until odfmax is known; for a subject that lives on one GPU the divisor is
//     decoded on the device, so nothing forces a host sync between subjects: the ring keeps rolling across subject
//     boundaries (batch mode).  Subjects split over several GPUs reduce the scalar on the host (no collective).
#include <emmintrin.h>
#include <sched.h>
#include <pthread.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <cctype>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include "host_pipeline.h"

namespace fibers {

const std::string& last_error();            // api.cu

// Split nz slices into ngpu contiguous z-slabs balanced by masked-voxel count.
std::vector<Shard> partition_slabs(const uint8_t* mask, int64_t nxny, int nz, int ngpu) {
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

namespace {

// ---- cross-slab reduction of odfmax for a subject that is split over several GPUs (host side) ----
struct Rendezvous {
    std::mutex mu; std::condition_variable cv;
    int arrived = 0, n = 0; float maxv = -INFINITY; bool failed = false;
    void arrive(float mine, bool ok) {
        std::lock_guard<std::mutex> lk(mu);
        if (!ok) failed = true;
        maxv = std::max(maxv, mine);
        if (++arrived >= n) cv.notify_all();
    }
    float wait(bool* ok) {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return arrived >= n; });
        *ok = !failed;
        return maxv;
    }
};

// ---- host copy threads ------------------------------------------------------------------------
class CopyPool {
public:
    CopyPool(int nthreads, const cpu_set_t* aff) {
        for (int i = 0; i < nthreads; ++i) {
            th_.emplace_back([this] { loop(); });
            if (aff) pthread_setaffinity_np(th_.back().native_handle(), sizeof(cpu_set_t), aff);
        }
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    // fn(i) for i in [0, n) on the pool and the calling thread; returns when every index is done
    void parallel_for(int n, const std::function<void(int)>& fn) {
        if (n <= 0) return;
        if (th_.empty() || n == 1) { for (int i = 0; i < n; ++i) fn(i); return; }
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; n_ = n; next_.store(0); active_ = (int)th_.size(); ++gen_;
        }
        cv_.notify_all();
        for (int i; (i = next_.fetch_add(1)) < n;) fn(i);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return active_ == 0; });
        fn_ = nullptr;
    }
private:
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)>* fn; int n;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_; fn = fn_; n = n_;
            }
            for (int i; (i = next_.fetch_add(1)) < n;) (*fn)(i);
            std::lock_guard<std::mutex> lk(mu_);
            if (--active_ == 0) done_.notify_all();
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_; std::condition_variable cv_, done_;
    const std::function<void(int)>* fn_ = nullptr; int n_ = 0; std::atomic<int> next_{0};
    int active_ = 0; uint64_t gen_ = 0; bool stop_ = false;
};

// Row copy with non-temporal stores: the rows are 0.25 - 1 MB and written once, so read-for-ownership traffic and cache
// pollution are pure cost (glibc's memcpy switches to streaming stores only well above these sizes).
void stream_copy(void* dst, const void* src, size_t n) {
    char* d = (char*)dst; const char* s = (const char*)src;
    if (n < 4096) { memcpy(d, s, n); return; }
    const size_t head = (16 - ((uintptr_t)d & 15)) & 15;
    memcpy(d, s, head); d += head; s += head; n -= head;
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(s + i)), b = _mm_loadu_si128((const __m128i*)(s + i + 16));
        const __m128i c = _mm_loadu_si128((const __m128i*)(s + i + 32)), e = _mm_loadu_si128((const __m128i*)(s + i + 48));
        _mm_stream_si128((__m128i*)(d + i), a); _mm_stream_si128((__m128i*)(d + i + 16), b);
        _mm_stream_si128((__m128i*)(d + i + 32), c); _mm_stream_si128((__m128i*)(d + i + 48), e);
    }
    _mm_sfence();
    memcpy(d + i, s + i, n - i);
}

// ---- CPU affinity: the CPUs local to the GPU's PCIe root ------------------------------------------
bool gpu_local_cpus(int device, cpu_set_t* out) {
    const char* env = getenv("FIBERS_CUDA_AFFINITY");
    if (env && !strcmp(env, "0")) return false;
    char bus[64] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return false; }
    for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return false;
    char line[4096] = {0};
    const bool got = fgets(line, sizeof(line), f) != nullptr;
    fclose(f);
    if (!got) return false;
    cpu_set_t local; CPU_ZERO(&local);
    for (const char* s = line; *s && *s != '\n';) {
        char* e; long a = strtol(s, &e, 10); if (e == s) break;
        long b = a; s = e;
        if (*s == '-') { b = strtol(s + 1, &e, 10); s = e; }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) CPU_SET((int)c, &local);
        if (*s == ',') ++s;
    }
    cpu_set_t cur; CPU_ZERO(&cur);
    if (sched_getaffinity(0, sizeof(cur), &cur) != 0) return false;
    cpu_set_t both; CPU_AND(&both, &local, &cur);
    if (CPU_COUNT(&both) == 0 || CPU_EQUAL(&both, &cur)) return false;      // nothing to narrow
    *out = both;
    return true;
}

bool host_is_pinned(const void* p) {
    if (!p) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

enum HostPath { PATH_AUTO = 0, PATH_DIRECT = 1, PATH_BOUNCE = 2 };
int host_path_choice() {
    const char* e = getenv("FIBERS_CUDA_HOST_PATH");
    if (e && !strcmp(e, "direct")) return PATH_DIRECT;
    if (e && !strcmp(e, "bounce")) return PATH_BOUNCE;
    return PATH_AUTO;
}

int elem_size(int dtype) {
    return dtype == FIBERS_F32 ? 4 : dtype == FIBERS_F64 ? 8 : dtype == FIBERS_I32 ? 4 : dtype == FIBERS_U8 ? 1 : 2;
}

// ---- per-device context: streams, slab ring, staging, QA scratch and the last plans ----------------
constexpr int NSLOT = 3;
constexpr int NQA = 2;

struct SlotLayout {          // byte offsets inside one device slab (and, from out0 on, inside one host staging slot)
    size_t dwi = 0, raw = 0, mask = 0, out0 = 0, idx = 0, valid = 0, end = 0;
    std::vector<size_t> out, out2;
};

SlotLayout make_layout(const HostJob& job, int64_t cp) {
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    SlotLayout L; size_t o = 0;
    L.dwi = o; o = al(o + sizeof(float) * job.nvol * cp);
    L.raw = o; if (job.dtype != FIBERS_F32) o = al(o + (size_t)elem_size(job.dtype) * job.nvol * cp);
    L.mask = o; o = al(o + cp);
    L.out0 = o;
    for (auto& a : job.out_f32) { L.out.push_back(o); o = al(o + sizeof(float) * a.second * cp); }
    for (auto& a : job.out2_f32) { L.out2.push_back(o); o = al(o + sizeof(float) * a.second * cp); }
    L.idx = o; o = al(o + 6 * cp);
    L.valid = o; o = al(o + cp);
    L.end = o;
    return L;
}

struct PendingScatter { bool active = false; const HostJob* job = nullptr; int64_t g0 = 0, cn = 0, cp = 0; SlotLayout L; };
struct PendingFinal { bool active = false; bool staged = false; const HostJob* job = nullptr; int64_t v0 = 0, n = 0; };

struct DeviceCtx {
    std::mutex mu; bool busy = false;
    int device = -1;
    cudaStream_t st[NSLOT] = {}, fin = nullptr;
    cudaEvent_t ev_h2d[NSLOT] = {}, ev_d2h[NSLOT] = {}, ev_tail[NSLOT] = {}, ev_init = nullptr, ev_fin[NQA] = {};
    char* slab[NSLOT] = {}; size_t slab_bytes = 0;
    char* h_in[NSLOT] = {}; size_t h_in_bytes = 0;
    char* h_out[NSLOT] = {}; size_t h_out_bytes = 0;
    float* qa[NQA] = {}; size_t qa_bytes = 0; int32_t* stats[NQA] = {};
    float* h_qa[NQA] = {}; size_t h_qa_bytes = 0;
    Plan* plan = nullptr; uint64_t plan_key = 0;
    Plan* plan2 = nullptr; uint64_t plan2_key = 0;
    CopyPool* pool = nullptr;
    uint64_t nchunk = 0; int nqa = 0;
    bool h2d_used[NSLOT] = {};
    PendingScatter pend[NSLOT]; PendingFinal pfin[NQA];
};
DeviceCtx g_ctx[64];

struct Err { int code = 0; std::string msg; };

#define P_CUDA(expr)                                                                          \
    do { cudaError_t _e = (expr); if (_e != cudaSuccess) {                                    \
        e.msg = std::string(#expr) + ": " + cudaGetErrorString(_e);                           \
        e.code = _e == cudaErrorMemoryAllocation ? FIBERS_ERR_NOMEM : FIBERS_ERR_CUDA; return e.code; } } while (0)
#define P_CALL(expr) do { int _rc = (expr); if (_rc) { e.code = _rc; e.msg = last_error(); return _rc; } } while (0)

int ctx_init(DeviceCtx& c, int device, Err& e) {
    c.device = device;
    for (int s = 0; s < NSLOT; ++s) {
        if (!c.st[s]) P_CUDA(cudaStreamCreateWithFlags(&c.st[s], cudaStreamNonBlocking));
        if (!c.ev_h2d[s]) P_CUDA(cudaEventCreateWithFlags(&c.ev_h2d[s], cudaEventDisableTiming));
        if (!c.ev_d2h[s]) P_CUDA(cudaEventCreateWithFlags(&c.ev_d2h[s], cudaEventDisableTiming));
        if (!c.ev_tail[s]) P_CUDA(cudaEventCreateWithFlags(&c.ev_tail[s], cudaEventDisableTiming));
    }
    if (!c.fin) P_CUDA(cudaStreamCreateWithFlags(&c.fin, cudaStreamNonBlocking));
    if (!c.ev_init) P_CUDA(cudaEventCreateWithFlags(&c.ev_init, cudaEventDisableTiming));
    for (int q = 0; q < NQA; ++q) {
        if (!c.ev_fin[q]) P_CUDA(cudaEventCreateWithFlags(&c.ev_fin[q], cudaEventDisableTiming));
        if (!c.stats[q]) P_CUDA(cudaMalloc(&c.stats[q], 2 * sizeof(int32_t)));
    }
    return 0;
}

void ctx_free(DeviceCtx& c) {           // device must be current
    for (int s = 0; s < NSLOT; ++s) {
        if (c.slab[s]) cudaFree(c.slab[s]);
        if (c.h_in[s]) cudaFreeHost(c.h_in[s]);
        if (c.h_out[s]) cudaFreeHost(c.h_out[s]);
        if (c.st[s]) cudaStreamDestroy(c.st[s]);
        if (c.ev_h2d[s]) cudaEventDestroy(c.ev_h2d[s]);
        if (c.ev_d2h[s]) cudaEventDestroy(c.ev_d2h[s]);
        if (c.ev_tail[s]) cudaEventDestroy(c.ev_tail[s]);
        c.slab[s] = c.h_in[s] = c.h_out[s] = nullptr; c.st[s] = nullptr; c.ev_h2d[s] = c.ev_d2h[s] = c.ev_tail[s] = nullptr;
        c.pend[s] = PendingScatter(); c.h2d_used[s] = false;
    }
    for (int q = 0; q < NQA; ++q) {
        if (c.qa[q]) cudaFree(c.qa[q]);
        if (c.stats[q]) cudaFree(c.stats[q]);
        if (c.h_qa[q]) cudaFreeHost(c.h_qa[q]);
        if (c.ev_fin[q]) cudaEventDestroy(c.ev_fin[q]);
        c.qa[q] = nullptr; c.stats[q] = nullptr; c.h_qa[q] = nullptr; c.ev_fin[q] = nullptr; c.pfin[q] = PendingFinal();
    }
    if (c.fin) cudaStreamDestroy(c.fin);
    if (c.ev_init) cudaEventDestroy(c.ev_init);
    c.fin = nullptr; c.ev_init = nullptr;
    if (c.plan) plan_free(c.plan);
    if (c.plan2) plan_free(c.plan2);
    c.plan = c.plan2 = nullptr; c.plan_key = c.plan2_key = 0;
    c.slab_bytes = c.h_in_bytes = c.h_out_bytes = c.qa_bytes = c.h_qa_bytes = 0;
    delete c.pool; c.pool = nullptr;
}

// rows of one chunk: caller's [frames][nvox] arrays <-> staging [frames][cp]
void scatter_chunk(DeviceCtx& c, int s, const PendingScatter& p) {
    struct Row { char* dst; const char* src; size_t bytes; };
    std::vector<Row> rows;
    const char* base = c.h_out[s] - p.L.out0;                      // staging holds the slab's output region
    auto add = [&](const std::vector<std::pair<void*, int>>& arrs, const std::vector<size_t>& offs) {
        for (size_t i = 0; i < arrs.size(); ++i) {
            if (!arrs[i].first) continue;
            for (int f = 0; f < arrs[i].second; ++f)
                rows.push_back({(char*)arrs[i].first + ((int64_t)f * p.job->nvox + p.g0) * 4, base + offs[i] + (size_t)f * p.cp * 4, (size_t)p.cn * 4});
        }
    };
    add(p.job->out_f32, p.L.out); add(p.job->out2_f32, p.L.out2);
    if (p.job->peak_idx)
        for (int f = 0; f < 3; ++f)
            rows.push_back({(char*)p.job->peak_idx + ((int64_t)f * p.job->nvox + p.g0) * 2, base + p.L.idx + (size_t)f * p.cp * 2, (size_t)p.cn * 2});
    if (p.job->valid) rows.push_back({(char*)p.job->valid + p.g0, base + p.L.valid, (size_t)p.cn});
    c.pool->parallel_for((int)rows.size(), [&](int i) { stream_copy(rows[i].dst, rows[i].src, rows[i].bytes); });
}

int complete_scatter(DeviceCtx& c, int s, Err& e) {
    if (!c.pend[s].active) return 0;
    P_CUDA(cudaEventSynchronize(c.ev_d2h[s]));
    scatter_chunk(c, s, c.pend[s]);
    c.pend[s].active = false;
    return 0;
}

int complete_final(DeviceCtx& c, int q, Err& e) {
    PendingFinal& f = c.pfin[q];
    if (!f.active) return 0;
    P_CUDA(cudaEventSynchronize(c.ev_fin[q]));
    if (f.staged)
        c.pool->parallel_for(3, [&](int k) { memcpy(f.job->qa[k] + f.v0, c.h_qa[q] + (size_t)k * f.n, sizeof(float) * f.n); });
    f.active = false;
    return 0;
}

// everything queued on this device has landed in the caller's arrays
int ctx_drain(DeviceCtx& c, Err& e) {
    for (int s = 0; s < NSLOT; ++s) if (complete_scatter(c, s, e)) return e.code;
    for (int q = 0; q < NQA; ++q) if (complete_final(c, q, e)) return e.code;
    for (int s = 0; s < NSLOT; ++s) if (c.st[s]) P_CUDA(cudaStreamSynchronize(c.st[s]));
    if (c.fin) P_CUDA(cudaStreamSynchronize(c.fin));
    return 0;
}

void ctx_abort(DeviceCtx& c) {          // after an error: nothing may still be writing into buffers we are about to reuse / free
    for (int s = 0; s < NSLOT; ++s) { if (c.st[s]) cudaStreamSynchronize(c.st[s]); c.pend[s].active = false; }
    if (c.fin) cudaStreamSynchronize(c.fin);
    for (int q = 0; q < NQA; ++q) c.pfin[q].active = false;
    cudaGetLastError();
}

template <typename T>
int grow_host(T** p, size_t* have, size_t need, int n, Err& e) {       // n pinned buffers of equal size
    if (*have >= need) return 0;
    for (int i = 0; i < n; ++i) { if (p[i]) cudaFreeHost(p[i]); p[i] = nullptr; }
    *have = 0;
    for (int i = 0; i < n; ++i) P_CUDA(cudaHostAlloc((void**)&p[i], need, cudaHostAllocDefault));
    *have = need;
    return 0;
}

struct Unit { const HostJob* job; Shard sh; Rendezvous* rv; };

// One unit = one subject's voxel range [v0, v1) on this device.  Returns once everything is QUEUED (the results land
// later: ctx_drain); `arrived` tells the caller whether the unit's rendezvous (if any) has been served.
int run_unit(DeviceCtx& c, const Unit& u, bool* arrived, Err& e) {
    const HostJob& job = *u.job;
    const int64_t n = u.sh.v1 - u.sh.v0;
    const bool recon = job.kind == PLAN_GQI || job.kind == PLAN_DSI;
    const bool fused = (bool)job.make_plan2;
    const int esz = elem_size(job.dtype);
    if (n <= 0) {
        if (u.rv) { u.rv->arrive(-INFINITY, true); *arrived = true; bool ok; u.rv->wait(&ok); }
        return 0;
    }
    // ---- plans (cached per device while the protocol stays the same) ----
    if (!(c.plan && c.plan_key == job.plan_key && job.plan_key != 0)) {
        if (ctx_drain(c, e)) return e.code;
        if (c.plan) { plan_free(c.plan); c.plan = nullptr; }
        P_CALL(job.make_plan(&c.plan, c.device));
        c.plan_key = job.plan_key;
    }
    if (fused && !(c.plan2 && c.plan2_key == job.plan2_key && job.plan2_key != 0)) {
        if (ctx_drain(c, e)) return e.code;
        if (c.plan2) { plan_free(c.plan2); c.plan2 = nullptr; }
        P_CALL(job.make_plan2(&c.plan2, c.device));
        c.plan2_key = job.plan2_key;
    }
    Plan* plan = c.plan; Plan* plan2 = fused ? c.plan2 : nullptr;
    // ---- transfer mode ----
    const int hp = host_path_choice();
    bool in_direct = hp == PATH_DIRECT || (hp == PATH_AUTO && host_is_pinned(job.dwi));
    bool out_direct = hp == PATH_DIRECT;
    if (hp == PATH_AUTO) {
        out_direct = true;
        for (auto& a : job.out_f32) out_direct = out_direct && host_is_pinned(a.first);
        for (auto& a : job.out2_f32) out_direct = out_direct && host_is_pinned(a.first);
        for (int k = 0; k < 3; ++k) out_direct = out_direct && host_is_pinned(job.qa[k]);
    }
    // ---- chunk size, slab ring, staging ----
    size_t free_b = 0, total_b = 0;
    P_CUDA(cudaMemGetInfo(&free_b, &total_b));
    int64_t chunk_max = (in_direct && out_direct) ? (1 << 18) : (1 << 17);
    if (const char* cv = getenv("FIBERS_CUDA_CHUNK_VOXELS")) { long v = atol(cv); if (v >= 4096) chunk_max = v; }
    int64_t chunk = std::min<int64_t>(n, chunk_max);
    chunk = (chunk + 63) / 64 * 64;
    while (chunk > 4096 && (double)make_layout(job, chunk).end * NSLOT > 0.6 * (double)(free_b + c.slab_bytes * NSLOT)) chunk = (chunk / 2 + 63) / 64 * 64;
    const int64_t cp = chunk;                  // device pitch (elements) inside a slot
    const SlotLayout L = make_layout(job, cp);
    if (c.slab_bytes < L.end) {
        if (ctx_drain(c, e)) return e.code;
        for (int s = 0; s < NSLOT; ++s) { if (c.slab[s]) cudaFree(c.slab[s]); c.slab[s] = nullptr; }
        c.slab_bytes = 0;
        for (int s = 0; s < NSLOT; ++s) P_CUDA(cudaMalloc(&c.slab[s], L.end));
        c.slab_bytes = L.end;
    }
    const size_t in_need = (size_t)esz * job.nvol * cp + cp;
    const size_t out_need = L.end - L.out0;
    if ((!in_direct && c.h_in_bytes < in_need) || (!out_direct && c.h_out_bytes < out_need)) {
        if (ctx_drain(c, e)) return e.code;
        if (!in_direct && grow_host(c.h_in, &c.h_in_bytes, in_need, NSLOT, e)) return e.code;
        if (!out_direct && grow_host(c.h_out, &c.h_out_bytes, out_need, NSLOT, e)) return e.code;
    }
    if (!c.pool) {
        int nt = 16;
        if (const char* tv = getenv("FIBERS_CUDA_COPY_THREADS")) nt = std::max(1, atoi(tv));
        cpu_set_t cur; CPU_ZERO(&cur);
        const bool have = sched_getaffinity(0, sizeof(cur), &cur) == 0;       // the worker is already bound (run_worker)
        if (have) nt = std::min(nt, std::max(1, CPU_COUNT(&cur)));
        c.pool = new CopyPool(nt - 1, have ? &cur : nullptr);
    }
    // ---- QA scratch of this unit ----
    int q = 0;
    float* d_qa = nullptr; int32_t* d_stats = nullptr;
    if (recon) {
        q = c.nqa++ % NQA;
        if (complete_final(c, q, e)) return e.code;
        const size_t qa_need = sizeof(float) * 3 * (size_t)n;
        if (c.qa_bytes < qa_need) {
            if (ctx_drain(c, e)) return e.code;
            for (int i = 0; i < NQA; ++i) { if (c.qa[i]) cudaFree(c.qa[i]); c.qa[i] = nullptr; }
            c.qa_bytes = 0;
            for (int i = 0; i < NQA; ++i) P_CUDA(cudaMalloc(&c.qa[i], qa_need));
            c.qa_bytes = qa_need;
        }
        if (!out_direct && grow_host(c.h_qa, &c.h_qa_bytes, qa_need, NQA, e)) return e.code;
        d_qa = c.qa[q]; d_stats = c.stats[q];
        P_CALL(launch_stats_init(d_stats, c.fin));
        P_CUDA(cudaEventRecord(c.ev_init, c.fin));
    }
    // ---- chunks ----
    bool used[NSLOT] = {false, false, false};
    // Ramp-up: the first chunks of a unit are small (chunk / 8, / 4, / 2, then full size), so that the D2H engine -- the busier
    // direction: 4.87 GB out against 4.22 GB in for a cfg2 subject -- starts after 1 ms instead of after the H2D + kernel of a full
    // 300 MB chunk (7 ms of a 115 ms call).  FIBERS_CUDA_RAMP=0 disables.
    int64_t cur = chunk;
    { const char* rv = getenv("FIBERS_CUDA_RAMP"); if (!rv || atoi(rv) != 0) cur = std::min<int64_t>(chunk, std::max<int64_t>(4096, (chunk / 8 + 63) / 64 * 64)); }
    int64_t cn_step = 0;
    for (int64_t c0 = 0; c0 < n; c0 += cn_step) {
        const int s = (int)(c.nchunk++ % NSLOT);
        const int64_t cn = std::min(cur, n - c0);
        cn_step = cn; cur = std::min(chunk, cur * 2);
        const int64_t g0 = u.sh.v0 + c0;          // global voxel offset
        cudaStream_t st = c.st[s];
        char* base = c.slab[s];
        float* d_dwi = (float*)(base + L.dwi);
        char* d_raw = base + L.raw;
        uint8_t* d_mask = (uint8_t*)(base + L.mask);
        std::vector<float*> d_out, d_out2;
        for (size_t i = 0; i < L.out.size(); ++i) d_out.push_back((float*)(base + L.out[i]));
        for (size_t i = 0; i < L.out2.size(); ++i) d_out2.push_back((float*)(base + L.out2[i]));
        int16_t* d_idx = (int16_t*)(base + L.idx);
        uint8_t* d_valid = (uint8_t*)(base + L.valid);
        if (complete_scatter(c, s, e)) return e.code;                    // the slot's previous results have left the staging buffer
        if (recon && !used[s]) P_CUDA(cudaStreamWaitEvent(st, c.ev_init, 0));
        used[s] = true;
        // -- H2D: every volume contributes one contiguous run of cn voxels
        char* d_in = job.dtype == FIBERS_F32 ? (char*)d_dwi : d_raw;
        if (in_direct) {
            P_CUDA(cudaMemcpy2DAsync(d_in, cp * esz, (const char*)job.dwi + g0 * esz, job.nvox * esz, cn * esz, job.nvol,
                                     cudaMemcpyHostToDevice, st));
            P_CUDA(cudaMemcpyAsync(d_mask, job.mask + g0, cn, cudaMemcpyHostToDevice, st));
        } else {
            if (c.h2d_used[s]) P_CUDA(cudaEventSynchronize(c.ev_h2d[s]));   // the DMA engine is done with this staging slot
            char* hin = c.h_in[s];
            const char* src = (const char*)job.dwi;
            const int64_t nvox = job.nvox;
            c.pool->parallel_for(job.nvol + 1, [&](int k) {
                if (k < job.nvol) stream_copy(hin + (size_t)k * cp * esz, src + ((int64_t)k * nvox + g0) * esz, (size_t)cn * esz);
                else memcpy(hin + (size_t)job.nvol * cp * esz, job.mask + g0, (size_t)cn);
            });
            P_CUDA(cudaMemcpyAsync(d_in, hin, (size_t)job.nvol * cp * esz, cudaMemcpyHostToDevice, st));
            P_CUDA(cudaMemcpyAsync(d_mask, hin + (size_t)job.nvol * cp * esz, cn, cudaMemcpyHostToDevice, st));
            P_CUDA(cudaEventRecord(c.ev_h2d[s], st));
            c.h2d_used[s] = true;
        }
        if (job.dtype != FIBERS_F32) P_CALL(launch_convert(d_raw, job.dtype, d_dwi, (int64_t)job.nvol * cp, st));
        // -- kernels
        if (job.kind == PLAN_DTI) {
            P_CALL(launch_dti(plan, d_dwi, cp, d_mask, cn, cp, d_out.data(), job.valid ? d_valid : nullptr, st));
        } else if (job.kind == PLAN_ADC) {
            P_CALL(launch_adc(plan, d_dwi, cp, d_mask, cn, d_out[0], d_out[1], st));
        } else {
            ReconArgs a{};
            a.dwi = d_dwi; a.dwi_pitch = cp; a.mask = d_mask; a.nvox = cn; a.out_pitch = cp;
            int oi = 0;
            if (job.kind == PLAN_DSI) a.pdf = d_out[oi++];
            a.odf = d_out[oi++];
            for (int k = 0; k < 3; ++k) a.peak[k] = d_out[oi++];
            for (int k = 0; k < 3; ++k) a.qa[k] = d_qa + (size_t)k * n + c0;   // QA stays on device until odfmax is known
            a.peak_idx = job.peak_idx ? d_idx : nullptr; a.stats = d_stats;
            P_CALL(plan->kernel == FIBERS_KERNEL_TC ? launch_recon_tc(plan, a, st) : launch_recon_simt(plan, a, st));
        }
        if (fused) P_CALL(launch_dti(plan2, d_dwi, cp, d_mask, cn, cp, d_out2.data(), nullptr, st));
        // -- D2H
        if (out_direct) {
            for (size_t i = 0; i < job.out_f32.size(); ++i)
                if (job.out_f32[i].first)
                    P_CUDA(cudaMemcpy2DAsync((char*)job.out_f32[i].first + g0 * 4, job.nvox * 4, d_out[i], cp * 4, cn * 4,
                                             job.out_f32[i].second, cudaMemcpyDeviceToHost, st));
            for (size_t i = 0; i < job.out2_f32.size(); ++i)
                if (job.out2_f32[i].first)
                    P_CUDA(cudaMemcpy2DAsync((char*)job.out2_f32[i].first + g0 * 4, job.nvox * 4, d_out2[i], cp * 4, cn * 4,
                                             job.out2_f32[i].second, cudaMemcpyDeviceToHost, st));
            if (job.peak_idx)
                P_CUDA(cudaMemcpy2DAsync((char*)job.peak_idx + g0 * 2, job.nvox * 2, d_idx, cp * 2, cn * 2, 3, cudaMemcpyDeviceToHost, st));
            if (job.valid) P_CUDA(cudaMemcpyAsync(job.valid + g0, d_valid, cn, cudaMemcpyDeviceToHost, st));
        } else {
            // one contiguous copy of the slab's output region, minus leading arrays nobody wants (odf == NULL)
            size_t from = L.out0;
            for (size_t i = 0; i < job.out_f32.size() && !job.out_f32[i].first; ++i) from = i + 1 < L.out.size() ? L.out[i + 1] : (L.out2.empty() ? L.idx : L.out2[0]);
            size_t to = job.valid ? L.end : (job.peak_idx ? L.valid : L.idx);
            if (to > from) P_CUDA(cudaMemcpyAsync(c.h_out[s] + (from - L.out0), base + from, to - from, cudaMemcpyDeviceToHost, st));
            P_CUDA(cudaEventRecord(c.ev_d2h[s], st));
            PendingScatter& p = c.pend[s];
            p.active = true; p.job = &job; p.g0 = g0; p.cn = cn; p.cp = cp; p.L = L;
        }
    }
    if (!recon) return 0;
    // ---- QA: divide by odfmax = max over ALL voxels of the subject of mean(odf) (src/gqi.jl:164-168) ----
    for (int s = 0; s < NSLOT; ++s)
        if (used[s]) { P_CUDA(cudaEventRecord(c.ev_tail[s], c.st[s])); P_CUDA(cudaStreamWaitEvent(c.fin, c.ev_tail[s], 0)); }
    if (!u.rv) {
        P_CALL(launch_qa_scale(d_qa, d_qa + n, d_qa + 2 * n, n, d_stats, 0.f, c.fin));       // divisor decoded on the device
    } else {
        int32_t h[2];
        P_CUDA(cudaMemcpyAsync(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost, c.fin));
        P_CUDA(cudaStreamSynchronize(c.fin));
        u.rv->arrive(ord2f(h[0]), true); *arrived = true;
        bool ok; const float odfmax = u.rv->wait(&ok);
        if (!ok) { e.code = FIBERS_ERR_CUDA; e.msg = "another z-slab of this subject failed"; return e.code; }
        P_CALL(launch_qa_scale(d_qa, d_qa + n, d_qa + 2 * n, n, nullptr, odfmax, c.fin));
    }
    PendingFinal& f = c.pfin[q];
    f.active = true; f.job = &job; f.v0 = u.sh.v0; f.n = n; f.staged = !out_direct;
    if (out_direct) {
        for (int k = 0; k < 3; ++k)
            P_CUDA(cudaMemcpyAsync(job.qa[k] + u.sh.v0, d_qa + (size_t)k * n, sizeof(float) * n, cudaMemcpyDeviceToHost, c.fin));
    } else {
        P_CUDA(cudaMemcpyAsync(c.h_qa[q], d_qa, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c.fin));
    }
    P_CUDA(cudaEventRecord(c.ev_fin[q], c.fin));
    return 0;
}

struct Queue {
    std::vector<Unit> units; std::atomic<size_t> next{0}; std::atomic<bool> failed{false};
    std::mutex mu; int code = 0; std::string msg;
    void fail(const Err& e) {
        std::lock_guard<std::mutex> lk(mu);
        if (!failed.exchange(true)) { code = e.code ? e.code : FIBERS_ERR_CUDA; msg = e.msg; }
    }
};

void run_worker(Queue* qu, int device) {
    Err e;
    cpu_set_t saved; CPU_ZERO(&saved);
    bool narrowed = false;
    DeviceCtx* c = nullptr; bool cached = false;
    if (cudaSetDevice(device) != cudaSuccess) { e.code = FIBERS_ERR_CUDA; e.msg = "cudaSetDevice failed"; cudaGetLastError(); }
    if (!e.code) {
        cpu_set_t local;
        if (sched_getaffinity(0, sizeof(saved), &saved) == 0 && gpu_local_cpus(device, &local))
            narrowed = pthread_setaffinity_np(pthread_self(), sizeof(local), &local) == 0;
        if (device >= 0 && device < 64) {
            std::lock_guard<std::mutex> lk(g_ctx[device].mu);
            if (!g_ctx[device].busy) { g_ctx[device].busy = true; c = &g_ctx[device]; cached = true; }
        }
        if (!c) c = new DeviceCtx();            // a concurrent call owns this device's cache: private buffers
        ctx_init(*c, device, e);
    }
    for (;;) {
        const size_t i = qu->next.fetch_add(1);
        if (i >= qu->units.size()) break;
        const Unit& u = qu->units[i];
        bool arrived = false;
        if (!e.code && !qu->failed.load()) run_unit(*c, u, &arrived, e);
        if (e.code) qu->fail(e);
        if (u.rv && !arrived) u.rv->arrive(-INFINITY, false);       // never leave the other slabs of the subject waiting
    }
    if (c) {
        if (!e.code && !qu->failed.load()) { if (ctx_drain(*c, e)) qu->fail(e); }
        if (e.code || qu->failed.load()) ctx_abort(*c);
        if (cached) { std::lock_guard<std::mutex> lk(c->mu); c->busy = false; }
        else { ctx_free(*c); delete c; }
    }
    if (narrowed) pthread_setaffinity_np(pthread_self(), sizeof(saved), &saved);
}

}  // namespace

int run_host_jobs(const std::vector<HostJob>& jobs, int ngpu) {
    if (jobs.empty()) return 0;
    for (auto& job : jobs) {
        if (job.nvox <= 0) return fail(FIBERS_ERR_ARG, "empty volume");
        if (!job.dwi || !job.mask) return fail(FIBERS_ERR_ARG, "dwi / mask pointer is NULL");
    }
    std::vector<int> devs = device_list();
    if (devs.empty()) return fail(FIBERS_ERR_NODEV, "no CUDA device available (libfibers_cuda has no CPU fallback)");
    if (ngpu < 1) return fail(FIBERS_ERR_ARG, "ngpu must be >= 1");
    ngpu = std::min<int>(ngpu, (int)devs.size());
    // slabs per subject: 1 when there are at least as many subjects as GPUs, otherwise enough to keep every GPU busy
    const int nsub = (int)jobs.size();
    int nslab = nsub >= ngpu ? 1 : (ngpu + nsub - 1) / nsub;
    Queue qu;
    std::vector<std::unique_ptr<Rendezvous>> rvs;
    for (auto& job : jobs) {
        const int ns = std::min(nslab, job.nz);
        std::vector<Shard> sh = ns > 1 ? partition_slabs(job.mask, job.nxny, job.nz, ns) : std::vector<Shard>{{0, job.nvox}};
        Rendezvous* rv = nullptr;
        if (ns > 1 && (job.kind == PLAN_GQI || job.kind == PLAN_DSI)) { rvs.emplace_back(new Rendezvous()); rv = rvs.back().get(); rv->n = ns; }
        for (auto& s : sh) qu.units.push_back({&job, s, rv});
    }
    const int nworker = std::min<int>(ngpu, (int)qu.units.size());
    if (nworker == 1) run_worker(&qu, devs[0]);
    else {
        std::vector<std::thread> th;
        for (int g = 0; g < nworker; ++g) th.emplace_back(run_worker, &qu, devs[g]);
        for (auto& t : th) t.join();
    }
    if (qu.failed.load()) return fail(qu.code, qu.msg);
    return 0;
}

void release_host_caches() {
    for (int d = 0; d < 64; ++d) {
        DeviceCtx& c = g_ctx[d];
        std::lock_guard<std::mutex> lk(c.mu);
        if (c.busy || c.device < 0) continue;
        int cur = 0; cudaGetDevice(&cur); cudaSetDevice(d);
        ctx_free(c);
        c.device = -1;
        cudaSetDevice(cur);
    }
}

}  // namespace fibers
