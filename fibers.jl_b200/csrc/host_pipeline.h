// Host-pointer side of libfibers_cuda: subjects x z-slabs work queue, one worker thread + stream ring per GPU,
// pinned-direct or pageable (bounce ring) transfers.  See host_pipeline.cu.
#pragma once
#include <functional>
#include <string>
#include <utility>
#include <vector>
#include "common.cuh"

namespace fibers {

// One subject's reconstruction request on HOST arrays (Julia column-major [nx,ny,nz,frames]).
struct HostJob {
    int kind = 0;                                   // PLAN_DTI / PLAN_ADC / PLAN_GQI / PLAN_DSI
    int nvol = 0, dtype = FIBERS_F32;
    int64_t nvox = 0, nxny = 0; int nz = 0;
    const void* dwi = nullptr; const uint8_t* mask = nullptr;
    std::function<int(Plan**, int)> make_plan;
    uint64_t plan_key = 0;                          // hash of everything the plan depends on (context cache)
    // float32 outputs in kernel order (host pointer, frames).  A NULL host pointer = "do not copy back"
    // (the device buffer still exists: e.g. the ODF feeds the peak search but the caller only wants peaks / QA).
    std::vector<std::pair<void*, int>> out_f32;
    float* qa[3] = {nullptr, nullptr, nullptr};
    int16_t* peak_idx = nullptr; uint8_t* valid = nullptr;
    // optional companion DTI fit on the same resident slab (fused DTI + GQI): one H2D of the DWI feeds both
    std::function<int(Plan**, int)> make_plan2;
    uint64_t plan2_key = 0;
    std::vector<std::pair<void*, int>> out2_f32;    // the 10 DTI outputs in kernel order
};

struct Shard { int64_t v0, v1; };                   // voxel range [v0, v1), aligned to z-slab boundaries

std::vector<Shard> partition_slabs(const uint8_t* mask, int64_t nxny, int nz, int ngpu);

// Runs every job (subject) on up to `ngpu` devices.  njobs >= ngpu: whole subjects are handed to the devices from an
// ordered queue and the stream ring keeps rolling across subject boundaries (H2D of subject i+1 overlaps the
// kernels and D2H of subject i).  njobs < ngpu: every subject is also split into z-slabs (the one cross-slab datum,
// odfmax, is reduced on the host).  No inter-GPU collective.
int run_host_jobs(const std::vector<HostJob>& jobs, int ngpu);

std::vector<int> device_list();
void release_host_caches();
void plan_free(Plan* p);

}  // namespace fibers
