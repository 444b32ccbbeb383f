// Tensor-core (tcgen05) reconstruction path -- placeholder until the UMMA kernel lands.
#include "common.cuh"
namespace fibers {
int tc_plan_init(Plan* p) { (void)p; set_error("tensor-core kernel not built yet"); return 1; }
void tc_plan_free(Plan* p) { (void)p; }
int launch_recon_tc(Plan* p, const ReconArgs& a, cudaStream_t st) {
    (void)p; (void)a; (void)st;
    return fail(FIBERS_ERR_ARG, "tensor-core kernel not built yet");
}
}  // namespace fibers
