// tcgen05 bit-plane GEMM (prefill regime) -- placeholder until the kernel lands.
#include "pbllm_common.cuh"
namespace pbl {
bool gemm_tc_supported(const Layer&, const void*, int64_t, const void*, int64_t, int64_t) { return false; }
int launch_gemm_tc(const Layer&, const void*, int64_t, void*, int64_t, int64_t, cudaStream_t) {
    set_error("tcgen05 GEMM path not built");
    return PBL_ERR_UNSUPPORTED;
}
}  // namespace pbl
