// tcgen05 (UMMA) engine — placeholder until the kernels land; reports "unsupported" so that
// NSR_ENGINE_AUTO routes everything to the exact-fp32 engine.
#include "common.cuh"
namespace nsr {
bool conv_fprop_tc_supported(const NsrConv&) { return false; }
int conv_fprop_tc(const NsrConv&, cudaStream_t) { set_error("tcgen05 engine not built"); return NSR_E_INVALID; }
bool conv_wgrad_tc_supported(const NsrWgrad&) { return false; }
size_t conv_wgrad_workspace_tc(const NsrWgrad&) { return 0; }
int conv_wgrad_tc(const NsrWgrad&, cudaStream_t) { set_error("tcgen05 engine not built"); return NSR_E_INVALID; }
}  // namespace nsr
