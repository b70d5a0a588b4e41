// K3 tensor-core path (tcgen05 + TMEM + TMA).  Placeholder until the kernel lands: reports "unsupported"
// so tc_linear routes to the exact SIMT path.
#include "tc_common.cuh"
namespace tc {
bool linear_tc_supported(const tc_linear_args*) { return false; }
int linear_tc_launch(const tc_linear_args*, cudaStream_t) { set_error("tc_linear: tensor-core path not built"); return TC_ERR_DTYPE; }
}  // namespace tc
