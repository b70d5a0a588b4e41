// K4 tensor-core path (tcgen05 + TMEM).  Placeholder until the kernel lands.
#include "tc_common.cuh"
namespace tc {
bool attention_tc_supported(const tc_attention_args*) { return false; }
int attention_tc_launch(const tc_attention_args*, cudaStream_t) { set_error("tc_attention_fwd: tensor-core path not built"); return TC_ERR_DTYPE; }
}  // namespace tc
