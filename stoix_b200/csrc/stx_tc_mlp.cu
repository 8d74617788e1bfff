// bf16 tcgen05 path -- placeholder until the tensor-core kernels land (returns STX_E_UNSUPPORTED).
#include "stx_common.cuh"

namespace stx {
size_t tc_mlp_forward_workspace_bytes(const StxMlp*, int64_t) { return 256; }
int tc_mlp_forward(const StxMlp*, const void*, int64_t, const int32_t*, int64_t, float*, void*, size_t, cudaStream_t) {
  set_error("STX_PREC_BF16 forward is not built into this library");
  return STX_E_UNSUPPORTED;
}
size_t tc_ppo_workspace_bytes(const StxMlp*, const StxMlp*, int64_t) { return 256; }
int tc_ppo_minibatch_grads(const StxMlp*, const StxMlp*, const StxPpoBatch*, int64_t, int64_t, const StxPpoHyper*, float, float*, float*, void*, size_t, cudaStream_t) {
  set_error("STX_PREC_BF16 PPO update is not built into this library");
  return STX_E_UNSUPPORTED;
}
}  // namespace stx
