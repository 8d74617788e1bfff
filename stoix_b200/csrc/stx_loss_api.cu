// Stand-alone faces of the two reference loss utilities (forward value only):
//   ppo_clip_loss       stoix/utils/loss.py:17-32
//   clipped_value_loss  stoix/utils/loss.py:68-78
// The training path never calls these: K3 fuses the same arithmetic with its gradient.  They exist so
// that code written against stoix.utils.loss keeps working and to test the loss arithmetic in
// isolation.  out[0] receives the mean; deterministic two-level reduction.
#include "stx_common.cuh"

namespace stx {
namespace {

template <int KIND>
__global__ void __launch_bounds__(256) loss_value_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                         const float* __restrict__ c, int64_t n, float eps,
                                                         double* __restrict__ partials, unsigned int* counter,
                                                         float* __restrict__ out) {
  __shared__ double sm[32];
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    if (KIND == 0) {  // a = pi_log_prob_t, b = b_pi_log_prob_t, c = gae_t
      const float ratio = expf(a[i] - b[i]);
      const float l1 = ratio * c[i];
      const float l2 = fminf(fmaxf(ratio, 1.0f - eps), 1.0f + eps) * c[i];
      acc += (double)(-fminf(l1, l2));
    } else {  // a = pred_value_t, b = behavior_value_t, c = targets_t
      const float vclip = b[i] + fminf(fmaxf(a[i] - b[i], -eps), eps);
      const float e1 = a[i] - c[i], e2 = vclip - c[i];
      acc += (double)(0.5f * fmaxf(e1 * e1, e2 * e2));
    }
  }
  const double bs = block_sum<double>(acc, sm);
  if (threadIdx.x == 0) partials[blockIdx.x] = bs;
  if (last_block_ticket(counter, gridDim.x)) {
    double s = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += partials[i];
    s = block_sum<double>(s, sm);
    if (threadIdx.x == 0) out[0] = (float)(s / (double)n);
  }
}

template <int KIND>
int launch(const float* a, const float* b, const float* c, int64_t n, float eps, float* out, void* scratch, void* stream) {
  STX_REQUIRE(a && b && c && out && scratch && n > 0, STX_E_ARG, "stx loss: bad arguments");
  int64_t blocks = (n + 255) / 256;
  if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
  unsigned int* counter = reinterpret_cast<unsigned int*>(scratch);
  double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + 16);
  loss_value_kernel<KIND><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, b, c, n, eps, partials, counter, out);
  STX_LAUNCH_OK();
  return STX_OK;
}

}  // namespace
}  // namespace stx

extern "C" size_t stx_loss_scratch_bytes(void) { return 16 + sizeof(double) * 2 * stx::kNumSMs; }

extern "C" int stx_ppo_clip_loss(const float* pi_log_prob_t, const float* b_pi_log_prob_t, const float* gae_t,
                                 int64_t n, float epsilon, float* out, void* scratch, void* stream) {
  return stx::launch<0>(pi_log_prob_t, b_pi_log_prob_t, gae_t, n, epsilon, out, scratch, stream);
}

extern "C" int stx_clipped_value_loss(const float* pred_value_t, const float* behavior_value_t,
                                      const float* targets_t, int64_t n, float epsilon, float* out,
                                      void* scratch, void* stream) {
  return stx::launch<1>(pred_value_t, behavior_value_t, targets_t, n, epsilon, out, scratch, stream);
}
