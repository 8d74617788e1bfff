// Recurrent layer of rec_ppo (fp32): ScannedRNN(cell_type="gru") over a (T, E) sequence with episode resets, forward and
// backward-through-time.
//
// Reference: stoix/networks/base.py:124-159 (ScannedRNN: the carry is replaced by zeros where `resets` is set, THEN the cell
// runs; nn.scan over the leading axis) as used by RecurrentActor / RecurrentCritic (:162-222) in
// stoix/systems/ppo/anakin/rec_ppo.py:90-101 (one step per env step) and :216-247 (whole chunk inside the loss, under jax.grad).
// flax.linen.GRUCell is not vendored in the reference; its published definition is restated:
//     r = sigmoid(gi_r + gh_r)   z = sigmoid(gi_z + gh_z)   n = tanh(gi_n + r * (gh_n + b_hn))   h' = (1 - z) n + z h
// with gi = W_i x + b_i (all three input projections carry a bias), gh = W_h h (only the n part carries a bias).
//
// The input projections of ALL timesteps are one GEMM outside (they are the "head" of the pre-torso MLP); what is sequential is
// h @ W_h + the gate arithmetic, T times.  This first form keeps that chain as 2 launches per step (the small-grid GEMM of
// stx_simt_gemm.cuh + one gate kernel) issued by a host loop INSIDE the entry point, so the ABI is already the one a persistent
// sequence kernel (W_h resident in shared memory, h in registers) would have.  Backward: the gate kernel of step t turns
// d(h_t) into d(gi_t), d(gh_t); d(h_{t-1}) = d(gh_t) W_h^T + d(h_t) z_t (masked by reset_t); d(W_h) = sum_t hp_t^T d(gh_t) is ONE
// GEMM over the stored sequences after the loop (fixed summation order: deterministic).
#include "stx_common.cuh"
#include "stx_simt_gemm.cuh"

namespace stx {
namespace {

inline size_t ralign(size_t x) { return (x + 255) / 256 * 256; }

struct GruWs {
  float *hp_seq;                 // [T + 1][E][H]  state entering step t (after the reset); [T] = where(reset_T.., ..) unused
  float *r, *z, *n, *ghn;        // [T][E][H]
  float *gh;                     // [E][3H]
  float *d_gh_seq;               // [T][E][3H]
  float *dhp_gemm;               // [E][H]
  float *dhp_direct[2];          // [E][H]
  float *partials;               // [splits][H * 3H + 3H]
  int splits;
  size_t bytes;
};

int gru_splits(int64_t rows) {
  int64_t s = rows / 4096;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return (int)s;
}

GruWs carve_gru(int T, int64_t E, int H, char* base) {
  GruWs w{};
  size_t o = 0;
  auto take = [&](size_t floats) {
    float* p = base ? reinterpret_cast<float*>(base + o) : nullptr;
    o += ralign(floats * 4);
    return p;
  };
  const size_t eh = (size_t)E * H;
  w.hp_seq = take((size_t)(T + 1) * eh);
  w.r = take((size_t)T * eh), w.z = take((size_t)T * eh), w.n = take((size_t)T * eh), w.ghn = take((size_t)T * eh);
  w.gh = take(3 * eh);
  w.d_gh_seq = take((size_t)T * 3 * eh);
  w.dhp_gemm = take(eh);
  w.dhp_direct[0] = take(eh), w.dhp_direct[1] = take(eh);
  w.splits = gru_splits((int64_t)T * E);
  w.partials = take((size_t)w.splits * ((size_t)H * 3 * H + 3 * H));
  w.bytes = o;
  return w;
}

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

// hp_0 = where(reset_0, 0, h0)
__global__ void gru_init_kernel(const float* __restrict__ h0, const uint8_t* __restrict__ reset0, int64_t E, int H, float* __restrict__ hp0) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= E * H) return;
  hp0[i] = reset0[i / H] ? 0.f : h0[i];
}

__global__ void gru_gate_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ bhn,
                                    const float* __restrict__ hp, const uint8_t* __restrict__ reset_next, int64_t E, int H,
                                    float* __restrict__ h_out, float* __restrict__ hp_next, float* __restrict__ rs, float* __restrict__ zs,
                                    float* __restrict__ ns, float* __restrict__ ghns) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= E * H) return;
  const int64_t e = i / H;
  const int k = (int)(i % H);
  const float* a = gi + e * 3 * H;
  const float* b = gh + e * 3 * H;
  const float r = sigm(a[k] + b[k]);
  const float z = sigm(a[H + k] + b[H + k]);
  const float ghn = b[2 * H + k] + bhn[k];
  const float n = tanhf(a[2 * H + k] + r * ghn);
  const float hpv = hp[i];
  const float h = (1.f - z) * n + z * hpv;
  h_out[i] = h;
  rs[i] = r, zs[i] = z, ns[i] = n, ghns[i] = ghn;
  if (hp_next) hp_next[i] = (reset_next && reset_next[e]) ? 0.f : h;
}

// d(h_t) (from the layers above + from step t+1) -> d(gi_t), d(gh_t), the direct part of d(hp_t)
__global__ void gru_gate_bwd_kernel(const float* __restrict__ d_h_out, const float* __restrict__ dhp_gemm_next, const float* __restrict__ dhp_direct_next,
                                    const uint8_t* __restrict__ reset_next, const float* __restrict__ rs, const float* __restrict__ zs,
                                    const float* __restrict__ ns, const float* __restrict__ ghns, const float* __restrict__ hp, int64_t E, int H,
                                    float* __restrict__ d_gi, float* __restrict__ d_gh, float* __restrict__ dhp_direct) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= E * H) return;
  const int64_t e = i / H;
  const int k = (int)(i % H);
  float dh = d_h_out ? d_h_out[i] : 0.f;
  if (dhp_gemm_next && !reset_next[e]) dh += dhp_gemm_next[i] + dhp_direct_next[i];   // base.py:139-148: the reset cuts the chain
  const float r = rs[i], z = zs[i], n = ns[i], ghn = ghns[i];
  const float dn = dh * (1.f - z), dz = dh * (hp[i] - n);
  const float dpn = dn * (1.f - n * n);
  const float dpr = dpn * ghn * r * (1.f - r);
  const float dpz = dz * z * (1.f - z);
  float* a = d_gi + e * 3 * H;
  float* b = d_gh + e * 3 * H;
  a[k] = dpr, a[H + k] = dpz, a[2 * H + k] = dpn;
  b[k] = dpr, b[H + k] = dpz, b[2 * H + k] = dpn * r;
  dhp_direct[i] = dh * z;
}

// d(h0) = where(reset_0, 0, d(hp_0))
__global__ void gru_dh0_kernel(const float* __restrict__ dhp_gemm, const float* __restrict__ dhp_direct, const uint8_t* __restrict__ reset0, int64_t E,
                               int H, float* __restrict__ d_h0) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= E * H) return;
  d_h0[i] = reset0[i / H] ? 0.f : dhp_gemm[i] + dhp_direct[i];
}

}  // namespace
}  // namespace stx

using namespace stx;

extern "C" size_t stx_gru_workspace_bytes(int T, int64_t E, int H) {
  if (T <= 0 || E <= 0 || H <= 0) return 0;
  return carve_gru(T, E, H, nullptr).bytes;
}

extern "C" int stx_gru_sequence_forward(const float* gi, const uint8_t* reset, const float* h0, const float* w_h, const float* b_hn, int T, int64_t E,
                                        int H, float* h_seq, void* workspace, size_t workspace_bytes, void* stream) {
  STX_REQUIRE(gi && reset && h0 && w_h && b_hn && h_seq && workspace, STX_E_ARG, "stx_gru_sequence_forward: null pointer");
  STX_REQUIRE(T > 0 && E > 0 && H > 0, STX_E_SHAPE, "stx_gru_sequence_forward: T=%d E=%lld H=%d", T, (long long)E, H);
  STX_REQUIRE(workspace_bytes >= stx_gru_workspace_bytes(T, E, H), STX_E_WORKSPACE, "stx_gru_sequence_forward: workspace %zu < %zu", workspace_bytes,
              stx_gru_workspace_bytes(T, E, H));
  cudaStream_t st = (cudaStream_t)stream;
  GruWs ws = carve_gru(T, E, H, reinterpret_cast<char*>(workspace));
  const size_t eh = (size_t)E * H;
  const unsigned blocks = (unsigned)((eh + 255) / 256);
  gru_init_kernel<<<blocks, 256, 0, st>>>(h0, reset, E, H, ws.hp_seq);
  STX_LAUNCH_OK();
  for (int t = 0; t < T; ++t) {
    simt::GemmArgs g{};
    g.A = ws.hp_seq + (size_t)t * eh, g.lda = H, g.B = w_h, g.C = ws.gh, g.mask_act = -1;
    g.M = E, g.N = 3 * H, g.K = H;
    STX_CUDA_OK(simt::launch_gemm<simt::FWD>(g, 1, st));
    gru_gate_fwd_kernel<<<blocks, 256, 0, st>>>(gi + (size_t)t * 3 * eh, ws.gh, b_hn, ws.hp_seq + (size_t)t * eh,
                                                 t + 1 < T ? reset + (size_t)(t + 1) * E : nullptr, E, H, h_seq + (size_t)t * eh,
                                                 ws.hp_seq + (size_t)(t + 1) * eh, ws.r + (size_t)t * eh, ws.z + (size_t)t * eh, ws.n + (size_t)t * eh,
                                                 ws.ghn + (size_t)t * eh);
    STX_LAUNCH_OK();
  }
  return STX_OK;
}

extern "C" int stx_gru_sequence_backward(const float* d_h_seq, const uint8_t* reset, const float* w_h, int T, int64_t E, int H, void* workspace,
                                         size_t workspace_bytes, float* d_gi, float* d_w_h, float* d_b_hn, float grad_weight, int overwrite,
                                         float* d_h0, void* stream) {
  STX_REQUIRE(d_h_seq && reset && w_h && workspace && d_gi, STX_E_ARG, "stx_gru_sequence_backward: null pointer");
  STX_REQUIRE(T > 0 && E > 0 && H > 0, STX_E_SHAPE, "stx_gru_sequence_backward: T=%d E=%lld H=%d", T, (long long)E, H);
  STX_REQUIRE((d_w_h == nullptr) == (d_b_hn == nullptr), STX_E_ARG, "stx_gru_sequence_backward: d_w_h and d_b_hn go together");
  STX_REQUIRE(workspace_bytes >= stx_gru_workspace_bytes(T, E, H), STX_E_WORKSPACE, "stx_gru_sequence_backward: workspace %zu < %zu", workspace_bytes,
              stx_gru_workspace_bytes(T, E, H));
  cudaStream_t st = (cudaStream_t)stream;
  GruWs ws = carve_gru(T, E, H, reinterpret_cast<char*>(workspace));
  const size_t eh = (size_t)E * H;
  const unsigned blocks = (unsigned)((eh + 255) / 256);
  for (int t = T - 1; t >= 0; --t) {
    const bool last = (t == T - 1);
    float* direct = ws.dhp_direct[t & 1];
    gru_gate_bwd_kernel<<<blocks, 256, 0, st>>>(d_h_seq + (size_t)t * eh, last ? nullptr : ws.dhp_gemm, last ? nullptr : ws.dhp_direct[(t + 1) & 1],
                                                 last ? nullptr : reset + (size_t)(t + 1) * E, ws.r + (size_t)t * eh, ws.z + (size_t)t * eh,
                                                 ws.n + (size_t)t * eh, ws.ghn + (size_t)t * eh, ws.hp_seq + (size_t)t * eh, E, H,
                                                 d_gi + (size_t)t * 3 * eh, ws.d_gh_seq + (size_t)t * 3 * eh, direct);
    STX_LAUNCH_OK();
    if (t > 0 || d_h0) {  // d(hp_t) through W_h: dX form, W_h is (H x 3H) "Kout x Nr"
      simt::GemmArgs d{};
      d.A = ws.d_gh_seq + (size_t)t * 3 * eh, d.lda = 3 * H, d.B = w_h, d.C = ws.dhp_gemm, d.mask_act = -1;
      d.M = E, d.N = H, d.K = 3 * H;
      STX_CUDA_OK(simt::launch_gemm<simt::DX>(d, 1, st));
    }
  }
  if (d_h0) {
    gru_dh0_kernel<<<blocks, 256, 0, st>>>(ws.dhp_gemm, ws.dhp_direct[0], reset, E, H, d_h0);
    STX_LAUNCH_OK();
  }
  if (d_w_h) {
    // d(W_h) = hp_seq^T d_gh_seq over all (t, e) rows; column sums of the n part = d(b_hn)
    const int64_t rows = (int64_t)T * E;
    const int64_t np = (int64_t)H * 3 * H + 3 * H;
    simt::GemmArgs g{};
    g.A = ws.hp_seq, g.lda = H, g.B = ws.d_gh_seq, g.C = ws.partials, g.dbias = ws.partials + (int64_t)H * 3 * H, g.mask_act = -1;
    g.M = rows, g.N = 3 * H, g.K = H;
    g.rows_per_split = (rows + ws.splits - 1) / ws.splits;
    g.part_stride = np, g.dbias_stride = np;
    STX_CUDA_OK(simt::launch_gemm<simt::DW>(g, ws.splits, st));
    const int64_t nw = (int64_t)H * 3 * H;
    simt::reduce_partials_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(ws.partials, ws.splits, np, nw, grad_weight, d_w_h, overwrite);
    STX_LAUNCH_OK();
    simt::reduce_partials_kernel<<<(unsigned)((H + 255) / 256), 256, 0, st>>>(ws.partials + nw + 2 * H, ws.splits, np, H, grad_weight, d_b_hn, overwrite);
    STX_LAUNCH_OK();
  }
  return STX_OK;
}
