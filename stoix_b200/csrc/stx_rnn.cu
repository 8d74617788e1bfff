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
#include <cooperative_groups.h>

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

int gru_splits(int64_t rows) {   // d(W_h) GEMM: (H/64) x (3H/64) output tiles; 512 rows per split fill the GPU from 32k rows on
  int64_t s = rows / 512;
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


// ---- persistent sequence kernels -------------------------------------------------------------------------------------------
// One CTA owns R sequences (rows) for ALL T steps with the whole recurrent matrix W_h (H x 3H fp32, 196 KB at H = 128) resident in
// shared memory, so the recurrence needs no global synchronisation and no per-step launch: per step the CTA streams gi_t in,
// does the (R x H) @ (H x 3H) product out of shared memory (thread n owns gate column n for the R rows; W rows are padded to
// 3H + 1 floats so that both the forward walk (lanes along n) and the transposed walk of the backward (lanes along k) are free
// of bank conflicts), applies the gates (thread j owns hidden unit j) and writes h_t plus what the backward needs.  The step time
// is set by reading W_h once from shared memory (3H * H * 4 B / 128 B per clock ~ 1.5k cycles at H = 128), not by launches.
// Backward: same residency, reverse time; d(gh_t) is kept for the ONE GEMM that forms d(W_h) afterwards.
template <int R, int KS>
struct GruSmem {
  __host__ __device__ static size_t w_floats(int H) { return ((size_t)H * (3 * H + 1) + 3) / 4 * 4; }   // keeps the float4 arrays behind it 16-byte aligned
  static size_t fwd_bytes(int H) { return (w_floats(H) + (size_t)H * R + (size_t)KS * R * 3 * H) * 4; }
  static size_t bwd_bytes(int H) { return (w_floats(H) + (size_t)3 * H * R + 3 * (size_t)KS * R * H + (size_t)R * H) * 4; }
};
constexpr int kGruItems = 3;   // gate items (row, hidden unit) per thread: H * R <= kGruItems * blockDim for every launch shape

__device__ __forceinline__ void gru_load_w(float* __restrict__ Ws, const float* __restrict__ w_h, int H) {
  const int n3 = 3 * H, total = H * n3;
  constexpr int U = 8;
  for (int base = threadIdx.x; base < total; base += blockDim.x * U) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * blockDim.x;
      v[u] = i < total ? __ldg(w_h + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * blockDim.x;
      if (i < total) Ws[(i / n3) * (n3 + 1) + i % n3] = v[u];
    }
  }
}

// KS thread groups split the reduction of the per-step product (partials added in a fixed order); what the gate phase of a step
// needs from global memory (gi_t, reset_{t+1}; backward: the saved gates of step t - 1) is requested BEFORE the product of the
// step, so its latency hides behind the shared-memory walk instead of sitting in front of the gate arithmetic.
template <int R, int KS>
__global__ void __launch_bounds__(1024, 1)
    gru_seq_fwd_kernel(const float* __restrict__ gi, const uint8_t* __restrict__ reset, const float* __restrict__ h0, const float* __restrict__ w_h,
                       const float* __restrict__ b_hn, int T, int64_t E, int H, float* __restrict__ h_seq, float* __restrict__ hp_seq,
                       float* __restrict__ rs, float* __restrict__ zs, float* __restrict__ ns, float* __restrict__ ghns) {
  extern __shared__ __align__(16) float gsm[];
  const int n3 = 3 * H;
  float* Ws = gsm;                               // [H][3H + 1]
  float* hs = Ws + GruSmem<R, KS>::w_floats(H);  // [H][R]   state entering the step, k-major
  float* ghp = hs + (size_t)H * R;               // [KS][R][3H] partial products
  const int kh = threadIdx.x / n3, n = threadIdx.x % n3;   // reduction group, gate column
  const int kper = (H + KS - 1) / KS;
  const int k_begin = kh * kper, k_end = (k_begin + kper < H) ? k_begin + kper : H;
  const int64_t row0 = (int64_t)blockIdx.x * R;
  const int items = H * R;
  gru_load_w(Ws, w_h, H);
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int r = i / H, j = i % H;
    const int64_t row = row0 + r;
    float v = 0.f;
    if (row < E) {
      v = reset[row] ? 0.f : h0[row * H + j];
      hp_seq[row * H + j] = v;
    }
    hs[j * R + r] = v;
  }
  __syncthreads();
  const int64_t eh = E * (int64_t)H;
  for (int t = 0; t < T; ++t) {
    // ---- requests for the gate phase ----
    float g_r[kGruItems], g_z[kGruItems], g_n[kGruItems], bh[kGruItems];
    bool cut[kGruItems];
#pragma unroll
    for (int q = 0; q < kGruItems; ++q) {
      const int i = threadIdx.x + q * blockDim.x;
      g_r[q] = g_z[q] = g_n[q] = bh[q] = 0.f, cut[q] = false;
      if (i < items) {
        const int r = i / H, j = i % H;
        const int64_t row = row0 + r;
        if (row < E) {
          const float* a = gi + ((int64_t)t * E + row) * n3;
          g_r[q] = __ldg(a + j), g_z[q] = __ldg(a + H + j), g_n[q] = __ldg(a + 2 * H + j), bh[q] = __ldg(b_hn + j);
          cut[q] = (t + 1 < T) && reset[(int64_t)(t + 1) * E + row] != 0;
        }
      }
    }
    // ---- (R x H) @ (H x 3H), this group's share of k ----
    if (kh < KS) {
      float2 acc[R / 2];
#pragma unroll
      for (int r = 0; r < R / 2; ++r) acc[r] = make_float2(0.f, 0.f);
      const float* wcol = Ws + n;
#pragma unroll 4
      for (int k = k_begin; k < k_end; ++k) {
        const float w = wcol[(size_t)k * (n3 + 1)];
        const float* hk = hs + k * R;
#pragma unroll
        for (int r4 = 0; r4 < R; r4 += 4) {
          const float4 hv = *reinterpret_cast<const float4*>(hk + r4);
          fma2(acc[r4 / 2], w, make_float2(hv.x, hv.y)), fma2(acc[r4 / 2 + 1], w, make_float2(hv.z, hv.w));
        }
      }
#pragma unroll
      for (int r = 0; r < R / 2; ++r) ghp[((size_t)kh * R + 2 * r) * n3 + n] = acc[r].x, ghp[((size_t)kh * R + 2 * r + 1) * n3 + n] = acc[r].y;
    }
    __syncthreads();
    // ---- gates: consecutive threads -> consecutive hidden units of one row ----
#pragma unroll
    for (int q = 0; q < kGruItems; ++q) {
      const int i = threadIdx.x + q * blockDim.x;
      if (i >= items) continue;
      const int r = i / H, j = i % H;
      const int64_t row = row0 + r;
      if (row >= E) continue;
      float b_r = 0.f, b_z = 0.f, b_n = 0.f;
#pragma unroll
      for (int g = 0; g < KS; ++g) {   // fixed order over the reduction groups
        const float* b = ghp + ((size_t)g * R + r) * n3;
        b_r += b[j], b_z += b[H + j], b_n += b[2 * H + j];
      }
      const float rg = sigm(g_r[q] + b_r);
      const float zg = sigm(g_z[q] + b_z);
      const float ghn = b_n + bh[q];
      const float ng = tanhf(g_n[q] + rg * ghn);
      const float hpv = hs[j * R + r];
      const float h = (1.f - zg) * ng + zg * hpv;
      const int64_t o = (int64_t)t * eh + row * H + j;
      h_seq[o] = h, rs[o] = rg, zs[o] = zg, ns[o] = ng, ghns[o] = ghn;
      const float hn = cut[q] ? 0.f : h;
      hp_seq[o + eh] = hn;
      hs[j * R + r] = hn;   // only this thread touches (j, r) in this phase; the product above is behind the barrier
    }
    __syncthreads();
  }
}

template <int R, int KS>
__global__ void __launch_bounds__(1024, 1)
    gru_seq_bwd_kernel(const float* __restrict__ d_h_seq, const uint8_t* __restrict__ reset, const float* __restrict__ w_h, int T, int64_t E, int H,
                       const float* __restrict__ hp_seq, const float* __restrict__ rs, const float* __restrict__ zs, const float* __restrict__ ns,
                       const float* __restrict__ ghns, float* __restrict__ d_gi, float* __restrict__ d_gh_seq, float* __restrict__ d_h0) {
  extern __shared__ __align__(16) float gsm[];
  const int n3 = 3 * H;
  constexpr int NB = 3 * KS;                       // column blocks of the transposed product
  float* Ws = gsm;                                 // [H][3H + 1]
  float* dgT = Ws + GruSmem<R, KS>::w_floats(H);   // [3H][R]  d(gh_t), n-major
  float* part = dgT + (size_t)n3 * R;              // [NB][R][H] partial sums of d(gh) W_h^T over the column blocks
  float* dh_rec = part + (size_t)NB * R * H;       // [R][H]   gradient arriving from step t + 1 (already masked by its reset)
  const int64_t row0 = (int64_t)blockIdx.x * R;
  const int64_t eh = E * (int64_t)H;
  const int items = H * R;
  gru_load_w(Ws, w_h, H);
  for (int i = threadIdx.x; i < items; i += blockDim.x) dh_rec[i] = 0.f;
  // saved values of the step the gate phase will process next (requested one step ahead)
  float v_dh[kGruItems], v_r[kGruItems], v_z[kGruItems], v_n[kGruItems], v_g[kGruItems], v_hp[kGruItems];
  bool v_cut[kGruItems];
  auto request = [&](int t) {
#pragma unroll
    for (int q = 0; q < kGruItems; ++q) {
      const int i = threadIdx.x + q * blockDim.x;
      v_dh[q] = v_r[q] = v_z[q] = v_n[q] = v_g[q] = v_hp[q] = 0.f, v_cut[q] = false;
      if (i < items && t >= 0) {
        const int r = i / H, j = i % H;
        const int64_t row = row0 + r;
        if (row < E) {
          const int64_t o = (int64_t)t * eh + row * H + j;
          v_dh[q] = __ldg(d_h_seq + o), v_r[q] = __ldg(rs + o), v_z[q] = __ldg(zs + o), v_n[q] = __ldg(ns + o), v_g[q] = __ldg(ghns + o);
          v_hp[q] = __ldg(hp_seq + o);
          v_cut[q] = reset[(int64_t)t * E + row] != 0;
        }
      }
    }
  };
  request(T - 1);
  __syncthreads();
  const int kcol = (H + KS - 1) / KS;               // columns per block of the transposed product
  for (int t = T - 1; t >= 0; --t) {
    bool cut_t[kGruItems];
#pragma unroll
    for (int q = 0; q < kGruItems; ++q) {            // gate gradients of step t from the values requested earlier
      const int i = threadIdx.x + q * blockDim.x;
      cut_t[q] = v_cut[q];
      if (i >= items) continue;
      const int r = i / H, j = i % H;
      const int64_t row = row0 + r;
      float dpr = 0.f, dpz = 0.f, dpn = 0.f, rg = 0.f, direct = 0.f;
      if (row < E) {
        const float dh = v_dh[q] + dh_rec[r * H + j];
        rg = v_r[q];
        const float zg = v_z[q], ng = v_n[q], ghn = v_g[q], hpv = v_hp[q];
        const float dn = dh * (1.f - zg), dz = dh * (hpv - ng);
        dpn = dn * (1.f - ng * ng);
        dpr = dpn * ghn * rg * (1.f - rg);
        dpz = dz * zg * (1.f - zg);
        direct = dh * zg;
        float* a = d_gi + ((int64_t)t * E + row) * n3;
        float* b = d_gh_seq + ((int64_t)t * E + row) * n3;
        a[j] = dpr, a[H + j] = dpz, a[2 * H + j] = dpn;
        b[j] = dpr, b[H + j] = dpz, b[2 * H + j] = dpn * rg;
      }
      dgT[j * R + r] = dpr, dgT[(H + j) * R + r] = dpz, dgT[(2 * H + j) * R + r] = dpn * rg;
      dh_rec[r * H + j] = direct;   // the direct part of d(hp_t); the W_h^T part is added below
    }
    request(t - 1);                 // in flight under the product
    __syncthreads();
    if ((int)threadIdx.x < NB * H) {   // d(gh_t) W_h^T: thread (k, block) sums the columns of its block
      const int k = threadIdx.x % H, blk = threadIdx.x / H;
      const int gate = blk / KS, sub = blk % KS;                 // block = (gate, sub-range of its H columns)
      const int c0 = gate * H + sub * kcol, c1 = (sub * kcol + kcol < H) ? c0 + kcol : gate * H + H;
      float2 acc[R / 2];
#pragma unroll
      for (int r = 0; r < R / 2; ++r) acc[r] = make_float2(0.f, 0.f);
      const float* wrow = Ws + (size_t)k * (n3 + 1);
#pragma unroll 4
      for (int nn = c0; nn < c1; ++nn) {
        const float w = wrow[nn];
#pragma unroll
        for (int r4 = 0; r4 < R; r4 += 4) {
          const float4 dv = *reinterpret_cast<const float4*>(dgT + (size_t)nn * R + r4);
          fma2(acc[r4 / 2], w, make_float2(dv.x, dv.y)), fma2(acc[r4 / 2 + 1], w, make_float2(dv.z, dv.w));
        }
      }
#pragma unroll
      for (int r = 0; r < R / 2; ++r) part[((size_t)blk * R + 2 * r) * H + k] = acc[r].x, part[((size_t)blk * R + 2 * r + 1) * H + k] = acc[r].y;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kGruItems; ++q) {
      const int i = threadIdx.x + q * blockDim.x;
      if (i >= items) continue;
      const int r = i / H, j = i % H;
      const int64_t row = row0 + r;
      float dhp = dh_rec[i];
#pragma unroll
      for (int blk = 0; blk < NB; ++blk) dhp += part[((size_t)blk * R + r) * H + j];   // fixed order
      dh_rec[i] = (row < E && !cut_t[q]) ? dhp : 0.f;   // base.py:139-148: the reset of step t cuts the chain into step t - 1
    }
    __syncthreads();
  }
  if (d_h0)
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
      const int r = i / H, j = i % H;
      if (row0 + r < E) d_h0[(row0 + r) * H + j] = dh_rec[i];
    }
}

// ---- LSTM (flax.linen.LSTMCell), per-step form ----------------------------------------------------------------------------
// i = sigmoid(z_i), f = sigmoid(z_f), g = tanh(z_g), o = sigmoid(z_o) with z = gi_t + hp W_h (gi carries the four hidden biases);
// c' = f cp + i g, h' = o tanh(c'); carry = (c | h), both halves zeroed at a reset.  W_h (H x 4H fp32) is 262 KB at H = 128 and
// does not fit one SM's shared memory, so this cell runs as one small-grid GEMM + one gate kernel per step (a 2-CTA cluster
// holding half of the hidden units each is the persistent form it would need).  d(gi_t) = d(gh_t) = d(z_t): one buffer.
struct LstmWs {
  float *hp_seq, *cp_seq;          // [T + 1][E][H]  h / c entering step t (after the reset)
  float *gi_, *gf_, *gg_, *go_, *tc_;  // [T][E][H]
  float *gh;                       // [E][4H]
  float *dhp_gemm;                 // [E][H]
  float *dcp[2];                   // [E][H]
  float *partials;                 // [splits][H * 4H]
  int splits;
  size_t bytes;
};
LstmWs carve_lstm(int T, int64_t E, int H, char* base) {
  LstmWs w{};
  size_t o = 0;
  auto take = [&](size_t floats) {
    float* p = base ? reinterpret_cast<float*>(base + o) : nullptr;
    o += ralign(floats * 4);
    return p;
  };
  const size_t eh = (size_t)E * H;
  w.hp_seq = take((size_t)(T + 1) * eh), w.cp_seq = take((size_t)(T + 1) * eh);
  w.gi_ = take((size_t)T * eh), w.gf_ = take((size_t)T * eh), w.gg_ = take((size_t)T * eh), w.go_ = take((size_t)T * eh), w.tc_ = take((size_t)T * eh);
  w.gh = take(4 * eh);
  w.dhp_gemm = take(eh);
  w.dcp[0] = take(eh), w.dcp[1] = take(eh);
  w.splits = gru_splits((int64_t)T * E);
  w.partials = take((size_t)w.splits * (size_t)H * 4 * H);
  w.bytes = o;
  return w;
}

__global__ void lstm_init_kernel(const float* __restrict__ carry0, const uint8_t* __restrict__ reset0, int64_t E, int H, float* __restrict__ cp0,
                                 float* __restrict__ hp0) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= E * H) return;
  const int64_t e = i / H;
  const int k = (int)(i % H);
  const bool cut = reset0[e] != 0;
  cp0[i] = cut ? 0.f : carry0[e * 2 * H + k];
  hp0[i] = cut ? 0.f : carry0[e * 2 * H + H + k];
}

__global__ void lstm_gate_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ cp,
                                     const uint8_t* __restrict__ reset_next, int64_t E, int H, float* __restrict__ h_out, float* __restrict__ cp_next,
                                     float* __restrict__ hp_next, float* __restrict__ si, float* __restrict__ sf, float* __restrict__ sg,
                                     float* __restrict__ so, float* __restrict__ stc, float* __restrict__ carry_last) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= E * H) return;
  const int64_t e = idx / H;
  const int k = (int)(idx % H);
  const float* a = gi + e * 4 * H;
  const float* b = gh + e * 4 * H;
  const float ig = sigm(a[k] + b[k]), fg = sigm(a[H + k] + b[H + k]), gg = tanhf(a[2 * H + k] + b[2 * H + k]), og = sigm(a[3 * H + k] + b[3 * H + k]);
  const float c = fg * cp[idx] + ig * gg;
  const float tc = tanhf(c);
  const float h = og * tc;
  h_out[idx] = h;
  si[idx] = ig, sf[idx] = fg, sg[idx] = gg, so[idx] = og, stc[idx] = tc;
  const bool cut = reset_next && reset_next[e];
  cp_next[idx] = cut ? 0.f : c;
  hp_next[idx] = cut ? 0.f : h;
  if (carry_last) carry_last[e * 2 * H + k] = c, carry_last[e * 2 * H + H + k] = h;
}

__global__ void lstm_gate_bwd_kernel(const float* __restrict__ d_h_out, const float* __restrict__ dhp_gemm_next, const float* __restrict__ dcp_next,
                                     const uint8_t* __restrict__ reset_next, const float* __restrict__ si, const float* __restrict__ sf,
                                     const float* __restrict__ sg, const float* __restrict__ so, const float* __restrict__ stc,
                                     const float* __restrict__ cp, int64_t E, int H, float* __restrict__ d_z, float* __restrict__ dcp_out) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= E * H) return;
  const int64_t e = idx / H;
  const int k = (int)(idx % H);
  float dh = d_h_out[idx], dc = 0.f;
  if (dhp_gemm_next && !reset_next[e]) dh += dhp_gemm_next[idx], dc = dcp_next[idx];   // the reset cuts both halves of the carry
  const float ig = si[idx], fg = sf[idx], gg = sg[idx], og = so[idx], tc = stc[idx];
  dc += dh * og * (1.f - tc * tc);
  float* z = d_z + e * 4 * H;
  z[k] = dc * gg * ig * (1.f - ig);
  z[H + k] = dc * cp[idx] * fg * (1.f - fg);
  z[2 * H + k] = dc * ig * (1.f - gg * gg);
  z[3 * H + k] = dh * tc * og * (1.f - og);
  dcp_out[idx] = dc * fg;
}

__global__ void lstm_dcarry_kernel(const float* __restrict__ dhp_gemm, const float* __restrict__ dcp, const uint8_t* __restrict__ reset0, int64_t E, int H,
                                   float* __restrict__ d_carry0) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= E * H) return;
  const int64_t e = i / H;
  const int k = (int)(i % H);
  const bool cut = reset0[e] != 0;
  d_carry0[e * 2 * H + k] = cut ? 0.f : dcp[i];
  d_carry0[e * 2 * H + H + k] = cut ? 0.f : dhp_gemm[i];
}

// ---- LSTM, persistent form on a 2-CTA thread-block cluster ---------------------------------------------------------------
// W_h (H x 4H fp32 = 262 KB at H = 128) does not fit one SM, so a CLUSTER of two CTAs shares each tile of R sequences: CTA q owns
// the hidden units [q H/2, (q+1) H/2) -- all four gate columns of those units, 131 KB of W_h -- computes their gates and cell
// update, and after every step pushes its half of the new h into the PEER's shared memory (distributed shared memory) next to its
// own, into the buffer the next step reads (two h buffers alternate, so a fast CTA never overwrites what its peer still reads);
// one cluster barrier per step replaces the block barrier.  Backward: each CTA forms d(z) of its units, multiplies with its
// 2H columns of W_h^T for ALL H outputs and sends the half that belongs to the peer's units across; the two partial sums are
// added in rank order (deterministic).  No per-step launch, no global synchronisation.
namespace cg = cooperative_groups;
constexpr int kLstmR = 4, kLstmKS = 2;
struct LstmSmem {
  __host__ __device__ static size_t w_floats(int H) { return ((size_t)H * (2 * H + 1) + 3) / 4 * 4; }
  static size_t fwd_bytes(int H) { return (w_floats(H) + 2 * (size_t)H * kLstmR + (size_t)(H / 2) * kLstmR + (size_t)kLstmKS * kLstmR * 2 * H) * 4; }
  static size_t bwd_bytes(int H) {
    return (w_floats(H) + (size_t)2 * H * kLstmR + 2 * (size_t)kLstmKS * kLstmR * H + 2 * (size_t)kLstmR * (H / 2) + 2 * (size_t)kLstmR * (H / 2)) * 4;
  }
};

// this CTA's G * H / 2 columns of W_h (G gates): local column g * Hq + jl  <->  global column g * H + q * Hq + jl
__device__ __forceinline__ void cluster_load_w(float* __restrict__ Ws, const float* __restrict__ w_h, int H, int q, int G) {
  const int Hq = H / 2, nl = G * Hq, total = H * nl;
  constexpr int U = 8;
  for (int base = threadIdx.x; base < total; base += blockDim.x * U) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * blockDim.x;
      if (i < total) {
        const int k = i / nl, c = i % nl;
        v[u] = __ldg(w_h + (size_t)k * G * H + (c / Hq) * H + q * Hq + c % Hq);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * blockDim.x;
      if (i < total) Ws[(i / nl) * (nl + 1) + i % nl] = v[u];
    }
  }
}

__global__ void __launch_bounds__(1024, 1)
    lstm_seq_fwd_kernel(const float* __restrict__ gi, const uint8_t* __restrict__ reset, const float* __restrict__ carry0, const float* __restrict__ w_h,
                        int T, int64_t E, int H, float* __restrict__ h_seq, float* __restrict__ carry_last, float* __restrict__ hp_seq,
                        float* __restrict__ cp_seq, float* __restrict__ si, float* __restrict__ sf, float* __restrict__ sg, float* __restrict__ so,
                        float* __restrict__ stc) {
  constexpr int R = kLstmR, KS = kLstmKS;
  cg::cluster_group cluster = cg::this_cluster();
  const int q = (int)cluster.block_rank();
  extern __shared__ __align__(16) float gsm[];
  const int Hq = H / 2, nl = 2 * H;
  float* Ws = gsm;                                  // [H][2H + 1]
  float* hs = Ws + LstmSmem::w_floats(H);           // [2][H][R]  full h entering the step (k-major), two alternating buffers
  float* cs = hs + 2 * (size_t)H * R;               // [Hq][R]    c of this CTA's units
  float* ghp = cs + (size_t)Hq * R;                 // [KS][R][2H] partial products
  float* hs_peer = cluster.map_shared_rank(hs, q ^ 1);
  const int kh = threadIdx.x / nl, n = threadIdx.x % nl;
  const int kper = (H + KS - 1) / KS;
  const int k_begin = kh * kper, k_end = (k_begin + kper < H) ? k_begin + kper : H;
  const int64_t row0 = (int64_t)(blockIdx.x / 2) * R;
  const int items = Hq * R;
  const int64_t eh = E * (int64_t)H;
  cluster_load_w(Ws, w_h, H, q, 4);
  for (int i = threadIdx.x; i < H * R; i += blockDim.x) {     // every CTA initialises the FULL h of buffer 0 itself
    const int r = i / H, j = i % H;
    const int64_t row = row0 + r;
    float hv = 0.f;
    if (row < E && !reset[row]) hv = carry0[row * 2 * H + H + j];
    hs[j * R + r] = hv;
    if (row < E && j / Hq == q) {
      const float cv = reset[row] ? 0.f : carry0[row * 2 * H + j];
      cs[(j - q * Hq) * R + r] = cv;
      hp_seq[row * H + j] = hv, cp_seq[row * H + j] = cv;
    } else if (row >= E && j / Hq == q) {
      cs[(j - q * Hq) * R + r] = 0.f;
    }
  }
  cluster.sync();
  for (int t = 0; t < T; ++t) {
    const float* hcur = hs + (size_t)(t & 1) * H * R;
    const size_t nxt = (size_t)((t + 1) & 1) * H * R;
    // requests for the gate phase (one item per thread: Hq * R <= blockDim)
    float gz[4] = {0.f, 0.f, 0.f, 0.f};
    bool cut = false;
    const int i0 = threadIdx.x;
    const int r0 = i0 / Hq, jl0 = i0 % Hq;
    const int64_t rowg = row0 + r0;
    if (i0 < items && rowg < E) {
      const float* a = gi + ((int64_t)t * E + rowg) * 4 * H + q * Hq + jl0;
      gz[0] = __ldg(a), gz[1] = __ldg(a + H), gz[2] = __ldg(a + 2 * H), gz[3] = __ldg(a + 3 * H);
      cut = (t + 1 < T) && reset[(int64_t)(t + 1) * E + rowg] != 0;
    }
    if (kh < KS) {
      float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
      const float* wcol = Ws + n;
#pragma unroll 4
      for (int k = k_begin; k < k_end; ++k) {
        const float w = wcol[(size_t)k * (nl + 1)];
        const float4 hv = *reinterpret_cast<const float4*>(hcur + k * R);
        fma2(a01, w, make_float2(hv.x, hv.y)), fma2(a23, w, make_float2(hv.z, hv.w));
      }
      const float acc[R] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
      for (int r = 0; r < R; ++r) ghp[((size_t)kh * R + r) * nl + n] = acc[r];
    }
    __syncthreads();
    if (i0 < items) {
      const int j = q * Hq + jl0;
      float h = 0.f, hn = 0.f;
      if (rowg < E) {
        float z[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float b = 0.f;
#pragma unroll
          for (int kk = 0; kk < KS; ++kk) b += ghp[((size_t)kk * R + r0) * nl + g * Hq + jl0];   // fixed order
          z[g] = gz[g] + b;
        }
        const float ig = sigm(z[0]), fg = sigm(z[1]), gg = tanhf(z[2]), og = sigm(z[3]);
        const float c = fg * cs[jl0 * R + r0] + ig * gg;
        const float tc = tanhf(c);
        h = og * tc;
        const int64_t o = (int64_t)t * eh + rowg * H + j;
        h_seq[o] = h, si[o] = ig, sf[o] = fg, sg[o] = gg, so[o] = og, stc[o] = tc;
        hn = cut ? 0.f : h;
        const float cn = cut ? 0.f : c;
        hp_seq[o + eh] = hn, cp_seq[o + eh] = cn;
        cs[jl0 * R + r0] = cn;
        if (t == T - 1 && carry_last) carry_last[rowg * 2 * H + j] = c, carry_last[rowg * 2 * H + H + j] = h;
      }
      hs[nxt + (size_t)j * R + r0] = hn;         // my half of the next step's h: here ...
      hs_peer[nxt + (size_t)j * R + r0] = hn;    // ... and in the peer's shared memory
    }
    cluster.sync();   // both halves of h_{t+1} are in place in both CTAs; also this block's barrier
  }
}

__global__ void __launch_bounds__(1024, 1)
    lstm_seq_bwd_kernel(const float* __restrict__ d_h_seq, const uint8_t* __restrict__ reset, const float* __restrict__ w_h, int T, int64_t E, int H,
                        const float* __restrict__ cp_seq, const float* __restrict__ si, const float* __restrict__ sf, const float* __restrict__ sg,
                        const float* __restrict__ so, const float* __restrict__ stc, float* __restrict__ d_gi, float* __restrict__ d_carry0) {
  constexpr int R = kLstmR, KS = kLstmKS;
  cg::cluster_group cluster = cg::this_cluster();
  const int q = (int)cluster.block_rank();
  extern __shared__ __align__(16) float gsm[];
  const int Hq = H / 2, nl = 2 * H;
  float* Ws = gsm;                                  // [H][2H + 1]  (this CTA's columns)
  float* dzT = Ws + LstmSmem::w_floats(H);          // [2H][R]   d(z_t) of this CTA's units, column-major
  float* part = dzT + (size_t)nl * R;               // [2 KS][R][H]... laid out as [blk][R][H]: partial d(z) W^T over this CTA's column blocks
  float* inbox = part + 2 * (size_t)KS * R * H;     // [2][R][Hq]  the peer's partial sums for MY units (two alternating buffers)
  float* dh_rec = inbox + 2 * (size_t)R * Hq;       // [R][Hq]
  float* dc_rec = dh_rec + (size_t)R * Hq;          // [R][Hq]
  float* inbox_peer = cluster.map_shared_rank(inbox, q ^ 1);
  const int64_t row0 = (int64_t)(blockIdx.x / 2) * R;
  const int64_t eh = E * (int64_t)H;
  const int items = Hq * R;
  cluster_load_w(Ws, w_h, H, q, 4);
  for (int i = threadIdx.x; i < items; i += blockDim.x) dh_rec[i] = 0.f, dc_rec[i] = 0.f;
  const int i0 = threadIdx.x;
  const int r0 = i0 / Hq, jl0 = i0 % Hq;
  const int64_t rowg = row0 + r0;
  const int j0 = q * Hq + jl0;
  float v_dh = 0.f, v_i = 0.f, v_f = 0.f, v_g = 0.f, v_o = 0.f, v_tc = 0.f, v_cp = 0.f;
  bool v_cut = false;
  auto request = [&](int t) {
    v_dh = v_i = v_f = v_g = v_o = v_tc = v_cp = 0.f, v_cut = false;
    if (i0 < items && rowg < E && t >= 0) {
      const int64_t o = (int64_t)t * eh + rowg * H + j0;
      v_dh = __ldg(d_h_seq + o), v_i = __ldg(si + o), v_f = __ldg(sf + o), v_g = __ldg(sg + o), v_o = __ldg(so + o), v_tc = __ldg(stc + o);
      v_cp = __ldg(cp_seq + o);
      v_cut = reset[(int64_t)t * E + rowg] != 0;
    }
  };
  request(T - 1);
  cluster.sync();
  const int NB = 2 * KS;                     // column blocks of this CTA's 2H columns
  const int kcol = (nl + NB - 1) / NB;
  for (int t = T - 1; t >= 0; --t) {
    bool cut_t = v_cut;
    float dcp = 0.f;
    if (i0 < items) {
      float dz[4] = {0.f, 0.f, 0.f, 0.f};
      if (rowg < E) {
        const float dh = v_dh + dh_rec[r0 * Hq + jl0];
        const float dc = dc_rec[r0 * Hq + jl0] + dh * v_o * (1.f - v_tc * v_tc);
        dz[0] = dc * v_g * v_i * (1.f - v_i);
        dz[1] = dc * v_cp * v_f * (1.f - v_f);
        dz[2] = dc * v_i * (1.f - v_g * v_g);
        dz[3] = dh * v_tc * v_o * (1.f - v_o);
        dcp = dc * v_f;
        float* a = d_gi + ((int64_t)t * E + rowg) * 4 * H + j0;
        a[0] = dz[0], a[H] = dz[1], a[2 * H] = dz[2], a[3 * H] = dz[3];
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) dzT[(size_t)(g * Hq + jl0) * R + r0] = dz[g];
    }
    request(t - 1);
    __syncthreads();
    if ((int)threadIdx.x < NB * H) {           // partial d(z) W^T over this CTA's columns, for ALL H outputs k
      const int k = threadIdx.x % H, blk = threadIdx.x / H;
      const int c0 = blk * kcol, c1 = (c0 + kcol < nl) ? c0 + kcol : nl;
      float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
      const float* wrow = Ws + (size_t)k * (nl + 1);
#pragma unroll 4
      for (int nn = c0; nn < c1; ++nn) {
        const float w = wrow[nn];
        const float4 dv = *reinterpret_cast<const float4*>(dzT + (size_t)nn * R);
        fma2(a01, w, make_float2(dv.x, dv.y)), fma2(a23, w, make_float2(dv.z, dv.w));
      }
      const float acc[R] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
      for (int r = 0; r < R; ++r) part[((size_t)blk * R + r) * H + k] = acc[r];
    }
    __syncthreads();
    // my partial for the PEER's units goes into its inbox; my partial for my own units stays in a register
    float mine = 0.f;
    float* box = inbox + (size_t)(t & 1) * R * Hq;
    if (i0 < items) {
      float other = 0.f;
      const int jp = (q ^ 1) * Hq + jl0;
      for (int blk = 0; blk < NB; ++blk) {       // fixed order
        mine += part[((size_t)blk * R + r0) * H + j0];
        other += part[((size_t)blk * R + r0) * H + jp];
      }
      inbox_peer[(size_t)(t & 1) * R * Hq + r0 * Hq + jl0] = other;
    }
    cluster.sync();
    if (i0 < items) {
      const float theirs = box[r0 * Hq + jl0];
      const float dhp = (q == 0) ? (mine + theirs) : (theirs + mine);   // rank-0 partial first on both CTAs
      const bool keep = rowg < E && !cut_t;
      dh_rec[r0 * Hq + jl0] = keep ? dhp : 0.f;
      dc_rec[r0 * Hq + jl0] = keep ? dcp : 0.f;
    }
    __syncthreads();
  }
  if (d_carry0 && i0 < items && rowg < E) {
    d_carry0[rowg * 2 * H + j0] = dc_rec[r0 * Hq + jl0];
    d_carry0[rowg * 2 * H + H + j0] = dh_rec[r0 * Hq + jl0];
  }
  cluster.sync();   // no CTA exits while its peer may still write into its shared memory
}

// ---- GRU on 2-CTA clusters: the same split (half of the hidden units, 3 H / 2 gate columns and 99 KB of W_h per CTA) halves the
// issue-bound per-step product of gru_seq_*_kernel; used for the training sequences (T >= kPersistentMinT, rows < 1024).
struct GruClSmem {
  __host__ __device__ static size_t w_floats(int H) { return ((size_t)H * (3 * (H / 2) + 1) + 3) / 4 * 4; }
  static size_t fwd_bytes(int H) { return (w_floats(H) + 2 * (size_t)H * kLstmR + (size_t)kLstmKS * kLstmR * 3 * (H / 2)) * 4; }
  static size_t bwd_bytes(int H) { return (w_floats(H) + (size_t)3 * (H / 2) * kLstmR + 3 * (size_t)kLstmR * H + 3 * (size_t)kLstmR * (H / 2)) * 4; }
};

__global__ void __launch_bounds__(1024, 1)
    gru_cl_fwd_kernel(const float* __restrict__ gi, const uint8_t* __restrict__ reset, const float* __restrict__ h0, const float* __restrict__ w_h,
                      const float* __restrict__ b_hn, int T, int64_t E, int H, float* __restrict__ h_seq, float* __restrict__ hp_seq,
                      float* __restrict__ rs, float* __restrict__ zs, float* __restrict__ ns, float* __restrict__ ghns) {
  constexpr int R = kLstmR, KS = kLstmKS;
  cg::cluster_group cluster = cg::this_cluster();
  const int q = (int)cluster.block_rank();
  extern __shared__ __align__(16) float gsm[];
  const int Hq = H / 2, nl = 3 * Hq;
  float* Ws = gsm;                                  // [H][nl + 1]
  float* hs = Ws + GruClSmem::w_floats(H);          // [2][H][R]
  float* ghp = hs + 2 * (size_t)H * R;              // [KS][R][nl]
  float* hs_peer = cluster.map_shared_rank(hs, q ^ 1);
  const int kh = threadIdx.x / nl, n = threadIdx.x % nl;
  const int kper = (H + KS - 1) / KS;
  const int k_begin = kh * kper, k_end = (k_begin + kper < H) ? k_begin + kper : H;
  const int64_t row0 = (int64_t)(blockIdx.x / 2) * R;
  const int items = Hq * R;
  const int64_t eh = E * (int64_t)H;
  cluster_load_w(Ws, w_h, H, q, 3);
  for (int i = threadIdx.x; i < H * R; i += blockDim.x) {
    const int r = i / H, j = i % H;
    const int64_t row = row0 + r;
    float hv = 0.f;
    if (row < E && !reset[row]) hv = h0[row * H + j];
    hs[j * R + r] = hv;
    if (row < E && j / Hq == q) hp_seq[row * H + j] = hv;
  }
  cluster.sync();
  const int i0 = threadIdx.x;
  const int r0 = i0 / Hq, jl0 = i0 % Hq;
  const int64_t rowg = row0 + r0;
  const int j0 = q * Hq + jl0;
  for (int t = 0; t < T; ++t) {
    const float* hcur = hs + (size_t)(t & 1) * H * R;
    const size_t nxt = (size_t)((t + 1) & 1) * H * R;
    float g3[3] = {0.f, 0.f, 0.f}, bh = 0.f;
    bool cut = false;
    if (i0 < items && rowg < E) {
      const float* a = gi + ((int64_t)t * E + rowg) * 3 * H + j0;
      g3[0] = __ldg(a), g3[1] = __ldg(a + H), g3[2] = __ldg(a + 2 * H), bh = __ldg(b_hn + j0);
      cut = (t + 1 < T) && reset[(int64_t)(t + 1) * E + rowg] != 0;
    }
    if (kh < KS) {
      float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
      const float* wcol = Ws + n;
#pragma unroll 4
      for (int k = k_begin; k < k_end; ++k) {
        const float w = wcol[(size_t)k * (nl + 1)];
        const float4 hv = *reinterpret_cast<const float4*>(hcur + k * R);
        fma2(a01, w, make_float2(hv.x, hv.y)), fma2(a23, w, make_float2(hv.z, hv.w));
      }
      const float acc[R] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
      for (int r = 0; r < R; ++r) ghp[((size_t)kh * R + r) * nl + n] = acc[r];
    }
    __syncthreads();
    if (i0 < items) {
      float hn = 0.f;
      if (rowg < E) {
        float b[3];
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          float v = 0.f;
#pragma unroll
          for (int kk = 0; kk < KS; ++kk) v += ghp[((size_t)kk * R + r0) * nl + g * Hq + jl0];   // fixed order
          b[g] = v;
        }
        const float rg = sigm(g3[0] + b[0]), zg = sigm(g3[1] + b[1]);
        const float ghn = b[2] + bh;
        const float ng = tanhf(g3[2] + rg * ghn);
        const float hpv = hcur[j0 * R + r0];
        const float h = (1.f - zg) * ng + zg * hpv;
        const int64_t o = (int64_t)t * eh + rowg * H + j0;
        h_seq[o] = h, rs[o] = rg, zs[o] = zg, ns[o] = ng, ghns[o] = ghn;
        hn = cut ? 0.f : h;
        hp_seq[o + eh] = hn;
      }
      hs[nxt + (size_t)j0 * R + r0] = hn;
      hs_peer[nxt + (size_t)j0 * R + r0] = hn;
    }
    cluster.sync();
  }
}

__global__ void __launch_bounds__(1024, 1)
    gru_cl_bwd_kernel(const float* __restrict__ d_h_seq, const uint8_t* __restrict__ reset, const float* __restrict__ w_h, int T, int64_t E, int H,
                      const float* __restrict__ hp_seq, const float* __restrict__ rs, const float* __restrict__ zs, const float* __restrict__ ns,
                      const float* __restrict__ ghns, float* __restrict__ d_gi, float* __restrict__ d_gh_seq, float* __restrict__ d_h0) {
  constexpr int R = kLstmR;
  cg::cluster_group cluster = cg::this_cluster();
  const int q = (int)cluster.block_rank();
  extern __shared__ __align__(16) float gsm[];
  const int Hq = H / 2, nl = 3 * Hq;
  constexpr int NB = 3;                             // column blocks of this CTA's 3 Hq columns (one gate each)
  float* Ws = gsm;                                  // [H][nl + 1]
  float* dgT = Ws + GruClSmem::w_floats(H);         // [nl][R]
  float* part = dgT + (size_t)nl * R;               // [NB][R][H]
  float* inbox = part + (size_t)NB * R * H;         // [2][R][Hq]
  float* dh_rec = inbox + 2 * (size_t)R * Hq;       // [R][Hq]
  float* inbox_peer = cluster.map_shared_rank(inbox, q ^ 1);
  const int64_t row0 = (int64_t)(blockIdx.x / 2) * R;
  const int64_t eh = E * (int64_t)H;
  const int items = Hq * R;
  cluster_load_w(Ws, w_h, H, q, 3);
  for (int i = threadIdx.x; i < items; i += blockDim.x) dh_rec[i] = 0.f;
  const int i0 = threadIdx.x;
  const int r0 = i0 / Hq, jl0 = i0 % Hq;
  const int64_t rowg = row0 + r0;
  const int j0 = q * Hq + jl0;
  float v_dh = 0.f, v_r = 0.f, v_z = 0.f, v_n = 0.f, v_g = 0.f, v_hp = 0.f;
  bool v_cut = false;
  auto request = [&](int t) {
    v_dh = v_r = v_z = v_n = v_g = v_hp = 0.f, v_cut = false;
    if (i0 < items && rowg < E && t >= 0) {
      const int64_t o = (int64_t)t * eh + rowg * H + j0;
      v_dh = __ldg(d_h_seq + o), v_r = __ldg(rs + o), v_z = __ldg(zs + o), v_n = __ldg(ns + o), v_g = __ldg(ghns + o), v_hp = __ldg(hp_seq + o);
      v_cut = reset[(int64_t)t * E + rowg] != 0;
    }
  };
  request(T - 1);
  cluster.sync();
  for (int t = T - 1; t >= 0; --t) {
    const bool cut_t = v_cut;
    float direct = 0.f;
    if (i0 < items) {
      float dpr = 0.f, dpz = 0.f, dpn = 0.f, rg = 0.f;
      if (rowg < E) {
        const float dh = v_dh + dh_rec[r0 * Hq + jl0];
        rg = v_r;
        const float dn = dh * (1.f - v_z), dz = dh * (v_hp - v_n);
        dpn = dn * (1.f - v_n * v_n);
        dpr = dpn * v_g * rg * (1.f - rg);
        dpz = dz * v_z * (1.f - v_z);
        direct = dh * v_z;
        float* a = d_gi + ((int64_t)t * E + rowg) * 3 * H + j0;
        float* b = d_gh_seq + ((int64_t)t * E + rowg) * 3 * H + j0;
        a[0] = dpr, a[H] = dpz, a[2 * H] = dpn;
        b[0] = dpr, b[H] = dpz, b[2 * H] = dpn * rg;
      }
      dgT[(size_t)jl0 * R + r0] = dpr, dgT[(size_t)(Hq + jl0) * R + r0] = dpz, dgT[(size_t)(2 * Hq + jl0) * R + r0] = dpn * rg;
    }
    request(t - 1);
    __syncthreads();
    if ((int)threadIdx.x < NB * H) {
      const int k = threadIdx.x % H, blk = threadIdx.x / H;
      const int c0 = blk * Hq, c1 = c0 + Hq;
      float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
      const float* wrow = Ws + (size_t)k * (nl + 1);
#pragma unroll 4
      for (int nn = c0; nn < c1; ++nn) {
        const float w = wrow[nn];
        const float4 dv = *reinterpret_cast<const float4*>(dgT + (size_t)nn * R);
        fma2(a01, w, make_float2(dv.x, dv.y)), fma2(a23, w, make_float2(dv.z, dv.w));
      }
      const float acc[R] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
      for (int r = 0; r < R; ++r) part[((size_t)blk * R + r) * H + k] = acc[r];
    }
    __syncthreads();
    float mine = 0.f;
    float* box = inbox + (size_t)(t & 1) * R * Hq;
    if (i0 < items) {
      float other = 0.f;
      const int jp = (q ^ 1) * Hq + jl0;
#pragma unroll
      for (int blk = 0; blk < NB; ++blk) {   // fixed order
        mine += part[((size_t)blk * R + r0) * H + j0];
        other += part[((size_t)blk * R + r0) * H + jp];
      }
      inbox_peer[(size_t)(t & 1) * R * Hq + r0 * Hq + jl0] = other;
    }
    cluster.sync();
    if (i0 < items) {
      const float theirs = box[r0 * Hq + jl0];
      const float dhp = direct + ((q == 0) ? (mine + theirs) : (theirs + mine));   // rank-0 partial first on both CTAs
      dh_rec[r0 * Hq + jl0] = (rowg < E && !cut_t) ? dhp : 0.f;
    }
    __syncthreads();
  }
  if (d_h0 && i0 < items && rowg < E) d_h0[rowg * H + j0] = dh_rec[r0 * Hq + jl0];
  cluster.sync();
}

inline int gru_cl_threads(int H) { return ((kLstmKS * 3 * (H / 2) + 31) / 32) * 32; }
inline bool gru_cluster_ok(int H, int64_t E, int T) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("STX_GRU_CLUSTER");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  return enabled && E < 1024 && T >= 4 && H >= 2 && H % 2 == 0 && gru_cl_threads(H) <= 1024 && (H / 2) * kLstmR <= gru_cl_threads(H) &&
         3 * H <= gru_cl_threads(H) && GruClSmem::fwd_bytes(H) <= 227 * 1024 && GruClSmem::bwd_bytes(H) <= 227 * 1024;
}

inline int lstm_threads(int H) { return ((kLstmKS * 2 * H + 31) / 32) * 32; }
inline bool lstm_cluster_ok(int H, int64_t E) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("STX_LSTM_CLUSTER");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  (void)E;
  return enabled && H >= 2 && H % 2 == 0 && lstm_threads(H) <= 1024 && (H / 2) * kLstmR <= lstm_threads(H) && 2 * kLstmKS * H <= lstm_threads(H) &&
         LstmSmem::fwd_bytes(H) <= 227 * 1024 && LstmSmem::bwd_bytes(H) <= 227 * 1024;
}
template <typename... Args>
inline cudaError_t launch_cluster2(void (*kernel)(Args...), unsigned clusters, int threads, size_t smem, cudaStream_t st, Args... args) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters), cfg.blockDim = dim3(threads), cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

constexpr size_t kGruSmemLimit = 227 * 1024;
constexpr int kPersistentMinT = 4;   // below: one GEMM + one gate kernel per step
// launch shape: R = 8 rows per CTA from 1024 sequences up (the rollout step), else 4 with the reduction split over two thread groups
inline int gru_threads(int H, int ks) { return ((ks * 3 * H + 31) / 32) * 32; }
inline bool gru_persistent_ok(int H) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("STX_GRU_PERSISTENT");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  return enabled && H >= 2 && gru_threads(H, 2) <= 1024 && GruSmem<8, 1>::fwd_bytes(H) <= kGruSmemLimit && GruSmem<8, 1>::bwd_bytes(H) <= kGruSmemLimit &&
         GruSmem<4, 2>::fwd_bytes(H) <= kGruSmemLimit && GruSmem<4, 2>::bwd_bytes(H) <= kGruSmemLimit;
}
template <typename K>
inline int gru_opt_in(K kernel) {
  STX_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGruSmemLimit));
  return STX_OK;
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
  if (gru_cluster_ok(H, E, T)) {   // training sequences: 2-CTA clusters, half of the gate columns per CTA
    STX_CUDA_OK(launch_cluster2(gru_cl_fwd_kernel, (unsigned)((E + kLstmR - 1) / kLstmR), gru_cl_threads(H), GruClSmem::fwd_bytes(H), st, gi, reset, h0, w_h, b_hn,
                                T, E, H, h_seq, ws.hp_seq, ws.r, ws.z, ws.n, ws.ghn));
    STX_LAUNCH_OK();
    return STX_OK;
  }
  // short sequences (the rollout's T = 1 step over all envs) do not amortise filling shared memory with W_h: per-step form
  if (gru_persistent_ok(H) && T >= kPersistentMinT) {
    if (E >= 1024) {
      if (int rc = gru_opt_in(gru_seq_fwd_kernel<8, 1>)) return rc;
      gru_seq_fwd_kernel<8, 1><<<(unsigned)((E + 7) / 8), gru_threads(H, 1), GruSmem<8, 1>::fwd_bytes(H), st>>>(gi, reset, h0, w_h, b_hn, T, E, H, h_seq, ws.hp_seq,
                                                                                                                  ws.r, ws.z, ws.n, ws.ghn);
    } else {
      if (int rc = gru_opt_in(gru_seq_fwd_kernel<4, 2>)) return rc;
      gru_seq_fwd_kernel<4, 2><<<(unsigned)((E + 3) / 4), gru_threads(H, 2), GruSmem<4, 2>::fwd_bytes(H), st>>>(gi, reset, h0, w_h, b_hn, T, E, H, h_seq, ws.hp_seq,
                                                                                                                  ws.r, ws.z, ws.n, ws.ghn);
    }
    STX_LAUNCH_OK();
    return STX_OK;
  }
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
  const bool clustered = gru_cluster_ok(H, E, T);
  const bool persistent = clustered || (gru_persistent_ok(H) && T >= kPersistentMinT);
  if (clustered) {
    STX_CUDA_OK(launch_cluster2(gru_cl_bwd_kernel, (unsigned)((E + kLstmR - 1) / kLstmR), gru_cl_threads(H), GruClSmem::bwd_bytes(H), st, d_h_seq, reset, w_h, T, E, H,
                                (const float*)ws.hp_seq, (const float*)ws.r, (const float*)ws.z, (const float*)ws.n, (const float*)ws.ghn, d_gi, ws.d_gh_seq,
                                d_h0));
    STX_LAUNCH_OK();
  } else if (persistent) {
    if (E >= 1024) {
      if (int rc = gru_opt_in(gru_seq_bwd_kernel<8, 1>)) return rc;
      gru_seq_bwd_kernel<8, 1><<<(unsigned)((E + 7) / 8), gru_threads(H, 1), GruSmem<8, 1>::bwd_bytes(H), st>>>(d_h_seq, reset, w_h, T, E, H, ws.hp_seq, ws.r, ws.z,
                                                                                                                  ws.n, ws.ghn, d_gi, ws.d_gh_seq, d_h0);
    } else {
      if (int rc = gru_opt_in(gru_seq_bwd_kernel<4, 2>)) return rc;
      gru_seq_bwd_kernel<4, 2><<<(unsigned)((E + 3) / 4), gru_threads(H, 2), GruSmem<4, 2>::bwd_bytes(H), st>>>(d_h_seq, reset, w_h, T, E, H, ws.hp_seq, ws.r, ws.z,
                                                                                                                  ws.n, ws.ghn, d_gi, ws.d_gh_seq, d_h0);
    }
    STX_LAUNCH_OK();
  }
  for (int t = T - 1; t >= 0 && !persistent; --t) {
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
  if (d_h0 && !persistent) {
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

extern "C" size_t stx_lstm_workspace_bytes(int T, int64_t E, int H) {
  if (T <= 0 || E <= 0 || H <= 0) return 0;
  return carve_lstm(T, E, H, nullptr).bytes;
}

extern "C" int stx_lstm_sequence_forward(const float* gi, const uint8_t* reset, const float* carry0, const float* w_h, int T, int64_t E, int H,
                                         float* h_seq, float* carry_last, void* workspace, size_t workspace_bytes, void* stream) {
  STX_REQUIRE(gi && reset && carry0 && w_h && h_seq && workspace, STX_E_ARG, "stx_lstm_sequence_forward: null pointer");
  STX_REQUIRE(T > 0 && E > 0 && H > 0, STX_E_SHAPE, "stx_lstm_sequence_forward: T=%d E=%lld H=%d", T, (long long)E, H);
  STX_REQUIRE(workspace_bytes >= stx_lstm_workspace_bytes(T, E, H), STX_E_WORKSPACE, "stx_lstm_sequence_forward: workspace %zu < %zu", workspace_bytes,
              stx_lstm_workspace_bytes(T, E, H));
  cudaStream_t st = (cudaStream_t)stream;
  LstmWs ws = carve_lstm(T, E, H, reinterpret_cast<char*>(workspace));
  const size_t eh = (size_t)E * H;
  const unsigned blocks = (unsigned)((eh + 255) / 256);
  if (lstm_cluster_ok(H, E) && T >= kPersistentMinT) {
    STX_CUDA_OK(launch_cluster2(lstm_seq_fwd_kernel, (unsigned)((E + kLstmR - 1) / kLstmR), lstm_threads(H), LstmSmem::fwd_bytes(H), st, gi, reset, carry0, w_h,
                                T, E, H, h_seq, carry_last, ws.hp_seq, ws.cp_seq, ws.gi_, ws.gf_, ws.gg_, ws.go_, ws.tc_));
    STX_LAUNCH_OK();
    return STX_OK;
  }
  lstm_init_kernel<<<blocks, 256, 0, st>>>(carry0, reset, E, H, ws.cp_seq, ws.hp_seq);
  STX_LAUNCH_OK();
  for (int t = 0; t < T; ++t) {
    simt::GemmArgs g{};
    g.A = ws.hp_seq + (size_t)t * eh, g.lda = H, g.B = w_h, g.C = ws.gh, g.mask_act = -1;
    g.M = E, g.N = 4 * H, g.K = H;
    STX_CUDA_OK(simt::launch_gemm<simt::FWD>(g, 1, st));
    lstm_gate_fwd_kernel<<<blocks, 256, 0, st>>>(gi + (size_t)t * 4 * eh, ws.gh, ws.cp_seq + (size_t)t * eh, t + 1 < T ? reset + (size_t)(t + 1) * E : nullptr, E, H,
                                                  h_seq + (size_t)t * eh, ws.cp_seq + (size_t)(t + 1) * eh, ws.hp_seq + (size_t)(t + 1) * eh,
                                                  ws.gi_ + (size_t)t * eh, ws.gf_ + (size_t)t * eh, ws.gg_ + (size_t)t * eh, ws.go_ + (size_t)t * eh,
                                                  ws.tc_ + (size_t)t * eh, (t == T - 1) ? carry_last : nullptr);
    STX_LAUNCH_OK();
  }
  return STX_OK;
}

extern "C" int stx_lstm_sequence_backward(const float* d_h_seq, const uint8_t* reset, const float* w_h, int T, int64_t E, int H, void* workspace,
                                          size_t workspace_bytes, float* d_gi, float* d_w_h, float grad_weight, int overwrite, float* d_carry0,
                                          void* stream) {
  STX_REQUIRE(d_h_seq && reset && w_h && workspace && d_gi, STX_E_ARG, "stx_lstm_sequence_backward: null pointer");
  STX_REQUIRE(T > 0 && E > 0 && H > 0, STX_E_SHAPE, "stx_lstm_sequence_backward: T=%d E=%lld H=%d", T, (long long)E, H);
  STX_REQUIRE(workspace_bytes >= stx_lstm_workspace_bytes(T, E, H), STX_E_WORKSPACE, "stx_lstm_sequence_backward: workspace %zu < %zu", workspace_bytes,
              stx_lstm_workspace_bytes(T, E, H));
  cudaStream_t st = (cudaStream_t)stream;
  LstmWs ws = carve_lstm(T, E, H, reinterpret_cast<char*>(workspace));
  const size_t eh = (size_t)E * H;
  const unsigned blocks = (unsigned)((eh + 255) / 256);
  const bool clustered = lstm_cluster_ok(H, E) && T >= kPersistentMinT;
  if (clustered) {
    STX_CUDA_OK(launch_cluster2(lstm_seq_bwd_kernel, (unsigned)((E + kLstmR - 1) / kLstmR), lstm_threads(H), LstmSmem::bwd_bytes(H), st, d_h_seq, reset, w_h, T, E,
                                H, (const float*)ws.cp_seq, (const float*)ws.gi_, (const float*)ws.gf_, (const float*)ws.gg_, (const float*)ws.go_,
                                (const float*)ws.tc_, d_gi, d_carry0));
    STX_LAUNCH_OK();
  }
  for (int t = T - 1; t >= 0 && !clustered; --t) {
    const bool last = (t == T - 1);
    lstm_gate_bwd_kernel<<<blocks, 256, 0, st>>>(d_h_seq + (size_t)t * eh, last ? nullptr : ws.dhp_gemm, last ? nullptr : ws.dcp[(t + 1) & 1],
                                                  last ? nullptr : reset + (size_t)(t + 1) * E, ws.gi_ + (size_t)t * eh, ws.gf_ + (size_t)t * eh,
                                                  ws.gg_ + (size_t)t * eh, ws.go_ + (size_t)t * eh, ws.tc_ + (size_t)t * eh, ws.cp_seq + (size_t)t * eh, E, H,
                                                  d_gi + (size_t)t * 4 * eh, ws.dcp[t & 1]);
    STX_LAUNCH_OK();
    if (t > 0 || d_carry0) {
      simt::GemmArgs d{};
      d.A = d_gi + (size_t)t * 4 * eh, d.lda = 4 * H, d.B = w_h, d.C = ws.dhp_gemm, d.mask_act = -1;
      d.M = E, d.N = H, d.K = 4 * H;
      STX_CUDA_OK(simt::launch_gemm<simt::DX>(d, 1, st));
    }
  }
  if (d_carry0 && !clustered) {
    lstm_dcarry_kernel<<<blocks, 256, 0, st>>>(ws.dhp_gemm, ws.dcp[0], reset, E, H, d_carry0);
    STX_LAUNCH_OK();
  }
  if (d_w_h) {   // d(W_h) = hp_seq^T d(z)_seq over all (t, e) rows
    const int64_t rows = (int64_t)T * E;
    const int64_t np = (int64_t)H * 4 * H;
    simt::GemmArgs g{};
    g.A = ws.hp_seq, g.lda = H, g.B = d_gi, g.C = ws.partials, g.dbias = nullptr, g.mask_act = -1;
    g.M = rows, g.N = 4 * H, g.K = H;
    g.rows_per_split = (rows + ws.splits - 1) / ws.splits;
    g.part_stride = np, g.dbias_stride = np;
    STX_CUDA_OK(simt::launch_gemm<simt::DW>(g, ws.splits, st));
    simt::reduce_partials_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(ws.partials, ws.splits, np, np, grad_weight, d_w_h, overwrite);
    STX_LAUNCH_OK();
  }
  return STX_OK;
}
