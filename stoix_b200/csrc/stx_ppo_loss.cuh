// Fused PPO loss + analytic output gradients (one thread per minibatch row).
//
// Reference: _actor_loss_fn / _critic_loss_fn (stoix/systems/ppo/anakin/ff_ppo.py:191-235) with
// ppo_clip_loss (stoix/utils/loss.py:17-32) and clipped_value_loss (loss.py:68-78); the reference
// gets gradients from jax.grad (ff_ppo.py:238-247), here they are written in closed form
// (SURVEY.md 8a "Gradient formulas"):
//   dL/dlogp_i  = -(1/m) A_i ratio_i  [unclipped branch active]
//   dlogp/dz_j  = 1[j=a] - p_j ;  dH/dz_j = -p_j (log p_j + H)
//   dL/dv_i     = (vf/m) * { (v-tgt)            if (v-tgt)^2 > (vclip-tgt)^2
//                          { (vclip-tgt)*1[|v-vold|<eps]  otherwise }
// Advantage standardisation (multistep.py:138-139) is applied on load from (mean, rstd).
#pragma once
#include "stx_common.cuh"

namespace stx {

constexpr int kMaxActions = 32;  // logits per thread kept in registers

struct LossArgs {
  const float* logits;   // (mb, A)   nullable -> skip actor
  const float* value;    // (mb)      nullable -> skip critic
  int64_t value_ld;      // stride between consecutive rows of `value`
  const int32_t* idx;    // perm + mb_off (nullable -> identity from row0)
  int64_t row0;          // used when idx == nullptr
  const int32_t* action;
  const float* logp_old;
  const float* v_old;
  const float* adv;
  const float* tgt;
  const float* adv_stats;  // nullable
  float* dlogits;        // (mb, A)
  float* dvalue;         // (mb)
  int64_t mb;
  int A;
  float clip_eps, ent_coef, vf_coef;
  double* partials;      // [grid][6]
  unsigned int* counter;
  float* metrics;        // [6] accumulated with weight
  float weight;
};

__global__ void __launch_bounds__(256) ppo_loss_grad_kernel(LossArgs g) {
  const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  float s_actor = 0.f, s_ent = 0.f, s_vloss = 0.f, s_adv = 0.f, s_pred = 0.f, s_tgt = 0.f;
  const float inv_m = 1.0f / (float)g.mb;
  if (m < g.mb) {
    const int64_t src = g.idx ? (int64_t)g.idx[m] : g.row0 + m;
    if (g.logits) {
      float z[kMaxActions];
      const int A = g.A;
      float zmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < kMaxActions; ++j)
        if (j < A) {
          z[j] = g.logits[m * A + j];
          zmax = fmaxf(zmax, z[j]);
        }
      float se = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxActions; ++j)
        if (j < A) se += expf(z[j] - zmax);
      const float lse = zmax + logf(se);
      const int a = g.action[src];
      float adv = g.adv[src];
      if (g.adv_stats) adv = (adv - g.adv_stats[0]) * g.adv_stats[1];
      float ent = 0.f, logp_a = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxActions; ++j)
        if (j < A) {
          const float lp = z[j] - lse;
          ent -= expf(lp) * lp;
          if (j == a) logp_a = lp;
        }
      const float ratio = expf(logp_a - g.logp_old[src]);
      const float l1 = ratio * adv;
      const float rc = fminf(fmaxf(ratio, 1.0f - g.clip_eps), 1.0f + g.clip_eps);
      const float l2 = rc * adv;
      s_actor = -fminf(l1, l2);
      s_ent = ent;
      s_adv = adv;
      const bool in_band = (ratio >= 1.0f - g.clip_eps) && (ratio <= 1.0f + g.clip_eps);
      const float dlogp = ((l1 < l2) || in_band) ? -adv * ratio * inv_m : 0.f;
      const float ce = g.ent_coef * inv_m;
#pragma unroll
      for (int j = 0; j < kMaxActions; ++j)
        if (j < A) {
          const float lp = z[j] - lse, p = expf(lp);
          // d/dz_j [ L_clip - ent_coef * H ]
          g.dlogits[m * A + j] = dlogp * ((j == a ? 1.f : 0.f) - p) + ce * p * (lp + ent);
        }
    }
    if (g.value) {
      const float v = g.value[m * g.value_ld], vo = g.v_old[src], tg = g.tgt[src];
      const float diff = v - vo;
      const float vclip = vo + fminf(fmaxf(diff, -g.clip_eps), g.clip_eps);
      const float e1 = v - tg, e2 = vclip - tg;
      const float q1 = e1 * e1, q2 = e2 * e2;
      s_vloss = 0.5f * fmaxf(q1, q2);
      s_pred = v;
      s_tgt = tg;
      const float g2 = (fabsf(diff) < g.clip_eps) ? e2 : 0.f;
      const float dv = q1 > q2 ? e1 : (q1 < q2 ? g2 : 0.5f * (e1 + g2));
      g.dvalue[m] = g.vf_coef * dv * inv_m;
    }
  }
  __shared__ double sm[32];
  double red[6] = {(double)s_actor, (double)s_ent, (double)s_vloss, (double)s_adv, (double)s_pred, (double)s_tgt};
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double r = block_sum<double>(red[k], sm);
    if (threadIdx.x == 0) g.partials[(int64_t)blockIdx.x * 6 + k] = r;
  }
  if (last_block_ticket(g.counter, gridDim.x)) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double s = 0.0;
      for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += g.partials[(int64_t)i * 6 + k];
      s = block_sum<double>(s, sm);
      if (threadIdx.x == 0) g.metrics[k] += g.weight * (float)(s / (double)g.mb);
    }
  }
}

}  // namespace stx
