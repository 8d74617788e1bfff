// K2 -- generalized advantage estimation as a chunked parallel reverse scan over (T, E).
//
// Reference: batch_truncated_generalized_advantage_estimation, stoix/utils/multistep.py:14-145, as
// called at stoix/systems/ppo/anakin/ff_ppo.py:164-179 (time-major, truncation flags, standardise).
// The reference runs a T-iteration lax.scan of (E,)-wide fusions (multistep.py:119-130).  Here the
// first-order linear recurrence  acc_t = delta_t + c_t * acc_{t+1},  c_t = disc_t*lambda_t*(1-trunc_t)
// is evaluated as a composition of affine maps:
//   * a thread owns L=4 consecutive timesteps of VEC (=4, float4) adjacent envs: one pass of coalesced
//     128-bit streaming loads, a local scan in registers yielding its chunk map x -> B + A*x;
//   * the 32 chunk maps of one env (one 128-step segment) are suffix-composed with a 5-step warp
//     shuffle scan; longer rollouts walk segments from the end carrying one value per env;
//   * every element is read once and written once: 22 B/element in PPO form.
// Advantage standardisation statistics (jax.nn.standardize, multistep.py:138-139) are reduced in
// double precision with a deterministic two-level reduction finished by the last block to arrive.
#include "stx_common.cuh"

namespace stx {
namespace {

constexpr int kChunks = 32;  // chunk lanes per block == warp width (suffix scan by shuffles)

struct PpoIn {  // exactly what ff_ppo.py:164-169 feeds
  const float *reward, *v_tm1, *v_t;
  const uint8_t *done, *trunc;
  float gamma, lambda, reward_scale;
  template <int VEC>
  __device__ __forceinline__ void load(int64_t off, float* d, float* c, float* v) const {
    float r[VEC], vt[VEC], dn[VEC], tr[VEC];
    if constexpr (VEC == 4) {
      float4 a = ldg_stream4(reward + off), b = ldg_stream4(v_tm1 + off), e = ldg_stream4(v_t + off);
      uint32_t fd = ldg_stream_u32(done + off), ft = ldg_stream_u32(trunc + off);
      r[0] = a.x, r[1] = a.y, r[2] = a.z, r[3] = a.w;
      v[0] = b.x, v[1] = b.y, v[2] = b.z, v[3] = b.w;
      vt[0] = e.x, vt[1] = e.y, vt[2] = e.z, vt[3] = e.w;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        dn[k] = ((fd >> (8 * k)) & 0xffu) ? 1.f : 0.f;
        tr[k] = ((ft >> (8 * k)) & 0xffu) ? 1.f : 0.f;
      }
    } else if constexpr (VEC == 2) {
      float2 a = ldg_stream2(reward + off), b = ldg_stream2(v_tm1 + off), e = ldg_stream2(v_t + off);
      uint32_t fd = ldg_stream_u16(done + off), ft = ldg_stream_u16(trunc + off);
      r[0] = a.x, r[1] = a.y, v[0] = b.x, v[1] = b.y, vt[0] = e.x, vt[1] = e.y;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        dn[k] = ((fd >> (8 * k)) & 0xffu) ? 1.f : 0.f;
        tr[k] = ((ft >> (8 * k)) & 0xffu) ? 1.f : 0.f;
      }
    } else {
      r[0] = __ldg(reward + off), v[0] = __ldg(v_tm1 + off), vt[0] = __ldg(v_t + off);
      dn[0] = __ldg(done + off) ? 1.f : 0.f, tr[0] = __ldg(trunc + off) ? 1.f : 0.f;
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float disc = (1.0f - dn[k]) * gamma;                      // ff_ppo.py:167-168
      d[k] = r[k] * reward_scale + disc * vt[k] - v[k];               // multistep.py:116
      c[k] = disc * lambda * (1.0f - tr[k]);                          // multistep.py:123
    }
  }
};

struct GenericIn {  // float discount / lambda / truncation arrays (multistep.py:97-105)
  const float *r, *disc, *lam_t, *v_tm1, *v_t, *trunc;
  float lambda;
  template <int VEC>
  __device__ __forceinline__ void load(int64_t off, float* d, float* c, float* v) const {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float dk = __ldg(disc + off + k);
      const float lk = lam_t ? __ldg(lam_t + off + k) : lambda;
      const float tk = trunc ? __ldg(trunc + off + k) : 0.f;
      v[k] = __ldg(v_tm1 + off + k);
      d[k] = __ldg(r + off + k) + dk * __ldg(v_t + off + k) - v[k];
      c[k] = dk * lk * (1.0f - tk);
    }
  }
};

// MINB: resident blocks per SM the register allocation must allow.  A block runs load -> barrier -> scan -> barrier ->
// store with nothing overlapped inside it, so HBM only stays busy if ANOTHER block of the same SM is in its load phase
// meanwhile: the saturating shapes use 2 x 512 threads (<= 64 registers) instead of 1 x 1024.
template <class In, int VEC, int QUADS, int kL = 4, int MINB = 1>
__global__ void __launch_bounds__(QUADS* kChunks, MINB)
    gae_scan_kernel(In in, int T, int E, float* __restrict__ adv, float* __restrict__ tgt,
                    int want_stats, double2* __restrict__ partials, unsigned int* counter,
                    float* __restrict__ stats) {
  constexpr int ENVS = QUADS * VEC;  // envs per block
  constexpr int kSeg = kL * kChunks;  // timesteps covered per pass over the block
  const int q = threadIdx.x % QUADS, chunk = threadIdx.x / QUADS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e0 = (blockIdx.x * QUADS + q) * VEC;
  const bool env_ok = e0 < E;

  __shared__ float sA[kChunks][ENVS + 1], sB[kChunks][ENVS + 1];  // sA is reused for acc_in
  __shared__ float sCarry[ENVS];
  __shared__ double sRed[32];
  if (threadIdx.x < ENVS) sCarry[threadIdx.x] = 0.f;  // multistep.py:127: acc starts at zero

  float lsum = 0.f, lsq = 0.f;
  const int nseg = (T + kSeg - 1) / kSeg;
  for (int s = nseg - 1; s >= 0; --s) {
    const int tbase = s * kSeg + chunk * kL;
    float d[kL][VEC], c[kL][VEC], v[kL][VEC];
#pragma unroll
    for (int j = 0; j < kL; ++j) {
      const int t = tbase + j;
      if (env_ok && t < T) {
        in.template load<VEC>((int64_t)t * E + e0, d[j], c[j], v[j]);
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) d[j][k] = 0.f, c[j][k] = 1.f, v[j][k] = 0.f;  // identity map
      }
    }
    // local reverse scan: adv[j] = b[j] + a[j] * acc_in
    float a[kL][VEC], b[kL][VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float rb = 0.f, ra = 1.f;
#pragma unroll
      for (int j = kL - 1; j >= 0; --j) {
        rb = fmaf(c[j][k], rb, d[j][k]);
        ra = c[j][k] * ra;
        b[j][k] = rb;
        a[j][k] = ra;
      }
      sA[chunk][q * VEC + k] = ra;
      sB[chunk][q * VEC + k] = rb;
    }
    __syncthreads();
    // suffix composition over the 32 chunks of each env: warp w owns envs w*VEC .. w*VEC+VEC-1
    // (QUADS warps per block, ENVS = QUADS*VEC), lane = chunk index.
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int env = warp * VEC + k;
      float A = sA[lane][env], B = sB[lane][env];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const float A2 = __shfl_down_sync(0xffffffffu, A, off);
        const float B2 = __shfl_down_sync(0xffffffffu, B, off);
        if (lane + off < 32) {
          B = fmaf(A, B2, B);
          A = A * A2;
        }
      }
      const float carry = sCarry[env];
      const float Ae = __shfl_down_sync(0xffffffffu, A, 1), Be = __shfl_down_sync(0xffffffffu, B, 1);
      sA[lane][env] = (lane == 31) ? carry : fmaf(Ae, carry, Be);  // acc entering this chunk
      __syncwarp();
      if (lane == 0) sCarry[env] = fmaf(A, carry, B);  // acc entering the previous (earlier) segment
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kL; ++j) {
      const int t = tbase + j;
      float av[VEC], tv[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        av[k] = fmaf(a[j][k], sA[chunk][q * VEC + k], b[j][k]);
        tv[k] = v[j][k] + av[k];  // multistep.py:132: targets use the un-standardised advantage
      }
      if (env_ok && t < T) {
        const int64_t off = (int64_t)t * E + e0;
        if constexpr (VEC == 4) {
          stg_stream4(adv + off, make_float4(av[0], av[1], av[2], av[3]));
          stg_stream4(tgt + off, make_float4(tv[0], tv[1], tv[2], tv[3]));
        } else if constexpr (VEC == 2) {
          stg_stream2(adv + off, make_float2(av[0], av[1]));
          stg_stream2(tgt + off, make_float2(tv[0], tv[1]));
        } else {
          adv[off] = av[0];
          tgt[off] = tv[0];
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) lsum += av[k], lsq = fmaf(av[k], av[k], lsq);
      }
    }
  }
  if (!want_stats) return;
  const double bs = block_sum<double>((double)lsum, sRed);
  const double bq = block_sum<double>((double)lsq, sRed);
  if (threadIdx.x == 0) partials[blockIdx.x] = make_double2(bs, bq);
  if (last_block_ticket(counter, gridDim.x)) {
    double s = 0.0, q2 = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
      const double2 p = partials[i];
      s += p.x, q2 += p.y;
    }
    s = block_sum<double>(s, sRed);
    q2 = block_sum<double>(q2, sRed);
    if (threadIdx.x == 0) {
      const double n = (double)T * (double)E;
      const double mean = s / n;
      const double var = q2 / n - mean * mean;  // jax.nn.standardize: E[x^2] - E[x]^2
      stats[0] = (float)mean;
      stats[1] = (float)(1.0 / sqrt(var + 1e-5));
    }
  }
}

__global__ void standardize_inplace_kernel(float* __restrict__ adv, int64_t n,
                                           const float* __restrict__ stats) {
  const float mean = stats[0], rstd = stats[1];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride)
    adv[i] = (adv[i] - mean) * rstd;
}

int g_quads_override = 0;

template <class In>
int launch_gae(const In& in, int T, int E, bool vec4, int standardize, float* adv, float* tgt,
               float* stats, void* scratch, cudaStream_t st) {
  unsigned int* counter = reinterpret_cast<unsigned int*>(scratch);
  double2* partials = reinterpret_cast<double2*>(reinterpret_cast<char*>(scratch) + 16);
  const int want = standardize != 0;
  if (vec4 && (g_quads_override / 100 == 4 || (g_quads_override == 0 && E / 4 >= 65536))) {
    // several resident blocks per SM (code 4xx; the default for saturating shapes): 416 = 2 x 512 threads, 408 = 4 x 256
    const int quads = g_quads_override == 0 ? 16 : g_quads_override % 100, total = E / 4;
    const int grid = (total + quads - 1) / quads;
    if (quads == 16)
      gae_scan_kernel<In, 4, 16, 4, 2><<<grid, 16 * kChunks, 0, st>>>(in, T, E, adv, tgt, want, partials, counter, stats);
    else
      gae_scan_kernel<In, 4, 8, 4, 4><<<grid, 8 * kChunks, 0, st>>>(in, T, E, adv, tgt, want, partials, counter, stats);
  } else if (vec4 && g_quads_override / 100 == 3) {  // L = 2 timesteps per thread (fewer registers, more resident blocks): code 3xx
    const int quads = g_quads_override % 100, total = E / 4;
    const int grid = (total + quads - 1) / quads;
    if (quads == 32)
      gae_scan_kernel<In, 4, 32, 2><<<grid, 32 * kChunks, 0, st>>>(in, T, E, adv, tgt, want, partials, counter, stats);
    else if (quads == 16)
      gae_scan_kernel<In, 4, 16, 2><<<grid, 16 * kChunks, 0, st>>>(in, T, E, adv, tgt, want, partials, counter, stats);
    else
      gae_scan_kernel<In, 4, 8, 2><<<grid, 8 * kChunks, 0, st>>>(in, T, E, adv, tgt, want, partials, counter, stats);
  } else if (vec4 && g_quads_override / 100 == 2) {  // experimental float2 variants: tuning code 2xx
    const int groups = g_quads_override % 100, total = E / 2;
    const int grid = (total + groups - 1) / groups;
    if (groups == 32)
      gae_scan_kernel<In, 2, 32><<<grid, 32 * kChunks, 0, st>>>(in, T, E, adv, tgt, want, partials, counter, stats);
    else
      gae_scan_kernel<In, 2, 16><<<grid, 16 * kChunks, 0, st>>>(in, T, E, adv, tgt, want, partials, counter, stats);
  } else if (vec4) {
    const int quads_total = E / 4;
    // Small E is latency-bound: more, smaller blocks.  Large E: 128-byte rows per warp access.
    // measured on B200 (profiles/r01_gae_bench.txt): wider rows per block win once the grid is large
    int quads = quads_total >= 65536 ? 32 : (quads_total >= 8 * kNumSMs * 2 ? 8 : (quads_total >= 4 * kNumSMs ? 4 : 2));
    if (g_quads_override == 2 || g_quads_override == 4 || g_quads_override == 8 || g_quads_override == 16 ||
        g_quads_override == 32)
      quads = g_quads_override;
    const int grid = (quads_total + quads - 1) / quads;
    if (quads == 32)
      gae_scan_kernel<In, 4, 32><<<grid, 32 * kChunks, 0, st>>>(in, T, E, adv, tgt, want, partials, counter, stats);
    else if (quads == 16)
      gae_scan_kernel<In, 4, 16><<<grid, 16 * kChunks, 0, st>>>(in, T, E, adv, tgt, want, partials, counter, stats);
    else if (quads == 8)
      gae_scan_kernel<In, 4, 8><<<grid, 8 * kChunks, 0, st>>>(in, T, E, adv, tgt, want, partials, counter, stats);
    else if (quads == 4)
      gae_scan_kernel<In, 4, 4><<<grid, 4 * kChunks, 0, st>>>(in, T, E, adv, tgt, want, partials, counter, stats);
    else
      gae_scan_kernel<In, 4, 2><<<grid, 2 * kChunks, 0, st>>>(in, T, E, adv, tgt, want, partials, counter, stats);
  } else {
    const int grid = (E + 15) / 16;
    gae_scan_kernel<In, 1, 16><<<grid, 16 * kChunks, 0, st>>>(in, T, E, adv, tgt, want, partials, counter, stats);
  }
  STX_LAUNCH_OK();
  if (standardize == 2) {
    const int64_t n = (int64_t)T * E;
    const int grid = (int)((n + 1023) / 1024 < 4 * kNumSMs ? (n + 1023) / 1024 : 4 * kNumSMs);
    standardize_inplace_kernel<<<grid > 0 ? grid : 1, 256, 0, st>>>(adv, n, stats);
    STX_LAUNCH_OK();
  }
  return STX_OK;
}

}  // namespace
}  // namespace stx

using namespace stx;

extern "C" size_t stx_gae_scratch_bytes(int T, int E) {
  (void)T;
  const size_t blocks = (size_t)(E > 0 ? E : 1) / 8 + 2;  // smallest block covers 8 envs
  return 16 + blocks * sizeof(double2);
}

extern "C" void stx_gae_set_tuning(int quads) { g_quads_override = quads; }

static int check_common(int T, int E, int standardize, const void* adv, const void* tgt,
                        const void* stats, const void* scratch) {
  STX_REQUIRE(T > 0 && E > 0, STX_E_ARG, "stx_gae: T=%d E=%d must be positive", T, E);
  STX_REQUIRE(adv && tgt, STX_E_ARG, "stx_gae: null output");
  STX_REQUIRE(standardize >= 0 && standardize <= 2, STX_E_ARG, "stx_gae: standardize=%d", standardize);
  STX_REQUIRE(standardize == 0 || (stats && scratch), STX_E_ARG, "stx_gae: stats/scratch required when standardising");
  return STX_OK;
}

extern "C" int stx_gae_ppo_f32(const float* reward, const float* v_tm1, const float* v_t,
                               const uint8_t* done, const uint8_t* truncated, int T, int E,
                               float gamma, float lambda_, float reward_scale, int standardize,
                               float* adv, float* targets, float* stats, void* scratch,
                               void* stream) {
  if (int rc = check_common(T, E, standardize, adv, targets, stats, scratch)) return rc;
  STX_REQUIRE(reward && v_tm1 && v_t && done && truncated, STX_E_ARG, "stx_gae_ppo_f32: null input");
  const bool vec4 = (E % 4 == 0) && aligned16(reward) && aligned16(v_tm1) && aligned16(v_t) &&
                    aligned16(adv) && aligned16(targets) &&
                    (reinterpret_cast<uintptr_t>(done) % 4 == 0) &&
                    (reinterpret_cast<uintptr_t>(truncated) % 4 == 0);
  PpoIn in{reward, v_tm1, v_t, done, truncated, gamma, lambda_, reward_scale};
  return launch_gae(in, T, E, vec4, standardize, adv, targets, stats, scratch, (cudaStream_t)stream);
}

extern "C" int stx_gae_generic_f32(const float* r_t, const float* discount_t, const float* lambda_t,
                                   float lambda_, const float* v_tm1, const float* v_t,
                                   const float* truncation_t, int T, int E, int standardize,
                                   float* adv, float* targets, float* stats, void* scratch,
                                   void* stream) {
  if (int rc = check_common(T, E, standardize, adv, targets, stats, scratch)) return rc;
  STX_REQUIRE(r_t && discount_t && v_tm1 && v_t, STX_E_ARG, "stx_gae_generic_f32: null input");
  const bool vec4 = (E % 4 == 0) && aligned16(adv) && aligned16(targets);
  GenericIn in{r_t, discount_t, lambda_t, v_tm1, v_t, truncation_t, lambda_};
  return launch_gae(in, T, E, vec4, standardize, adv, targets, stats, scratch, (cudaStream_t)stream);
}
