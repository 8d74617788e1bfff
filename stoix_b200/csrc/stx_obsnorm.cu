// Observation normalisation: running mean / std of the raw observations (batched Welford) + normalise-on-copy.
//
// Reference: stoix/utils/running_statistics.py -- update_statistics :204-345 (batched Welford with psum over the
// mapped axes), normalize :348-363 -- as used by the Anakin ff_ppo update step (stoix/systems/ppo/anakin/ff_ppo.py:
// 90-94, 113-115 normalise the inputs of both networks with the statistics from BEFORE the update; :145-162 absorb
// the raw trajectory observations with std limits 5e-4 / 5e4, psum over "device" and "batch").
//
// The reference runs a chain of XLA fusions with two psums per update; here it is
//   stx_running_stats_accumulate   one pass over the raw (rows, D) fp32 batch: per-feature S1 = sum w (x - mean_old),
//                                  S2 = sum w (x - mean_old)^2, W = sum w, accumulated in double, block partials
//                                  reduced in a fixed order by the last block (deterministic)           [HBM-bound: 4 B/elem]
//   (N > 1 ranks: ONE all-reduce of the 2D+1 doubles -- the reference's two psums collapse into it because
//    sum_dev sum (x-m)(x-m_new) = S2_tot - delta * S1_tot with delta = S1_tot / count_new)
//   stx_running_stats_finalize     count += W; delta = S1/count; mean += delta; summed_variance += S2 - delta*S1;
//                                  std = clip(sqrt(clip(max(sv,0)/count, min^2, max^2)), min, max)         [D threads]
//   stx_obs_normalize              out = (x - mean) / std (optional symmetric clip), fp32 or bf16 output [6-8 B/elem]
#include "stx_common.cuh"

namespace stx {
namespace {

constexpr int kMaxD = 256;

struct AccScratch {
  unsigned int ticket;
  unsigned int pad;
  // followed by double partials[grid][2*D + 1]
};

// blockDim.x = L * rpp with L = D / V lanes per row (V = 4 floats per lane when D % 4 == 0, else 1) and rpp rows per
// pass; thread (r, l) walks rows r, r + rpp*grid, ... of its V features: consecutive threads read consecutive 4V bytes
// (fully coalesced), kUnroll independent row loads in flight per thread.
template <int V>
__global__ void __launch_bounds__(256) stats_accumulate_kernel(const float* __restrict__ x, const float* __restrict__ w, int64_t rows, int D,
                                                             const float* __restrict__ mean, double* __restrict__ sums, AccScratch* scratch) {
  extern __shared__ double sm[];  // [rpp][2*D] + [rpp]
  constexpr int kUnroll = 8;
  const int L = D / V;
  const int rpp = blockDim.x / L;
  const int l = threadIdx.x % L, r = threadIdx.x / L;
  float m[V];
#pragma unroll
  for (int k = 0; k < V; ++k) m[k] = mean[l * V + k];
  double s1[V], s2[V], sw = 0.0;
#pragma unroll
  for (int k = 0; k < V; ++k) s1[k] = 0.0, s2[k] = 0.0;
  const int64_t step = (int64_t)gridDim.x * rpp;
  for (int64_t row0 = (int64_t)blockIdx.x * rpp + r; row0 < rows; row0 += step * kUnroll) {
    float v[kUnroll][V], wt[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t row = row0 + u * step;
      const bool ok = row < rows;
      wt[u] = ok ? (w ? w[row] : 1.0f) : 0.0f;
      if (V == 4) {
        const float4 t = ok ? ldg_stream4(x + row * D + l * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[u][0] = t.x, v[u][1 % V] = t.y, v[u][2 % V] = t.z, v[u][3 % V] = t.w;
      } else {
        v[u][0] = ok ? __ldg(x + row * D + l) : 0.f;
      }
    }
    // the kUnroll rows of one pass are summed in fp32 (8 terms: ~1e-7 relative), the running sums in double: the fp64
    // pipe and the fp32->fp64 conversions see 1/8 of the elements (the all-double version ran at 24 % of HBM peak)
    float f1[V], f2[V], fw = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) f1[k] = 0.f, f2[k] = 0.f;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float diff = v[u][k] - m[k];
        const float wd = wt[u] * diff;
        f1[k] += wd;
        f2[k] = fmaf(wd, diff, f2[k]);
      }
      fw += wt[u];
    }
#pragma unroll
    for (int k = 0; k < V; ++k) s1[k] += (double)f1[k], s2[k] += (double)f2[k];
    sw += (double)fw;
  }
#pragma unroll
  for (int k = 0; k < V; ++k) {
    sm[(r * 2 + 0) * D + l * V + k] = s1[k];
    sm[(r * 2 + 1) * D + l * V + k] = s2[k];
  }
  if (l == 0) sm[2 * D * rpp + r] = sw;
  __syncthreads();
  double* partials = reinterpret_cast<double*>(scratch + 1);
  const int n = 2 * D + 1;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = 0.0;
    if (i < 2 * D) {
      const int which = i / D, dd = i % D;
      for (int rr = 0; rr < rpp; ++rr) acc += sm[(rr * 2 + which) * D + dd];
    } else {
      for (int rr = 0; rr < rpp; ++rr) acc += sm[2 * D * rpp + rr];
    }
    partials[(int64_t)blockIdx.x * n + i] = acc;
  }
  // last block: fixed-order sum over the block partials
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&scratch->ticket, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) scratch->ticket = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) acc += __ldcg(&partials[(int64_t)b * n + i]);
    sums[i] = acc;
  }
}

__global__ void stats_finalize_kernel(const double* __restrict__ sums, int D, int64_t* __restrict__ count,
                                      float* __restrict__ mean, float* __restrict__ summed_variance, float* __restrict__ stdv,
                                      float std_min, float std_max) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const double inc = sums[2 * D];  // summed weight of the batch (all ranks)
  const double cnt = (double)count[0] + inc;
  if (d < D && cnt > 0.0) {
    const double s1 = sums[d], s2 = sums[D + d];
    const double delta = s1 / cnt;                                  // mean_update (:301-302)
    const double mn = (double)mean[d] + delta;
    const double sv = (double)summed_variance[d] + (s2 - delta * s1);  // sum diff_old * diff_new (:305-309)
    double var = fmax(sv, 0.0) / cnt;                               // :333-337
    var = fmin(fmax(var, (double)std_min * std_min), (double)std_max * std_max);
    double sd = sqrt(var);
    sd = fmin(fmax(sd, (double)std_min), (double)std_max);          // :338-339
    mean[d] = (float)mn, summed_variance[d] = (float)sv, stdv[d] = (float)sd;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) count[0] = (int64_t)llrint(cnt);
}

template <bool BF16>
__global__ void __launch_bounds__(256) obs_normalize_kernel(const float* __restrict__ x, int64_t n4, int D, const float* __restrict__ mean,
                                                           const float* __restrict__ stdv, float max_abs, void* __restrict__ out) {
  // D % 4 == 0: every float4 stays inside one row; feature index of element 4*i is (4*i) % D
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = ldg_stream4(x + 4 * i);
    const int d = (int)((4 * i) % D);
    const float4 m = *reinterpret_cast<const float4*>(mean + d), s = *reinterpret_cast<const float4*>(stdv + d);
    float y[4] = {(v.x - m.x) / s.x, (v.y - m.y) / s.y, (v.z - m.z) / s.z, (v.w - m.w) / s.w};
    if (max_abs > 0.f) {
#pragma unroll
      for (int k = 0; k < 4; ++k) y[k] = fminf(fmaxf(y[k], -max_abs), max_abs);
    }
    if (BF16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(y[0], y[1]), hi = __floats2bfloat162_rn(y[2], y[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo), pk.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(out)[i] = pk;
    } else {
      stg_stream4(reinterpret_cast<float*>(out) + 4 * i, make_float4(y[0], y[1], y[2], y[3]));
    }
  }
}

template <bool BF16>
__global__ void __launch_bounds__(256) obs_normalize_scalar_kernel(const float* __restrict__ x, int64_t n, int D, const float* __restrict__ mean,
                                                                  const float* __restrict__ stdv, float max_abs, void* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const int d = (int)(i % D);
    float y = (x[i] - mean[d]) / stdv[d];
    if (max_abs > 0.f) y = fminf(fmaxf(y, -max_abs), max_abs);
    if (BF16) reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(y);
    else reinterpret_cast<float*>(out)[i] = y;
  }
}

int acc_grid(int64_t rows, int rpp) {
  int64_t blocks = (rows + rpp - 1) / rpp;
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;  // 4 resident blocks of <= 256 threads per SM: one wave
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace
}  // namespace stx

using namespace stx;

extern "C" size_t stx_running_stats_scratch_bytes(int D) { return sizeof(AccScratch) + sizeof(double) * (size_t)(4 * kNumSMs) * (2 * (size_t)D + 1); }

// sums: [2*D + 1] doubles (S1[D], S2[D], W).  scratch: stx_running_stats_scratch_bytes(D), zero-initialised once.
extern "C" int stx_running_stats_accumulate(const float* x, const float* weights, int64_t rows, int D, const float* mean, double* sums,
                                            void* scratch, void* stream) {
  STX_REQUIRE(x && mean && sums && scratch, STX_E_ARG, "stx_running_stats_accumulate: null pointer");
  STX_REQUIRE(D >= 1 && D <= kMaxD && rows >= 0, STX_E_SHAPE, "stx_running_stats_accumulate: feature dim %d unsupported (1..%d)", D, kMaxD);
  const bool vec = D % 4 == 0 && aligned16(x);
  const int L = vec ? D / 4 : D;
  const int rpp = 256 / L > 0 ? 256 / L : 1;
  const int threads = L * rpp;
  const size_t smem = sizeof(double) * ((size_t)2 * D * rpp + rpp);
  STX_REQUIRE(smem <= 48 * 1024, STX_E_SHAPE, "stx_running_stats_accumulate: feature dim %d needs %zu B of shared memory", D, smem);
  if (vec)
    stats_accumulate_kernel<4><<<acc_grid(rows, rpp), threads, smem, (cudaStream_t)stream>>>(x, weights, rows, D, mean, sums,
                                                                                            reinterpret_cast<AccScratch*>(scratch));
  else
    stats_accumulate_kernel<1><<<acc_grid(rows, rpp), threads, smem, (cudaStream_t)stream>>>(x, weights, rows, D, mean, sums,
                                                                                            reinterpret_cast<AccScratch*>(scratch));
  STX_LAUNCH_OK();
  return STX_OK;
}

// sums = the (all-reduced) output of stx_running_stats_accumulate; the state tensors are updated in place.
extern "C" int stx_running_stats_finalize(const double* sums, int D, int64_t* count, float* mean, float* summed_variance, float* std,
                                          float std_min_value, float std_max_value, void* stream) {
  STX_REQUIRE(sums && count && mean && summed_variance && std, STX_E_ARG, "stx_running_stats_finalize: null pointer");
  STX_REQUIRE(D >= 1 && D <= kMaxD, STX_E_SHAPE, "stx_running_stats_finalize: feature dim %d unsupported (1..%d)", D, kMaxD);
  STX_REQUIRE(std_min_value > 0.f && std_max_value >= std_min_value, STX_E_ARG, "stx_running_stats_finalize: bad std limits");
  stats_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(sums, D, count, mean, summed_variance, std, std_min_value, std_max_value);
  STX_LAUNCH_OK();
  return STX_OK;
}

extern "C" int stx_obs_normalize(const float* x, int64_t rows, int D, const float* mean, const float* std, float max_abs_value, void* out,
                                 int out_bf16, void* stream) {
  STX_REQUIRE(x && mean && std && out, STX_E_ARG, "stx_obs_normalize: null pointer");
  STX_REQUIRE(D >= 1 && rows >= 0, STX_E_SHAPE, "stx_obs_normalize: bad shape");
  const int64_t n = rows * D;
  if (n == 0) return STX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = D % 4 == 0 && aligned16(x) && aligned16(mean) && aligned16(std) && (reinterpret_cast<uintptr_t>(out) & (out_bf16 ? 7u : 15u)) == 0;
  if (vec) {
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
    if (out_bf16) obs_normalize_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(x, n4, D, mean, std, max_abs_value, out);
    else obs_normalize_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(x, n4, D, mean, std, max_abs_value, out);
  } else {
    int64_t blocks = (n + 255) / 256;
    if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
    if (out_bf16) obs_normalize_scalar_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(x, n, D, mean, std, max_abs_value, out);
    else obs_normalize_scalar_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(x, n, D, mean, std, max_abs_value, out);
  }
  STX_LAUNCH_OK();
  return STX_OK;
}
