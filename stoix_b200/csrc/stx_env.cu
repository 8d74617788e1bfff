// Synthetic Box environment step (the named benchmark env, BASELINE.json configs[1] / SURVEY.md 8d)
// and the keyed-bijection minibatch shuffle.
//
// Environment contract restated from the reference's call sites (stoa is not vendored):
//   * AutoResetWrapper(next_obs_in_extras=True)  stoix/utils/make_env.py:56-60: on the last step of
//     an episode timestep.observation is the RESET observation and extras["next_obs"] the true
//     successor (ff_ppo.py:110-116);
//   * discount == 0 <=> termination; last() & discount != 0 <=> truncation (ff_ppo.py:107-108);
//   * RecordEpisodeMetrics: running return/length, published with is_terminal_step on the final step
//     (same logic spelled out in stoix/wrappers/envpool.py:94-133).
// One thread per (env, 4-observation-feature group): Philox4x32-10 keyed by seed, counter =
// (env, step, lane, stream-tag), so the trajectory is a pure function of (seed, step, env) and the
// CPU oracle in tests/ regenerates it exactly.
#include "stx_common.cuh"

namespace stx {
namespace {

constexpr uint32_t kTagObs = 0x4f425331u;    // successor observation
constexpr uint32_t kTagReset = 0x52535431u;  // reset observation
constexpr uint32_t kTagStep = 0x53545031u;   // reward / termination / truncation

template <bool BF16>
__device__ __forceinline__ void store4(void* base, int64_t off, int n_valid, const float* v) {
  if (BF16) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(base) + off;
    for (int k = 0; k < n_valid; ++k) p[k] = __float2bfloat16_rn(v[k]);
  } else {
    float* p = reinterpret_cast<float*>(base) + off;
    for (int k = 0; k < n_valid; ++k) p[k] = v[k];
  }
}

template <bool BF16>
__global__ void synth_env_step_kernel(int64_t E, int D, uint64_t seed, uint64_t step_base,
                                      const uint64_t* __restrict__ dev_counter, float p_term,
                                      float p_trunc, void* __restrict__ obs_out,
                                      void* __restrict__ next_obs, float* __restrict__ reward,
                                      uint8_t* __restrict__ done, uint8_t* __restrict__ truncated,
                                      float* __restrict__ run_return, int32_t* __restrict__ run_length,
                                      float* __restrict__ ep_return, int32_t* __restrict__ ep_length,
                                      uint8_t* __restrict__ is_terminal) {
  const int groups = (D + 3) / 4;
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (gid >= E * groups) return;
  const int64_t e = gid / groups;
  const int grp = (int)(gid % groups);
  const uint64_t step = step_base + (dev_counter ? *dev_counter : 0ull);
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t e_lo = (uint32_t)e, st_lo = (uint32_t)step;
  const uint32_t hi = ((uint32_t)((uint64_t)e >> 32) << 16) ^ (uint32_t)(step >> 32);

  // step outcome (recomputed by every feature-group thread of the env: 1 Philox call, no sync needed)
  const uint4 rs = Philox::rand4(make_uint4(e_lo, st_lo, hi, kTagStep), key);
  const float2 nr = normal2(rs.x, rs.y);
  const bool term = u01(rs.z) < p_term;
  const bool trunc = !term && (u01(rs.w) < p_trunc);
  const bool last = term || trunc;

  const uint4 ro = Philox::rand4(make_uint4(e_lo, st_lo, hi ^ ((uint32_t)grp << 8), kTagObs), key);
  const float2 a = normal2(ro.x, ro.y), b = normal2(ro.z, ro.w);
  const float nxt[4] = {a.x, a.y, b.x, b.y};
  const int nv = D - grp * 4 < 4 ? D - grp * 4 : 4;
  const int64_t off = e * D + grp * 4;
  store4<BF16>(next_obs, off, nv, nxt);
  if (last) {
    const uint4 rr = Philox::rand4(make_uint4(e_lo, st_lo, hi ^ ((uint32_t)grp << 8), kTagReset), key);
    const float2 c = normal2(rr.x, rr.y), d2 = normal2(rr.z, rr.w);
    const float rst[4] = {c.x, c.y, d2.x, d2.y};
    store4<BF16>(obs_out, off, nv, rst);
  } else {
    store4<BF16>(obs_out, off, nv, nxt);
  }
  if (grp == 0) {
    const float r = nr.x;
    reward[e] = r;
    done[e] = term ? 1 : 0;
    truncated[e] = trunc ? 1 : 0;
    const float ret = run_return[e] + r;
    const int32_t len = run_length[e] + 1;
    ep_return[e] = ret;   // running totals; the finished episode's totals when is_terminal
    ep_length[e] = len;
    is_terminal[e] = last ? 1 : 0;
    run_return[e] = last ? 0.f : ret;
    run_length[e] = last ? 0 : len;
  }
}

// ---- keyed bijection on [0, n): cycle-walking 6-round Feistel over the next power of 4 ------
// Round function: a 32-bit avalanche mixer (two multiply / xor-shift stages) of (right half + round key); the six
// round keys come from one Philox draw per thread.  (A Philox block per round made this shuffle a 36 us compute-bound
// kernel; a shuffle needs a keyed bijection with good diffusion, not a cryptographic round function.)
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x21f0aaadu;
  x ^= x >> 15;
  x *= 0x735a2d97u;
  x ^= x >> 15;
  return x;
}
__device__ __forceinline__ uint32_t feistel(uint32_t x, int half_bits, const uint32_t (&rk)[6]) {
  const uint32_t mask = (1u << half_bits) - 1u;
  uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
  for (int round = 0; round < 6; ++round) {
    const uint32_t nl = r, nr = l ^ (mix32(r + rk[round]) & mask);
    l = nl, r = nr;
  }
  return (l << half_bits) | r;
}

__global__ void permutation_kernel(int32_t* __restrict__ perm, int64_t n, int half_bits, uint64_t seed,
                                   uint64_t stream_base, const uint64_t* __restrict__ dev_counter) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t sid = stream_base + (dev_counter ? *dev_counter : 0ull);
  const uint64_t k = seed ^ (sid * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull);
  const uint2 key = make_uint2((uint32_t)k, (uint32_t)(k >> 32));
  const uint4 k0 = Philox::rand4(make_uint4(0u, 0u, 0x5045524du, 0u), key), k1 = Philox::rand4(make_uint4(1u, 0u, 0x5045524du, 0u), key);
  const uint32_t rk[6] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y};
  uint32_t x = (uint32_t)i;
  do {
    x = feistel(x, half_bits, rk);  // a bijection on [0, 4^half_bits); walk the cycle back into [0, n)
  } while (x >= (uint64_t)n);
  perm[i] = (int32_t)x;
}

}  // namespace
}  // namespace stx

using namespace stx;

__global__ void counter_add_kernel(uint64_t* c, uint64_t inc) { *c += inc; }

extern "C" int stx_counter_add(uint64_t* counter, uint64_t inc, void* stream) {
  STX_REQUIRE(counter, STX_E_ARG, "stx_counter_add: null counter");
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, inc);
  STX_LAUNCH_OK();
  return STX_OK;
}

extern "C" int stx_synth_env_step(int64_t E, int D, uint64_t seed, uint64_t step,
                                  const uint64_t* dev_counter, float p_term,
                                  float p_trunc, const int32_t* action, void* obs_out,
                                  void* next_obs, int obs_bf16, float* reward, uint8_t* done,
                                  uint8_t* truncated, float* run_return, int32_t* run_length,
                                  float* ep_return, int32_t* ep_length, uint8_t* is_terminal,
                                  void* stream) {
  (void)action;  // the synthetic dynamics ignore the action (it only fixes shapes and bandwidth)
  STX_REQUIRE(E > 0 && D > 0, STX_E_SHAPE, "stx_synth_env_step: E=%lld D=%d", (long long)E, D);
  STX_REQUIRE(obs_out && next_obs && reward && done && truncated && run_return && run_length &&
                  ep_return && ep_length && is_terminal,
              STX_E_ARG, "stx_synth_env_step: null pointer");
  const int groups = (D + 3) / 4;
  const int64_t total = E * groups;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (obs_bf16)
    synth_env_step_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(E, D, seed, step, dev_counter, p_term, p_trunc, obs_out, next_obs, reward, done, truncated, run_return, run_length, ep_return, ep_length, is_terminal);
  else
    synth_env_step_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(E, D, seed, step, dev_counter, p_term, p_trunc, obs_out, next_obs, reward, done, truncated, run_return, run_length, ep_return, ep_length, is_terminal);
  STX_LAUNCH_OK();
  return STX_OK;
}

extern "C" int stx_make_permutation(int32_t* perm, int64_t n, uint64_t seed, uint64_t stream_id,
                                    const uint64_t* dev_counter, void* stream) {
  STX_REQUIRE(perm && n > 0 && n < (1ll << 31), STX_E_ARG, "stx_make_permutation: n=%lld", (long long)n);
  int half_bits = 1;
  while ((1ll << (2 * half_bits)) < n) ++half_bits;
  permutation_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(perm, n, half_bits, seed, stream_id, dev_counter);
  STX_LAUNCH_OK();
  return STX_OK;
}
