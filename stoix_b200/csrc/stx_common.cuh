// Shared device/host helpers for libstoixb200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/stx.h"

#ifndef __CUDA_ARCH__
#define STX_HOST_ONLY 1
#endif

namespace stx {

// acc.{x,y} += w * v.{x,y} as ONE instruction (sm_100 FFMA2 with a scalar-broadcast operand): two IEEE fused multiply-adds, the
// same bits as two FFMAs, half the issue slots -- what the issue-bound fp32 inner loops (small GEMMs, recurrent products) need.
__device__ __forceinline__ void fma2(float2& acc, float w, float2 v) { acc = __ffma2_rn(make_float2(w, w), v, acc); }


// ---- error plumbing -----------------------------------------------------------------------
void set_error(const char* fmt, ...);  // defined in stx_api.cu (thread-local buffer)

#define STX_REQUIRE(cond, code, ...)            \
  do {                                          \
    if (!(cond)) {                              \
      ::stx::set_error(__VA_ARGS__);            \
      return (code);                            \
    }                                           \
  } while (0)

#define STX_CUDA_OK(expr)                                                         \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      ::stx::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                       __FILE__, __LINE__);                                       \
      return (int)_e;                                                             \
    }                                                                             \
  } while (0)

// Checks the launch that just happened (no sync; capture-safe) and counts it (stx_launch_count()).
void count_launch();
#define STX_LAUNCH_OK()                  \
  do {                                   \
    ::stx::count_launch();               \
    STX_CUDA_OK(cudaPeekAtLastError());  \
  } while (0)

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// The optimiser step is a strict chain of short kernels (K3a -> K3b -> reduce -> K4, 64 times per update).  A kernel
// launched through launch_pdl() may become resident while its stream predecessor is still running (its prologue --
// barrier init, TMEM allocation, smem clearing -- overlaps the predecessor's tail and the launch latency is hidden);
// it MUST execute griddep_wait() before touching any global memory the predecessor chain reads or writes.
// STX_PDL=0 in the environment launches the same kernels fully serialised (then griddep_wait() is a no-op).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr, cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ---- L2 residency hints ---------------------------------------------------------------------------
// One optimiser step streams ~280 MB of activations (K3a writes, K3b reads) through the 126 MB L2 and would evict the
// few MB that the latency-bound tail of the step (split-K partials, gradients, parameters, Adam moments) needs: those
// kernels are chains of dependent round trips, so an L2 hit instead of a DRAM access shortens every link.  Streaming
// data is tagged evict_first, the small reused set evict_last.
#ifdef __CUDACC__
__device__ __forceinline__ uint64_t l2_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy(int mode) {  // 0 evict_first, 1 evict_normal, 2 evict_last
  uint64_t p;
  if (mode == 0) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else if (mode == 1) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void st_hint(uint4* ptr, uint4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void st_hint(float4* ptr, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void st_hint(uint2* ptr, uint2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1, %2}, %3;" ::"l"(ptr), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(float* ptr, float v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(ptr), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ float ld_hint(const float* ptr, uint64_t pol) {
  float v;
  asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(ptr), "l"(pol));
  return v;
}
__device__ __forceinline__ float4 ld_hint(const float4* ptr, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(pol));
  return v;
}
#endif

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- Philox4x32-10 (counter-based RNG; Salmon et al. 2011) ----------------------------------
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
    hi = __umulhi(a, b);
    lo = a * b;
#else
    uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#endif
  }
  // counter = (c0..c3), key = (k0,k1) -> 4 random words
  __host__ __device__ static inline uint4 rand4(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(M0, c.x, hi0, lo0);
      mulhilo(M1, c.z, hi1, lo1);
      c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
      k.x += W0;
      k.y += W1;
    }
    return c;
  }
};

// uniform in (0,1): never 0 or 1.  23 random bits + half-step offset: k + 0.5 is exact in fp32 for every k < 2^23
// (with 24 bits the odd k >= 2^23 round half-to-even and k = 2^24 - 1 yields exactly 1.0, i.e. -log(-log(u)) = +inf
// in the Gumbel-max sampler).
__host__ __device__ static inline float u01(uint32_t x) { return ((x >> 9) + 0.5f) * (1.0f / 8388608.0f); }

#ifdef __CUDACC__
// two N(0,1) from two words (Box-Muller, full-precision log/sincos: this is a synthetic data source
// that the CPU test oracle re-computes, so it must not depend on -use_fast_math intrinsics).
__device__ static inline float2 normal2(uint32_t a, uint32_t b) {
  float r = sqrtf(-2.0f * logf(u01(a)));
  float s, c;
  sincospif(2.0f * u01(b), &s, &c);
  return make_float2(r * c, r * s);
}

// Bounded spin for cross-block / cross-GPU flag waits: a protocol error traps (with a message) after ~4 s instead of
// hanging the GPU.  `spins` is the caller's loop counter; the wall clock is only consulted every 1024 polls.
__device__ __forceinline__ void spin_guard(unsigned int& spins, unsigned long long& t0, const char* what) {
  if ((++spins & 1023u) != 0u) return;
  unsigned long long now;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
  if (t0 == 0ull) t0 = now;
  else if (now - t0 > 4000000000ull) {
    printf("[stx] spin-wait timeout (%s): block %d thread %d\n", what, blockIdx.x, threadIdx.x);
    __trap();
  }
}

// ---- deterministic block reductions ---------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the block; result valid in thread 0. `sm` needs 32 elements. Fixed order -> deterministic.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* sm) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  T r = T(0);
  if (w == 0) {
    r = lane < nw ? sm[lane] : T(0);
    r = warp_sum(r);
  }
  return r;
}

// "Last block done" ticket: returns true in exactly one block (the last to arrive), for all of its
// threads, after every other block's prior global writes are visible. Resets the counter for reuse.
__device__ __forceinline__ bool last_block_ticket(unsigned int* counter, unsigned int nblocks) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == nblocks - 1);
    if (is_last) *counter = 0u;
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float2 ldg_stream2(const float* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream2(float* p, float2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ uint32_t ldg_stream_u16(const void* p) {
  uint16_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
#endif  // __CUDACC__

}  // namespace stx
