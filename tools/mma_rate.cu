// Micro-benchmark: tcgen05.mma issue-to-completion rate per SM for the shapes K3a uses (diagnostic only).
//   kind::f16 bf16, M=128, K=16 per instruction; N in {64,128,256}; A from TMEM (TS) or smem (SS).
#include <cstdio>
#include <cuda_runtime.h>
#include "stx_tc_ptx.cuh"
using namespace stx::tc;

template <int N, bool TS, bool WITH_LD>
__global__ void k(int n_mma, long long* out, uint32_t* sink) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&slot, 512);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_async_proxy();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot, sbase = smem_u32(smem);
  constexpr uint32_t idesc = idesc_bf16(128, N, 0, 1);
  long long t0 = 0, t1 = 0;
  uint32_t acc = 0;
  if (warp == 0) {
    t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < n_mma; ++i) {
        const int k = i & 15, pt = (i >> 4) % (256 / N);
        if (TS)
          mma_ts(tmem + pt * N, tmem + 256 + k * 8, smem_desc(sbase + pt * (N / 64) * 32768 + k * 2048, 32768, 1024, SWIZZLE_128B), idesc, k > 0);
        else
          mma_ss(tmem + pt * N, smem_desc(sbase + 131072 + (k & 3) * 32, 16, 1024, SWIZZLE_128B),
                 smem_desc(sbase + pt * (N / 64) * 32768 + k * 2048, 32768, 1024, SWIZZLE_128B), idesc, k > 0);
      }
      mma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0, 1);
    t1 = clock64();
  } else if (WITH_LD && warp >= 4) {
    // epilogue-like TMEM traffic from 8 warps while the MMAs run: ld32 of D + st16 into the A region
    const uint32_t t = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    for (int i = 0; i < n_mma / 4; ++i) {
      uint32_t r[32], pk[16];
      tmem_ld32(t + ((i + (warp >> 2)) & 7) * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = r[2 * j] ^ r[2 * j + 1];
      tmem_st16(t + 384 + (i & 7) * 16, pk);
      tmem_st_wait();
      acc ^= pk[3];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345u) sink[threadIdx.x] = acc;
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int N, bool TS, bool LD>
void run(const char* name, long long* out, uint32_t* sink) {
  const int n = 1024;
  cudaFuncSetAttribute(k<N, TS, LD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    k<N, TS, LD><<<148, 384, 170 * 1024>>>(n, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(e)); exit(1); }
  }
  long long h[148];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-28s N=%3d: %.1f cycles per MMA (nominal %d), %.0f MAC/clk/SM\n", name, N, (double)h[0] / n, N / 2, 128.0 * N * 16 * n / h[0]);
}

int main() {
  long long* out;
  uint32_t* sink;
  cudaMalloc(&out, 148 * 8);
  cudaMalloc(&sink, 4096);
  run<64, true, false>("TS (A in TMEM)", out, sink);
  run<128, true, false>("TS (A in TMEM)", out, sink);
  run<256, true, false>("TS (A in TMEM)", out, sink);
  run<64, false, false>("SS (A in smem)", out, sink);
  run<128, false, false>("SS (A in smem)", out, sink);
  run<256, false, false>("SS (A in smem)", out, sink);
  run<64, true, true>("TS + 8 warps ld/st TMEM", out, sink);
  run<128, true, true>("TS + 8 warps ld/st TMEM", out, sink);
  run<256, true, true>("TS + 8 warps ld/st TMEM", out, sink);
  return 0;
}
