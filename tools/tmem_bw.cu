// Micro-benchmark: tcgen05.ld throughput per SM as a function of the number of reading warps (diagnostic only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I stoix_b200/csrc tools/tmem_bw.cu -o tools/bin/tmem_bw
#include <cstdio>
#include <cuda_runtime.h>
#include "stx_tc_ptx.cuh"
using namespace stx::tc;

template <int MODE>  // 0: ld32 only, 1: ld32 + st16 (epilogue-like), 2: ld32 + 4x st.global.v4
__global__ void k(int iters, long long* out, uint32_t* sink, uint4* gbuf) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[32];
    const int c = (i + (warp >> 2)) & 7;
    tmem_ld32(tmem + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= r[j];
    if (MODE == 1) {
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = r[2 * j] + r[2 * j + 1];
      tmem_st16(tmem + 256 + c * 16, pk);
      tmem_st_wait();
    }
    if (MODE == 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        gbuf[((size_t)(blockIdx.x * 64 + (i & 63)) * 64 + (warp * 4 + j)) * 32 + lane] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
  long long* out;
  uint32_t* sink;
  uint4* gbuf;
  cudaMalloc(&out, 148 * 8);
  cudaMalloc(&sink, 4096);
  cudaMalloc(&gbuf, (size_t)148 * 64 * 64 * 32 * 16);  // [CTA][iteration & 63][store slot <= 16 warps x 4][lane] x 16 B
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(iters, out, sink, gbuf);
        if (mode == 1) k<1><<<148, warps * 32>>>(iters, out, sink, gbuf);
        if (mode == 2) k<2><<<148, warps * 32>>>(iters, out, sink, gbuf);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long h[148];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      const double bytes = (double)warps * iters * 4096.0;
      printf("mode %d warps %2d: %lld cycles, %.1f B/clk/SM TMEM read (%.1f cycles per warp-ld)\n", mode, warps, h[0], bytes / h[0],
             (double)h[0] / iters);
    }
  return 0;
}
