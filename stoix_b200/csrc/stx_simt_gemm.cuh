// fp32 CUDA-core GEMM family for the parity path (STX_PREC_F32).
//
// The reference is fp32 everywhere (SURVEY.md "facts"), so this path keeps every product and
// accumulation in fp32 and is what the tight-tolerance parity tests run.  The bf16 tcgen05 path
// (stx_tc_*.cu) is the throughput path.  Three flavours of one 64x64x16 register-tiled kernel:
//   FWD : Y[m,n]  = act(sum_k X[row(m),k] * W[k,n] + b[n])                 nn.Dense, torso.py:26
//   DX  : dX[m,k] = (sum_n dY[m,n] * W[k,n]) * (H[m,k] > 0)                 backward through relu
//   DW  : dWp[z][k,n] = sum_{m in slice z} X[row(m),k] * dY[m,n],  dbp[z][n] = sum dY[m,n]
// All shapes are bounds-checked (D=4, A=2 of CartPole work).  Split-M partials of DW are reduced in
// a fixed order by reduce_partials_kernel -> run-to-run deterministic gradients.
#pragma once
#include "stx_common.cuh"

namespace stx {
namespace simt {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
constexpr int kThreads = (BM / TM) * (BN / TN);  // 256

enum Mode { FWD = 0, DX = 1, DW = 2 };

struct GemmArgs {
  // FWD: A=X (M x K, lda, optional row gather), B=W (K x N), C=Y (M x N)
  // DX : A=dY (M x Nr, lda=Nr), B=W (Kout x Nr) used transposed, C=dX (M x Kout), mask=H (M x Kout)
  // DW : A=X (rows x Kout, gathered), B=dY (rows x N), C=partials [z][Kout x N]
  const float* A;
  const float* B;
  float* C;
  const float* bias;     // FWD
  const float* mask;     // DX: post-activation tensor of the producing layer (ld = ldc)
  const int32_t* rowidx; // FWD/DW: gather index for A rows (nullable)
  float* dbias;          // DW: partial db [z][N]
  int64_t M;             // FWD/DX: output rows.  DW: total sample rows (reduction length)
  int N;                 // output columns
  int K;                 // FWD: reduction (in dim).  DX: reduction (= layer out dim).  DW: output rows (in dim)
  int64_t lda;
  int relu;              // FWD epilogue
  int64_t rows_per_split;  // DW
  int64_t part_stride;   // DW: floats between split partials of this layer
  int64_t dbias_stride;  // DW
};

template <int MODE>
__global__ void __launch_bounds__(kThreads) gemm_kernel(GemmArgs g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int64_t m0 = (int64_t)blockIdx.y * BM;  // output row tile
  const int n0 = blockIdx.x * BN;               // output col tile
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float dbacc[TN] = {0.f, 0.f, 0.f, 0.f};

  // reduction range
  int64_t r_begin = 0, r_end;
  if (MODE == FWD) r_end = g.K;
  else if (MODE == DX) r_end = g.K;
  else {
    r_begin = (int64_t)blockIdx.z * g.rows_per_split;
    r_end = r_begin + g.rows_per_split;
    if (r_end > g.M) r_end = g.M;
  }
  const int out_rows = (MODE == DW) ? g.K : 0;  // DW: output rows = in-dim

  for (int64_t r0 = r_begin; r0 < r_end; r0 += BK) {
    // ---- load A tile into As[kk][m] ----
    if (MODE == FWD) {
      // X[row(m), r0+kk]: kk fastest in memory
      for (int i = tid; i < BM * BK; i += kThreads) {
        const int kk = i % BK, mm = i / BK;
        const int64_t m = m0 + mm, k = r0 + kk;
        float v = 0.f;
        if (m < g.M && k < r_end) {
          const int64_t row = g.rowidx ? (int64_t)g.rowidx[m] : m;
          v = g.A[row * g.lda + k];
        }
        As[kk][mm] = v;
      }
    } else if (MODE == DX) {
      // dY[m, r0+kk]
      for (int i = tid; i < BM * BK; i += kThreads) {
        const int kk = i % BK, mm = i / BK;
        const int64_t m = m0 + mm, k = r0 + kk;
        As[kk][mm] = (m < g.M && k < r_end) ? g.A[m * g.lda + k] : 0.f;
      }
    } else {
      // DW: A(i, r) = X[row(r), i]; i (feature) fastest in memory. As[kk=r][mm=i]
      for (int i = tid; i < BM * BK; i += kThreads) {
        const int mm = i % BM, kk = i / BM;
        const int64_t r = r0 + kk, f = m0 + mm;
        float v = 0.f;
        if (r < r_end && f < out_rows) {
          const int64_t row = g.rowidx ? (int64_t)g.rowidx[r] : r;
          v = g.A[row * g.lda + f];
        }
        As[kk][mm] = v;
      }
    }
    // ---- load B tile into Bs[kk][n] ----
    if (MODE == FWD) {
      for (int i = tid; i < BN * BK; i += kThreads) {
        const int nn = i % BN, kk = i / BN;
        const int64_t k = r0 + kk;
        const int n = n0 + nn;
        Bs[kk][nn] = (k < r_end && n < g.N) ? g.B[k * g.N + n] : 0.f;
      }
    } else if (MODE == DX) {
      // B(kk, j) = W[j, r0+kk], W is (Kout x Nr) row-major; output col j = n0+nn; reduction fastest
      for (int i = tid; i < BN * BK; i += kThreads) {
        const int kk = i % BK, nn = i / BK;
        const int64_t k = r0 + kk;
        const int j = n0 + nn;
        Bs[kk][nn] = (k < r_end && j < g.N) ? g.B[(int64_t)j * g.K + k] : 0.f;
      }
    } else {
      for (int i = tid; i < BN * BK; i += kThreads) {
        const int nn = i % BN, kk = i / BN;
        const int64_t r = r0 + kk;
        const int n = n0 + nn;
        Bs[kk][nn] = (r < r_end && n < g.N) ? g.B[r * g.N + n] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (MODE == DW) {
#pragma unroll
        for (int j = 0; j < TN; ++j) dbacc[j] += b[j];
      }
    }
    __syncthreads();
  }

  // ---- epilogue ----
  if (MODE == FWD) {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int64_t m = m0 + ty * TM + i;
      if (m >= g.M) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * TN + j;
        if (n >= g.N) continue;
        float v = acc[i][j] + (g.bias ? g.bias[n] : 0.f);
        if (g.relu) v = fmaxf(v, 0.f);
        g.C[m * g.N + n] = v;
      }
    }
  } else if (MODE == DX) {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int64_t m = m0 + ty * TM + i;
      if (m >= g.M) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * TN + j;
        if (n >= g.N) continue;
        const float h = g.mask ? g.mask[m * g.N + n] : 1.f;
        g.C[m * g.N + n] = h > 0.f ? acc[i][j] : 0.f;
      }
    }
  } else {
    float* Cp = g.C + (int64_t)blockIdx.z * g.part_stride;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int64_t f = m0 + ty * TM + i;
      if (f >= out_rows) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * TN + j;
        if (n < g.N) Cp[f * g.N + n] = acc[i][j];
      }
    }
    if (blockIdx.y == 0 && ty == 0 && g.dbias) {
      float* dbp = g.dbias + (int64_t)blockIdx.z * g.dbias_stride;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * TN + j;
        if (n < g.N) dbp[n] = dbacc[j];
      }
    }
  }
}

// grad[i] += w * sum_z part[z*stride + i]   (fixed order over z -> deterministic)
__global__ void reduce_partials_kernel(const float* __restrict__ part, int splits, int64_t stride,
                                       int64_t n, float w, float* __restrict__ grad, int overwrite) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[(int64_t)z * stride + i];
  grad[i] = overwrite ? w * s : grad[i] + w * s;
}

}  // namespace simt
}  // namespace stx
