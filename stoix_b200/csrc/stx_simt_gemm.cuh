// fp32 CUDA-core GEMM family for the parity path (STX_PREC_F32).
//
// The reference is fp32 everywhere (SURVEY.md "facts"), so this path keeps every product and
// accumulation in fp32 and is what the tight-tolerance parity tests run.  The bf16 tcgen05 path
// (stx_tc_*.cu) is the throughput path.  Three flavours of one 64x64x16 register-tiled kernel:
//   FWD : U[m,n]  = sum_k A'[row(m),k] * W[k,n] (+ b[n])                    nn.Dense, torso.py:26
//   DX  : dU_prev[m,k] = (sum_n dY[m,n] * W[k,n]) * f'(U_prev[m,k])         backward through the activation (no-LayerNorm torso)
//   DW  : dWp[z][k,n] = sum_{m in slice z} A'[row(m),k] * dY[m,n],  dbp[z][n] = sum dY[m,n]
// Hidden layers are stored PRE-activation (U = the Dense output): the operand A' of the next GEMM is rebuilt on load,
//   A' = f(U)                               MLPTorso(activation=f)                       torso.py:31-32, networks/utils.py:9-24
//   A' = f(LN(U) * scale + bias)            MLPTorso(use_layer_norm=True): Dense without bias, nn.LayerNorm (eps 1e-6) torso.py:26-30
// which keeps one code path for every activation (f' is evaluated from U; for relu f(U) > 0 <=> U > 0, so the relu
// numbers are those of the former post-activation storage) and lets LayerNorm use per-row statistics computed once.
// All shapes are bounds-checked (D=4, A=2 of CartPole work).  Split-M partials of DW are reduced in
// a fixed order by reduce_partials_kernel -> run-to-run deterministic gradients.
#pragma once
#include "stx_common.cuh"

namespace stx {
namespace simt {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
constexpr int kThreads = (BM / TM) * (BN / TN);  // 256

enum Mode { FWD = 0, DX = 1, DW = 2 };

// Activation table of stoix/networks/utils.py:9-24 (flax.linen names): value and derivative from the pre-activation.
__device__ __forceinline__ float act_fwd(int kind, float z) {
  switch (kind) {
    case STX_ACT_RELU: return fmaxf(z, 0.f);
    case STX_ACT_TANH: return tanhf(z);
    case STX_ACT_SILU: return z / (1.f + expf(-z));
    case STX_ACT_ELU: return z > 0.f ? z : expm1f(z);
    case STX_ACT_GELU: {  // nn.gelu default approximate=True
      const float t = tanhf(0.7978845608028654f * (z + 0.044715f * z * z * z));
      return 0.5f * z * (1.f + t);
    }
    case STX_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
    case STX_ACT_SOFTPLUS: return fmaxf(z, 0.f) + log1pf(expf(-fabsf(z)));
    default: return z;  // identity / none
  }
}
__device__ __forceinline__ float act_grad(int kind, float z) {
  switch (kind) {
    case STX_ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case STX_ACT_TANH: {
      const float t = tanhf(z);
      return 1.f - t * t;
    }
    case STX_ACT_SILU: {
      const float s = 1.f / (1.f + expf(-z));
      return s * (1.f + z * (1.f - s));
    }
    case STX_ACT_ELU: return z > 0.f ? 1.f : expf(z);
    case STX_ACT_GELU: {
      const float c = 0.7978845608028654f, inner = c * (z + 0.044715f * z * z * z), t = tanhf(inner);
      return 0.5f * (1.f + t) + 0.5f * z * (1.f - t * t) * c * (1.f + 3.f * 0.044715f * z * z);
    }
    case STX_ACT_SIGMOID: {
      const float s = 1.f / (1.f + expf(-z));
      return s * (1.f - s);
    }
    case STX_ACT_SOFTPLUS: return 1.f / (1.f + expf(-z));
    default: return 1.f;
  }
}

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
  // operand transform on load (FWD / DW): A' = f(A) or f(LN(A) * gamma + beta); a_act < 0: A is used as it is (network input)
  int a_act;
  const float* a_stats;  // [rows][2] (mean, rstd) of A's rows, LayerNorm torsos only
  const float* a_gamma;  // [K] / [features]
  const float* a_beta;
  int mask_act;          // DX epilogue: activation whose derivative (from `mask` = U) multiplies the result; < 0: none
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
          if (g.a_act >= 0) {
            if (g.a_stats) v = (v - g.a_stats[2 * row]) * g.a_stats[2 * row + 1] * g.a_gamma[k] + g.a_beta[k];
            v = act_fwd(g.a_act, v);
          }
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
          if (g.a_act >= 0) {
            if (g.a_stats) v = (v - g.a_stats[2 * row]) * g.a_stats[2 * row + 1] * g.a_gamma[f] + g.a_beta[f];
            v = act_fwd(g.a_act, v);
          }
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
        g.C[m * g.N + n] = acc[i][j] + (g.bias ? g.bias[n] : 0.f);  // pre-activation (see the header)
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
        const float d = (g.mask && g.mask_act >= 0) ? act_grad(g.mask_act, g.mask[m * g.N + n]) : 1.f;
        g.C[m * g.N + n] = acc[i][j] * d;
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

// ---- LayerNorm (flax nn.LayerNorm defaults: epsilon 1e-6, scale + bias, over the feature axis) ---------------------
// per-row statistics of U (M x N): stats[m] = (mean, 1/sqrt(var + eps)); one warp per row
__global__ void ln_stats_kernel(const float* __restrict__ U, int64_t M, int N, float* __restrict__ stats) {
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* u = U + row * N;
  float s = 0.f;
  for (int k = lane; k < N; k += 32) s += u[k];
  const float mean = warp_sum(s) / (float)N;
  float q = 0.f;
  for (int k = lane; k < N; k += 32) {
    const float d = u[k] - mean;
    q = fmaf(d, d, q);
  }
  const float var = warp_sum(q) / (float)N;
  if (lane == 0) stats[2 * row] = mean, stats[2 * row + 1] = rsqrtf(var + 1e-6f);
}

// Backward through h = f(z), z = uhat * gamma + beta, uhat = (u - mean) * rstd for the rows of one layer:
//   dz = dh * f'(z);  dgamma += sum_rows dz * uhat;  dbeta += sum_rows dz;  dhat = dz * gamma;
//   du = rstd * (dhat - mean_k(dhat) - uhat * mean_k(dhat * uhat))
// One warp per row; warp w of a block walks rows r0 + w, r0 + w + kLnWarps, ... in order and keeps its column sums of
// (dz * uhat | dz) in registers (lane l owns columns l, l + 32, ...); the warps' sums are merged through shared memory in
// warp order -> part[block][2N], reduced over blocks in a fixed order by reduce_partials_kernel: deterministic.
// D (M x N) holds dh on entry and du on exit.  N <= 32 * kLnMaxCols.
constexpr int kLnWarps = 8, kLnMaxCols = 32;
__global__ void __launch_bounds__(32 * kLnWarps) ln_backward_kernel(float* __restrict__ D, const float* __restrict__ U,
                                                                  const float* __restrict__ stats, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, int act, int64_t M, int N,
                                                                  int rows_per_block, float* __restrict__ part) {
  extern __shared__ float swarp[];  // [kLnWarps][2N]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ag[kLnMaxCols], ab[kLnMaxCols];
#pragma unroll
  for (int j = 0; j < kLnMaxCols; ++j) ag[j] = 0.f, ab[j] = 0.f;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  for (int64_t row = r0 + warp; row < r0 + rows_per_block && row < M; row += kLnWarps) {
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    float* d = D + row * N;
    const float* u = U + row * N;
    float s1 = 0.f, s2 = 0.f;
    for (int k = lane; k < N; k += 32) {
      const float uh = (u[k] - mean) * rstd;
      const float dhat = d[k] * act_grad(act, uh * gamma[k] + beta[k]) * gamma[k];
      s1 += dhat;
      s2 = fmaf(dhat, uh, s2);
    }
    s1 = warp_sum(s1) / (float)N;
    s2 = warp_sum(s2) / (float)N;
#pragma unroll
    for (int j = 0; j < kLnMaxCols; ++j) {
      const int k = lane + 32 * j;
      if (k < N) {
        const float uh = (u[k] - mean) * rstd;
        const float dz = d[k] * act_grad(act, uh * gamma[k] + beta[k]);
        d[k] = rstd * (dz * gamma[k] - s1 - uh * s2);
        ag[j] = fmaf(dz, uh, ag[j]);
        ab[j] += dz;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kLnMaxCols; ++j) {
    const int k = lane + 32 * j;
    if (k < N) swarp[(warp * 2 + 0) * N + k] = ag[j], swarp[(warp * 2 + 1) * N + k] = ab[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) {
    const int which = i / N, k = i % N;
    float acc = 0.f;
    for (int w = 0; w < kLnWarps; ++w) acc += swarp[(w * 2 + which) * N + k];
    part[(int64_t)blockIdx.x * 2 * N + i] = acc;
  }
}

}  // namespace simt
}  // namespace stx
