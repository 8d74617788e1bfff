// K1 (bf16 path): fused 3-layer MLP forward on tcgen05 tensor cores.
//
// Reference semantics: FeedForwardActor / FeedForwardCritic.apply with an MLPTorso[256,256] and a Dense
// head (stoix/networks/base.py:18-59, torso.py:24-33, heads.py:36,134):  out = (relu(relu(x W0 + b0) W1 +
// b1)) W2 + b2.  The reference runs three XLA GEMMs + fusions per apply; here ONE persistent kernel keeps
// all weights of the network resident in shared memory (W0 32 KB + W1 128 KB + W2 8 KB, bf16) and chains
// the three GEMMs of a 128-row tile through tensor memory:
//
//   TMA(X tile, 128B swizzle) -> smem --tcgen05.mma SS--> D0 (TMEM, fp32) --epilogue: +b0, relu, bf16-->
//   A1 (TMEM, packed bf16) --tcgen05.mma TS (A from TMEM)--> D1 --epilogue--> A2 --TS--> D2 --> +b2 -> HBM
//
// so hidden activations never leave the SM (SURVEY.md 7 "Keeping activations out of HBM").
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM
// allocator, warps 2-17 = epilogue (warp w owns TMEM lanes 32*(w%4)..+31, one row per thread, 2 column chunks).
// Operand layouts: X K-major SW128; W0/W1 row-major (in,out) loaded by TMA as [rows x 64-col] blocks =
// MN-major SW128 B operands (the same image serves as the K-major B operand of the backward pass);
// W2 (256 x A<=16) is staged by hand in the un-swizzled core-matrix layout.
// Shapes: hidden = [256,256] exactly, input dim D <= 64 (multiple of 8; zero-filled by TMA), head <= 16.
#include <cuda.h>

#include "stx_common.cuh"
#include "stx_tc_ptx.cuh"

namespace stx {
namespace tc {

constexpr int kTileM = 128;
constexpr int kH = 256;
constexpr int kXStages = 2;
constexpr int kEpiWarps = 16;                 // 4 per TMEM lane quarter: hides the tcgen05.ld / pack latencies
constexpr int kThreads = 32 * (2 + kEpiWarps);  // warp 0 TMA, warp 1 MMA, warps 2.. epilogue

// shared-memory map (bytes from a 1024-aligned base)
constexpr uint32_t kOffW1 = 0;                         // 4 x [256 rows x 128 B]
constexpr uint32_t kOffW0 = 131072;                    // 4 x [ 64 rows x 128 B]
constexpr uint32_t kOffW2 = 163840;                    // 32 x 2 core matrices of 128 B
constexpr uint32_t kOffX = 172032;                     // kXStages x [128 rows x 128 B]
constexpr uint32_t kOffBias = kOffX + kXStages * 16384;  // b0[256] b1[256] b2[16]
constexpr uint32_t kOffBar = kOffBias + (256 + 256 + 16) * 4;
constexpr uint32_t kSmemBytes = kOffBar + 128 + 1024;  // + alignment slack

struct FwdParams {
  const __nv_bfloat16* w2;  // [256 x A] row-major (bf16 shadow arena)
  const float *b0, *b1, *b2;
  float* out;               // [M x A]
  float* dbg_h1;            // optional [M x 256] post-activation of layer 0 (bring-up / tests)
  float* dbg_h2;            // optional [M x 256]
  int64_t M;
  int A;
  int num_tiles;
};

__global__ void __launch_bounds__(kThreads, 1)
    tc_mlp_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW0,
                      const __grid_constant__ CUtensorMap tmW1, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  // align by OFFSET (not through an integer cast) so that the compiler keeps the shared address space (LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  float* s_b0 = reinterpret_cast<float*>(smem + kOffBias);
  float* s_b1 = s_b0 + 256;
  float* s_b2 = s_b1 + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* x_full = bars;             // [kXStages]
  uint64_t* x_empty = bars + 2;        // [kXStages]
  uint64_t* w_full = bars + 4;
  uint64_t* d_ready = bars + 5;      // [4] MMA -> epilogue: 64-column part p of the current layer's accumulator is complete
  uint64_t* chunk_done = bars + 9;   // [4] epilogue -> MMA: part p consumed and its slice of the next A operand written
  uint64_t* head_ready = bars + 13;  // head GEMM complete
  uint64_t* head_done = bars + 14;   // head read out of tensor memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kXStages; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], 1);
    }
    mbar_init(w_full, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&d_ready[i], 1);
      mbar_init(&chunk_done[i], 8);  // 4 lane quarters x 2 chunks of 32 columns
    }
    mbar_init(head_ready, 1);
    mbar_init(head_done, 4);
    fence_barrier_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW0);
    tma_prefetch_desc(&tmW1);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  // W2 image: element (j, n) at (j/8)*256 + (n/8)*128 + (j%8)*16 + (n%8)*2, zero for n >= A
  {
    // one thread per W2 row: all of its (<=16) loads are independent and in flight together
    uint8_t* w2s = smem + kOffW2;
    for (int j = threadIdx.x; j < kH; j += kThreads) {
      uint32_t pk[8];
      if (p.A == 8 && (reinterpret_cast<uintptr_t>(p.w2) & 15) == 0) {
        const uint4 v = *reinterpret_cast<const uint4*>(p.w2 + j * 8);
        pk[0] = v.x, pk[1] = v.y, pk[2] = v.z, pk[3] = v.w, pk[4] = pk[5] = pk[6] = pk[7] = 0u;
      } else {
        unsigned short e[16];
#pragma unroll
        for (int n = 0; n < 16; ++n) e[n] = n < p.A ? reinterpret_cast<const unsigned short*>(p.w2)[j * p.A + n] : (unsigned short)0;
#pragma unroll
        for (int n = 0; n < 8; ++n) pk[n] = (uint32_t)e[2 * n] | ((uint32_t)e[2 * n + 1] << 16);
      }
      uint8_t* dst = w2s + (j >> 3) * 256 + (j & 7) * 16;   // element (j, n) at (j/8)*256 + (n/8)*128 + (j%8)*16 + (n%8)*2
      *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(dst + 128) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    for (int i = threadIdx.x; i < 256; i += kThreads) s_b0[i] = p.b0[i], s_b1[i] = p.b1[i];
    if (threadIdx.x < 16) s_b2[threadIdx.x] = threadIdx.x < p.A ? p.b2[threadIdx.x] : 0.f;
    fence_async_proxy();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int my_tiles = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, 32768 + 131072);
      for (int j = 0; j < 4; ++j) tma_load_2d(smem + kOffW0 + j * 8192, &tmW0, w_full, j * 64, 0);
      for (int j = 0; j < 4; ++j) tma_load_2d(smem + kOffW1 + j * 32768, &tmW1, w_full, j * 64, 0);
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % kXStages;
        if (it >= kXStages) mbar_wait(&x_empty[s], ((it / kXStages) & 1) ^ 1, 1);
        const int tile = blockIdx.x + it * gridDim.x;
        mbar_arrive_expect_tx(&x_full[s], 16384);
        tma_load_2d(smem + kOffX + s * 16384, &tmX, &x_full[s], 0, tile * kTileM);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Every 256-wide GEMM is issued as four N=64 column parts; part p of the NEXT layer only needs D part p consumed
    // and the K-chunks of its A operand written (chunk_done[]), so the tensor pipe trails the epilogue part by part
    // (same scheme as K3a in stx_tc_ppo.cu).  h1 and h2 live in separate TMEM regions for that reason.
    constexpr uint32_t idesc_n64 = idesc_bf16(128, 64, 0, 1);
    constexpr uint32_t idesc_n16 = idesc_bf16(128, 16, 0, 1);
    const uint32_t tmem_d = tmem, tmem_a1 = tmem + 256, tmem_a2 = tmem + 384;
    // rolled part/group loops with incremental descriptors (see stx_tc_ppo.cu: code size / instruction cache)
    auto adv = [](uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); };
    const uint64_t dW0 = smem_desc(sbase + kOffW0, 8192, 1024, SWIZZLE_128B), dW1 = smem_desc(sbase + kOffW1, 32768, 1024, SWIZZLE_128B);
    const uint64_t dW2 = smem_desc(sbase + kOffW2, 256, 128, SWIZZLE_NONE);
    mbar_wait(w_full, 0, 2);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % kXStages;
      mbar_wait(&x_full[s], (it / kXStages) & 1, 3);
      if (it > 0) mbar_wait(head_done, (it - 1) & 1, 4);  // the previous tile's head has left D columns 0..15
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dX = smem_desc(sbase + kOffX + s * 16384, 16, 1024, SWIZZLE_128B);
#pragma unroll 1
        for (int pt = 0; pt < 4; ++pt) {  // layer 0: D0 = X (K-major SW128) * W0 (MN-major SW128), K = 64
          const uint64_t b0 = adv(dW0, pt * 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ss(tmem_d + pt * 64, adv(dX, k * 32), adv(b0, k * 2048), idesc_n64, k > 0);
          mma_commit(&d_ready[pt]);
        }
        mma_commit(&x_empty[s]);
      }
      __syncwarp();
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {  // layer 1: D1 = A1 (TMEM) * W1, K = 256, trailing E0
        mbar_wait(&chunk_done[j], 0, 34);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll 1
          for (int pt = 0; pt <= j; ++pt) {
#pragma unroll 1
            for (int g = (pt == j ? 0 : j); g <= j; ++g) {  // K steps 4g .. 4g+3
              const uint64_t b0 = adv(dW1, pt * 32768 + g * 8192);
              const uint32_t a0 = tmem_a1 + g * 32;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) mma_ts(tmem_d + pt * 64, a0 + kk * 8, adv(b0, kk * 2048), idesc_n64, (g | kk) != 0);
            }
            if (j == 3) mma_commit(&d_ready[pt]);
          }
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {  // head: D2 = A2 (TMEM) * W2 (un-swizzled core matrices), N = 16, trailing E1
        mbar_wait(&chunk_done[j], 1, 35);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t b0 = adv(dW2, j * 2048);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) mma_ts(tmem_d, tmem_a2 + j * 32 + kk * 8, adv(b0, kk * 512), idesc_n16, (j | kk) != 0);
          if (j == 3) mma_commit(head_ready);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps: lane quarter q = warp % 4, `sub` picks the column chunk of a step =============
    // step cc handles the 32-column chunks cc*kSub + sub (kSub warps per lane quarter): two 64-column parts per step
    constexpr int kSub = kEpiWarps / 4, kSteps = 8 / kSub;
    const int q = warp & 3, sub = (warp - 2) >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t tmem_d = tmem + lane_addr, tmem_a1 = tmem + lane_addr + 256, tmem_a2 = tmem + lane_addr + 384;
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int64_t row = (int64_t)tile * kTileM + q * 32 + lane;
#pragma unroll 1
      for (int layer = 0; layer < 2; ++layer) {
        const float* bias = layer == 0 ? s_b0 : s_b1;
        float* dbg = layer == 0 ? p.dbg_h1 : p.dbg_h2;
        const uint32_t ta = layer == 0 ? tmem_a1 : tmem_a2;
#pragma unroll 1
        for (int cc = 0; cc < kSteps; ++cc) {
          const int c = cc * kSub + sub, part = c >> 1;
          mbar_wait(&d_ready[part], layer, 7 + layer);  // layer 0 / 1 are the 1st / 2nd completion of d_ready per tile
          tc_fence_after();
          uint32_t r[32], pk[16];
          tmem_ld32(tmem_d + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            pk[j] = bias_relu_pack_bf16(r[2 * j], r[2 * j + 1], *reinterpret_cast<const float2*>(bias + c * 32 + 2 * j));
          }
          tmem_st16(ta + c * 16, pk);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&chunk_done[part]);
          if (dbg != nullptr && row < p.M) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&pk[j]);
              dbg[row * kH + c * 32 + 2 * j] = __bfloat162float(h.x);
              dbg[row * kH + c * 32 + 2 * j + 1] = __bfloat162float(h.y);
            }
          }
        }
      }
      // head (one row per thread: the sub == 0 warps)
      if (sub == 0) {
        mbar_wait(head_ready, it & 1, 9);
        tc_fence_after();
        uint32_t r[16];
        tmem_ld16(tmem_d, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(head_done);  // the head is in registers: layer 0 of the next tile may overwrite D
        if (row < p.M) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < p.A) p.out[row * p.A + j] = __uint_as_float(r[j]) + s_b2[j];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 2D bf16 tensor map: `rows` x `cols` with row pitch `pitch_elems`; box = box_rows x box_cols
static int make_map_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows,
                       uint32_t box_cols, CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode();
  STX_REQUIRE(enc != nullptr, STX_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  STX_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (pitch_elems * 2) % 16 == 0, STX_E_ALIGN,
              "TMA source must be 16-byte aligned with a 16-byte multiple row pitch (base %p pitch %llu elems)", base,
              (unsigned long long)pitch_elems);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  STX_REQUIRE(r == CUDA_SUCCESS, STX_E_ARG, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return STX_OK;
}

int make_map_2d_pub(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows,
                    uint32_t box_cols) {
  return make_map_2d(m, base, rows, cols, pitch_elems, box_rows, box_cols, CU_TENSOR_MAP_SWIZZLE_128B);
}

// Tiled activation matrix [tiles*128 rows x 8*colgroups] (layout of stx_tc_ppo.cu::tiled_ptr) seen as a 2D
// array of 8-byte words: inner = the 2 KB (128 rows x 16 B) block of one (tile, column group), outer =
// tile*colgroups + group.  Box = 64 rows (128 words) x all column groups, no swizzle.
int make_map_tiled(CUtensorMap* m, const void* base, uint64_t tiles, uint32_t colgroups) {
  EncodeTiledFn enc = get_encode();
  STX_REQUIRE(enc != nullptr, STX_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  STX_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, STX_E_ALIGN, "TMA source must be 16-byte aligned");
  cuuint64_t dims[2] = {256, tiles * colgroups};
  cuuint64_t strides[1] = {2048};
  cuuint32_t box[2] = {128, colgroups};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  STX_REQUIRE(r == CUDA_SUCCESS, STX_E_ARG, "cuTensorMapEncodeTiled (tiled activations) failed with CUresult %d", (int)r);
  return STX_OK;
}

static int g_forward_ctas = 0;  // stx_tc_set_forward_ctas: grid cap of the persistent forward kernel (0 = all SMs)

static bool tc_shape_ok(const StxMlp* m) {
  return m->n_layers == 3 && m->sizes[1] == kH && m->sizes[2] == kH && m->sizes[0] <= 64 && m->sizes[0] % 8 == 0 &&
         m->sizes[3] >= 1 && m->sizes[3] <= 16 && m->activation == STX_ACT_RELU && !m->use_layer_norm;
}

int tc_forward_impl(const StxMlp* m, const void* x, int64_t ldx, int64_t M, float* out, float* dbg_h1, float* dbg_h2,
                    cudaStream_t st) {
  STX_REQUIRE(tc_shape_ok(m), STX_E_SHAPE,
              "STX_PREC_BF16 MLP kernels need a relu torso without LayerNorm and sizes [D<=64 (mult of 8), 256, 256, head<=16]; got %d layers [%d,%d,%d,%d]",
              m->n_layers, m->sizes[0], m->sizes[1], m->sizes[2], m->sizes[3]);
  STX_REQUIRE(m->params_bf16 != nullptr, STX_E_ARG, "STX_PREC_BF16 needs StxMlp.params_bf16 (bf16 shadow of the arena)");
  const int D = m->sizes[0], A = m->sizes[3];
  const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(m->params_bf16);
  const int64_t off_w1 = (int64_t)D * kH + kH, off_w2 = off_w1 + (int64_t)kH * kH + kH;
  CUtensorMap tmX, tmW0, tmW1;
  if (int rc = make_map_2d(&tmX, x, (uint64_t)M, (uint64_t)D, (uint64_t)ldx, kTileM, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (int rc = make_map_2d(&tmW0, w, (uint64_t)D, kH, kH, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (int rc = make_map_2d(&tmW1, w + off_w1, kH, kH, kH, 256, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  FwdParams p{};
  p.w2 = w + off_w2;
  p.b0 = m->params + (int64_t)D * kH;
  p.b1 = m->params + off_w1 + (int64_t)kH * kH;
  p.b2 = m->params + off_w2 + (int64_t)kH * A;
  p.out = out, p.dbg_h1 = dbg_h1, p.dbg_h2 = dbg_h2, p.M = M, p.A = A;
  p.num_tiles = (int)((M + kTileM - 1) / kTileM);
  static bool attr_set = false;
  if (!attr_set) {
    STX_CUDA_OK(cudaFuncSetAttribute(tc_mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_set = true;
  }
  const int cap = (g_forward_ctas > 0 && g_forward_ctas < kNumSMs) ? g_forward_ctas : kNumSMs;
  const int grid = p.num_tiles < cap ? p.num_tiles : cap;
  tc_mlp_fwd_kernel<<<grid, kThreads, kSmemBytes, st>>>(tmX, tmW0, tmW1, p);
  STX_LAUNCH_OK();
  return STX_OK;
}

}  // namespace tc

size_t tc_mlp_forward_workspace_bytes(const StxMlp*, int64_t) { return 256; }

int tc_mlp_forward(const StxMlp* mlp, const void* x, int64_t ldx, const int32_t* row_idx, int64_t M, float* out, void*, size_t,
                   cudaStream_t st) {
  STX_REQUIRE(row_idx == nullptr, STX_E_UNSUPPORTED, "STX_PREC_BF16 stx_mlp_forward does not take a row gather");
  return tc::tc_forward_impl(mlp, x, ldx, M, out, nullptr, nullptr, st);
}

}  // namespace stx

extern "C" void stx_tc_set_forward_ctas(int n) { stx::tc::g_forward_ctas = n; }

// Bring-up / test hook: forward that also returns the bf16-rounded hidden activations.
extern "C" int stx_tc_debug_forward(const StxMlp* mlp, const void* x, int64_t ldx, int64_t M, float* out, float* h1, float* h2,
                                    void* stream) {
  return stx::tc::tc_forward_impl(mlp, x, ldx, M, out, h1, h2, (cudaStream_t)stream);
}
