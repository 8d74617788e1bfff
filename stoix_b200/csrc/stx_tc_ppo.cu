// K3 (bf16 path): the PPO minibatch gradient step on tcgen05 tensor cores.
//
// Reference semantics: _actor_loss_fn / _critic_loss_fn + jax.grad for one minibatch
// (stoix/systems/ppo/anakin/ff_ppo.py:184-247; losses stoix/utils/loss.py:17-32,68-78).  Three launches:
//
//  tc_ppo_fwd_bwd_kernel (K3a)  one persistent CTA per SM; CTAs [0, n_a) run the actor, the rest the critic.
//      Per 128-row tile (rows gathered through the shuffle permutation straight into 128B-swizzled smem):
//        G0 X*W0 -> E0 h1=relu(.+b0) -> G1 h1*W1 -> E1 h2 -> G2 h2*W2 -> E2 loss + d(logits|value) ->
//        G3 dz*W2^T -> E3 dh2 = .*(h2>0) -> G4 dh2*W1^T -> E4 dh1 = .*(h1>0)
//      All five GEMMs run on tcgen05 with the hidden activations resident in tensor memory (A operand
//      from TMEM); W1's single smem image serves as MN-major B (forward) and K-major B (backward).
//      Every 256-wide GEMM is issued as four N=64 column parts with their own mbarriers, so the tensor pipe
//      trails the epilogue part by part (TMEM is full: two tiles cannot be in flight).
//      Epilogue warps also emit h1, h2, dh2, dh1, dz (bf16) for the weight-gradient GEMMs -- deferred, one
//      512-byte store at a time between the column groups of the next chunk, because the SM's store path
//      (32 B/clk) is this kernel's roofline --, reduce the layer-1 / head bias gradients with a shuffle
//      butterfly and accumulate the loss metrics.
//  tc_dw_kernel (K3b)  split-K GEMMs dW = A^T * B over the minibatch rows with both operands MN-major,
//      3-stage TMA/mbarrier pipeline, 256 x N fp32 accumulators in TMEM:
//        dW1 = h1^T dh2 (N=256), dW2 = h2^T dz (N=16), dW0^T = dh1^T x (N=64) + db0 = dh1^T 1 (ones-operand MMA).
//      The activations travel K3a -> K3b in a tiled layout [tile of 128 rows][8-column group][row][8]:
//      a warp of K3a (one row per thread) stores 512 contiguous bytes per instruction, and one TMA box of
//      K3b lands 64 rows of every column group as un-swizzled UMMA core matrices (8 rows x 16 B).
//  tc_reduce_kernel     fixed-order reduction of the per-CTA partials into the flat fp32 gradient arena
//      (deterministic; also the bias gradients and the six loss metrics); optionally fused with clip + Adam
//      (stx_ppo_minibatch_update).
// The three kernels (and K4 behind them) are chained with programmatic dependent launch; activations carry
// evict_first, partials / gradients evict_last L2 policies (see stx_common.cuh).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "stx_common.cuh"
#include "stx_tc_ptx.cuh"

namespace stx {
namespace tc {

int make_map_2d_pub(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows,
                    uint32_t box_cols);  // stx_tc_mlp.cu (SWIZZLE_128B bf16)
int make_map_tiled(CUtensorMap* m, const void* base, uint64_t tiles, uint32_t colgroups);  // stx_tc_mlp.cu

// optional timeline instrumentation (scripts/profile_k3_timeline.py): clock64 stamps of CTA 0, tile 1
__device__ long long* g_clock_buf = nullptr;
#define STX_STAMP_AT(slot, cond)                                   \
  do {                                                             \
    if (clk_buf != nullptr && (cond)) clk_buf[slot] = clock64();   \
  } while (0)
#define STX_STAMP(slot) STX_STAMP_AT(slot, it == 1 && lane == 0)
// optional per-CTA cycle accounting of the fused launch (scripts/profile_k3_fused.py): [gridDim][8] long long
//   K3a CTAs: 0 role id (1), 1 total, 2 cycles inside flag_signal (epilogue warp 5), 3 tiles
//   dW CTAs:  0 role id (2 + job), 1 total, 2 TMA thread waiting for flags, 3 TMA thread waiting for a free stage,
//             4 MMA thread waiting for a full stage, 5 pipeline iterations
__device__ long long* g_prof_buf = nullptr;

constexpr int kTileM = 128;
constexpr int kH = 256;
// warps 0..3 gather producers (one row per thread), warp 4 MMA, warps 5.. epilogue (8 or 16 of them)
constexpr int fb_threads(int nepi) { return 32 * (5 + nepi); }
constexpr int kFbMmaWarp = 4;

// ---- K3a shared-memory map ------------------------------------------------------------------------
constexpr uint32_t kOffW1 = 0;
constexpr uint32_t kOffW0 = 131072;
constexpr uint32_t kOffW2 = 163840;
constexpr uint32_t kOffX = 172032;                 // 2 stages x 16 KB
constexpr uint32_t kOffDz = kOffX + 2 * 16384;     // 4 KB: dz tile, K-major core matrices
constexpr uint32_t kOffBias = kOffDz + 4096;       // b0[256] b1[256] b2[16]
constexpr uint32_t kOffDb = kOffBias + 528 * 4;    // [4 lane quarters][db0[256] db1[256] db2[16]] accumulators
constexpr uint32_t kOffBar = kOffDb + 4 * 528 * 4;
constexpr uint32_t kFbSmemBytes = kOffBar + 128 + 1024;

struct FbNet {
  const __nv_bfloat16* w2;   // [256 x A] bf16
  const float *b0, *b1, *b2;
  __nv_bfloat16 *h1, *h2, *dh1, *dh2;  // [mb x 256], tiled layout (see tiled_ptr)
  __nv_bfloat16* dz;                   // [mb x 16] tiled (columns >= A are zero)
  float* db_part;                      // [n_cta_of_net][528]
  int A;
  int is_actor;
};

struct FbParams {
  FbNet net[2];
  int n_cta[2];             // CTAs per net (actor first)
  const __nv_bfloat16* obs; // [B x D]
  __nv_bfloat16* xg;        // [mb x 64] gathered input rows, tiled layout (written by the actor CTAs)
  const int32_t* idx;       // perm + mb_off, or nullptr (then rows are row0 + i)
  int64_t row0;
  const int32_t* action;
  const float *logp_old, *v_old, *adv, *tgt, *adv_stats;
  float* metric_part;       // [gridDim][8]
  int D;
  int mb;
  float clip_eps, ent_coef, vf_coef;
  // fused launch (tc_ppo_fused_kernel): per-tile completion counters polled by the weight-gradient CTAs (nullptr: none)
  uint32_t* flag_t;   // [2][tiles] every activation (h1, h2, dz, dh2, dh1) of (net, tile) stored -> kFlagEpi arrivals (one per epilogue warp)
  uint32_t* flag_x;   // [tiles]    gathered input rows xg of tile stored                         -> kFlagGather arrivals (actor CTAs)
  int act_policy;     // L2 policy of the activation stores: 0 evict_first (read once, after a kernel boundary), 1 normal, 2 evict_last
  int ring_tiles;     // activation workspace = ring of this many tile slots per network (tile t -> slot t % ring_tiles); 0: one slot per tile
};
constexpr uint32_t kFlagEpi = 8, kFlagGather = 4;

// Release "this warp's stores of the tile are issued" to the consumer CTAs: the warp barrier orders the other lanes' stores
// before lane 0's gpu-scope fence (the same barrier + single-thread fence pattern cooperative-groups grid.sync relies on).
__device__ __forceinline__ void flag_signal(uint32_t* f, int lane) {
  __syncwarp();
  if (lane == 0) {
    __threadfence();
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(f), "r"(1u) : "memory");
  }
}
__device__ __forceinline__ void flag_signal_timed(uint32_t* f, int lane, long long& acc) {
  const long long t0 = clock64();
  flag_signal(f, lane);
  acc += clock64() - t0;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* f) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Tiled activation layout shared by K3a (writer) and K3b (reader): element (row, col) of a [mb x 8*CG]
// matrix lives at ((row/128 * CG + col/8) * 128 + row%128) * 8 + col%8.
__device__ __forceinline__ uint4* tiled_ptr(__nv_bfloat16* base, int64_t row, int cg, int CG) {
  return reinterpret_cast<uint4*>(base + (((row >> 7) * CG + cg) * 128 + (row & 127)) * 8);
}

// sum over the 32 lanes of a warp of v[i], i = 0..31: afterwards v[0] of lane l holds the total of
// column l (reduce-scatter butterfly: 31 shuffles instead of 32 x 5).
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const bool hi = (lane & s) != 0;
      const float send = hi ? v[i] : v[i + s];
      const float keep = hi ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// Actor head of one row: softmax statistics with ONE exponential per logit (e_j = exp(z_j - zmax), p_j = e_j / sum(e),
// log p_j = (z_j - zmax) - log(sum(e))), the clipped-surrogate loss and d(loss)/d(logits) incl. the entropy bonus.
// Fast intrinsics: this path carries bf16-level error anyway.  NA = compile-time bound on the action count.
template <int NA>
__device__ __forceinline__ void actor_head(const uint32_t (&r)[16], const float* s_b2, int A, int a, float logp_old, float adv,
                                           float clip_eps, float ent_coef, float inv_m, float (&dz)[16], float& loss,
                                           float& ent_out) {
  float zs[NA], ex[NA], zmax = -INFINITY;
#pragma unroll
  for (int j = 0; j < NA; ++j) {
    zs[j] = __uint_as_float(r[j]) + s_b2[j];
    if (j < A) zmax = fmaxf(zmax, zs[j]);
  }
  float se = 0.f;
#pragma unroll
  for (int j = 0; j < NA; ++j) {
    zs[j] -= zmax;
    ex[j] = j < A ? __expf(zs[j]) : 0.f;
    se += ex[j];
  }
  const float lse = __logf(se), inv_se = 1.0f / se;
  float ent = 0.f, logp_a = 0.f;
#pragma unroll
  for (int j = 0; j < NA; ++j)
    if (j < A) {
      const float lp = zs[j] - lse;
      ent -= ex[j] * inv_se * lp;
      if (j == a) logp_a = lp;
    }
  const float ratio = __expf(logp_a - logp_old);
  const float l1 = ratio * adv;
  const float l2 = fminf(fmaxf(ratio, 1.0f - clip_eps), 1.0f + clip_eps) * adv;
  const bool in_band = (ratio >= 1.0f - clip_eps) && (ratio <= 1.0f + clip_eps);
  const float dlogp = ((l1 < l2) || in_band) ? -adv * ratio * inv_m : 0.f;
  const float ce = ent_coef * inv_m;
#pragma unroll
  for (int j = 0; j < NA; ++j)
    if (j < A) {
      const float lp = zs[j] - lse, pr = ex[j] * inv_se;
      dz[j] = dlogp * ((j == a ? 1.f : 0.f) - pr) + ce * pr * (lp + ent);
    }
  loss = -fminf(l1, l2);
  ent_out = ent;
}

// The K3a CTA: called by tc_ppo_fwd_bwd_kernel (split launches) and by the producer CTAs of tc_ppo_fused_kernel.
template <int NEPI>
__device__ __forceinline__ void fb_role(uint8_t* smem_raw, const CUtensorMap& tmW0a, const CUtensorMap& tmW1a, const CUtensorMap& tmW0c,
                                        const CUtensorMap& tmW1c, const FbParams& p) {
  // align by OFFSET (not through an integer cast) so that the compiler keeps the shared address space: LDS/STS, not generic LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const long long prof_t0 = clock64();
  long long* const clk_buf = blockIdx.x == 0 ? g_clock_buf : nullptr;  // one load; the stamps themselves stay cheap
  STX_STAMP_AT(56, threadIdx.x == 0);
  float* s_b0 = reinterpret_cast<float*>(smem + kOffBias);
  float* s_b1 = s_b0 + 256;
  float* s_b2 = s_b1 + 256;
  float* s_db = reinterpret_cast<float*>(smem + kOffDb);  // [4][528]: one slice per lane quarter -> no atomics, fixed order
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* x_full = bars;       // [2]
  uint64_t* x_empty = bars + 2;  // [2]
  uint64_t* w_full = bars + 4;
  uint64_t* d_ready = bars + 5;      // [4] MMA -> epilogue: column part p of the current GEMM is complete
  uint64_t* chunk_done = bars + 9;   // [4] epilogue -> MMA: part p consumed (and its slice of the next A operand written)
  uint64_t* head_ready = bars + 13;  // G2 complete
  uint64_t* head_done = bars + 14;   // dz tile in smem
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t pol_stream = l2_policy(p.act_policy);  // default evict_first: the activations are written once and read once (by K3b)
  const int which = (int)blockIdx.x < p.n_cta[0] ? 0 : 1;
  const FbNet& net = p.net[which];
  const int cta_in_net = which == 0 ? (int)blockIdx.x : (int)blockIdx.x - p.n_cta[0];
  const int ncta = p.n_cta[which];
  const int num_tiles = p.mb / kTileM;
  const int ring = p.ring_tiles > 0 ? p.ring_tiles : (num_tiles > 0 ? num_tiles : 1);
  const int my_tiles = cta_in_net < num_tiles ? (num_tiles - cta_in_net + ncta - 1) / ncta : 0;
  const CUtensorMap* tmW0 = which == 0 ? &tmW0a : &tmW0c;
  const CUtensorMap* tmW1 = which == 0 ? &tmW1a : &tmW1c;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&x_full[s], 4);  // one arrival per producer warp
      mbar_init(&x_empty[s], 1);
    }
    mbar_init(w_full, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&d_ready[i], 1);
      mbar_init(&chunk_done[i], 8);  // 4 lane quarters x 2 column halves
    }
    mbar_init(head_ready, 1);
    mbar_init(head_done, 4);
    fence_barrier_init();
    tma_prefetch_desc(tmW0);
    tma_prefetch_desc(tmW1);
  }
  griddep_launch();  // K3b cannot co-reside with this CTA anyway; lets its launch be queued early
  if (warp == kFbMmaWarp) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < 4 * 528; i += fb_threads(NEPI)) s_db[i] = 0.f;
  griddep_wait();  // everything above overlapped the previous optimiser kernel; the weights below are its output
  // Start every long-latency fetch of the prologue at once: the 160 KB of W0/W1 (TMA), the first shuffle indices of the
  // gather and loss-input rows, then the W2 / bias staging below; all of them are in flight together.
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(w_full, 32768 + 131072);
    const uint64_t pol_keep = l2_evict_last();  // the bf16 weights (K4's output) stay L2-resident for all 148 CTAs and the next step
    for (int j = 0; j < 4; ++j) tma_load_2d_hint(smem + kOffW0 + j * 8192, tmW0, w_full, j * 64, 0, pol_keep);
    for (int j = 0; j < 4; ++j) tma_load_2d_hint(smem + kOffW1 + j * 32768, tmW1, w_full, j * 64, 0, pol_keep);
  }
  int32_t idx_first = 0;  // producers: row of tile 0 owned by this thread; epilogue (sub 0): its loss-input row of tile 0
  if (p.idx != nullptr && my_tiles > 0) {
    if (warp < 4) idx_first = p.idx[(int64_t)cta_in_net * kTileM + threadIdx.x];
    else if (warp >= 5 && warp < 9) idx_first = p.idx[(int64_t)cta_in_net * kTileM + (warp & 3) * 32 + lane];
  }
  {
    // one thread per W2 row: all of its (<=16) loads are independent and in flight together
    uint8_t* w2s = smem + kOffW2;
    for (int j = threadIdx.x; j < kH; j += fb_threads(NEPI)) {
      uint32_t pk[8];
      if (net.A == 8 && (reinterpret_cast<uintptr_t>(net.w2) & 15) == 0) {
        const uint4 v = *reinterpret_cast<const uint4*>(net.w2 + j * 8);
        pk[0] = v.x, pk[1] = v.y, pk[2] = v.z, pk[3] = v.w, pk[4] = pk[5] = pk[6] = pk[7] = 0u;
      } else {
        unsigned short e[16];
#pragma unroll
        for (int n = 0; n < 16; ++n) e[n] = n < net.A ? reinterpret_cast<const unsigned short*>(net.w2)[j * net.A + n] : (unsigned short)0;
#pragma unroll
        for (int n = 0; n < 8; ++n) pk[n] = (uint32_t)e[2 * n] | ((uint32_t)e[2 * n + 1] << 16);
      }
      uint8_t* dst = w2s + (j >> 3) * 256 + (j & 7) * 16;   // element (j, n) at (j/8)*256 + (n/8)*128 + (j%8)*16 + (n%8)*2
      *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(dst + 128) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    for (int i = threadIdx.x; i < 256; i += fb_threads(NEPI)) s_b0[i] = net.b0[i], s_b1[i] = net.b1[i];
    if (threadIdx.x < 16) s_b2[threadIdx.x] = threadIdx.x < net.A ? net.b2[threadIdx.x] : 0.f;
    fence_async_proxy();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  STX_STAMP_AT(57, threadIdx.x == 0);

  float m_acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // actor_loss, entropy, value_loss, adv, pred value, target

  if (warp < 4) {
    // ===================== producers: X rows gathered one row per thread =====================
    const int dchunks = p.D >> 3;  // 16-byte chunks per observation row
    const int r = threadIdx.x;     // row of the tile owned by this thread
    int32_t gidx = idx_first;      // shuffle index of this thread's row: fetched one tile ahead, kept raw until used
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it & 1;
      const int tile = cta_in_net + it * ncta;
      const int64_t mrow = (int64_t)tile * kTileM + r;
      const int64_t src = p.idx ? (int64_t)gidx : p.row0 + mrow;
      if (p.idx && it + 1 < my_tiles) gidx = p.idx[mrow + (int64_t)ncta * kTileM];
      uint4 v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c)  // 8 independent 16-byte loads in flight per thread
        v[c] = c < dchunks ? *reinterpret_cast<const uint4*>(p.obs + src * p.D + c * 8) : make_uint4(0u, 0u, 0u, 0u);
      if (warp == 0) STX_STAMP(48);
      if (it >= 2) mbar_wait(&x_empty[s], ((it >> 1) & 1) ^ 1, 1);
      if (warp == 0) STX_STAMP(49);
      uint8_t* xs = smem + kOffX + s * 16384 + r * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(xs + ((c ^ (r & 7)) << 4)) = v[c];  // Swizzle<3,4,3>
      if (which == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) st_hint(tiled_ptr(p.xg, (int64_t)(tile % ring) * kTileM + r, c, 8), v[c], pol_stream);
      }
      if (warp == 0) STX_STAMP(50);
      fence_async_proxy();
      if (warp == 0) STX_STAMP(51);
      __syncwarp();
      if (lane == 0) mbar_arrive(&x_full[s]);
      if (which == 0 && p.flag_x != nullptr) flag_signal(p.flag_x + tile, lane);  // xg rows of this warp are on their way
    }
  } else if (warp == kFbMmaWarp) {
    // ===================== MMA issuer =====================
    // Every 256-wide GEMM is issued as four N=64 column parts.  Part p of the NEXT GEMM only needs (a) D part p
    // consumed by the running epilogue and (b) the K-chunks of its A operand written, both announced per part
    // through chunk_done[]; so the tensor pipe trails the epilogue part by part instead of waiting for all of it.
    // Code shape: groups of four K steps are unrolled (the issuing thread must stay ahead of the 33 cycles an N=64 MMA
    // takes), the loops over parts and groups are rolled and the descriptors advance by adding byte offsets: fully
    // unrolled, the ~200 MMA issues of a tile were a third of the kernel's code and the instruction-cache misses at
    // every phase change stalled the SM for ~2.5k cycles per tile; fully rolled, the issue itself became the bottleneck.
    constexpr uint32_t idesc_fwd = idesc_bf16(128, 64, 0, 1);    // B = W (in,out) image as MN-major
    constexpr uint32_t idesc_head = idesc_bf16(128, 16, 0, 1);
    constexpr uint32_t idesc_bwd = idesc_bf16(128, 64, 0, 0);    // B = same images read as K-major
    constexpr uint32_t idesc_bwd_full = idesc_bf16(128, 256, 0, 0);
    const uint32_t tmem_d = tmem, tmem_a1 = tmem + 256, tmem_a2 = tmem + 384;
    auto adv = [](uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); };  // start-address field, no carry (< 256 KB)
    const uint64_t dW0 = smem_desc(sbase + kOffW0, 8192, 1024, SWIZZLE_128B);
    const uint64_t dW1f = smem_desc(sbase + kOffW1, 32768, 1024, SWIZZLE_128B);  // forward: MN-major
    const uint64_t dW1b = smem_desc(sbase + kOffW1, 16, 1024, SWIZZLE_128B);     // backward: K-major
    const uint64_t dW2f = smem_desc(sbase + kOffW2, 256, 128, SWIZZLE_NONE);
    mbar_wait(w_full, 0, 2);
    STX_STAMP_AT(58, lane == 0);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it & 1;
      STX_STAMP(0);
      mbar_wait(&x_full[s], (it >> 1) & 1, 3);
      const uint64_t dX = smem_desc(sbase + kOffX + s * 16384, 16, 1024, SWIZZLE_128B);
#pragma unroll 1
      for (int pt = 0; pt < 4; ++pt) {  // G0: D = X * W0, trailing E4 of the previous tile
        if (it > 0) mbar_wait(&chunk_done[pt], 1, 4);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t b0 = adv(dW0, pt * 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ss(tmem_d + pt * 64, adv(dX, k * 32), adv(b0, k * 2048), idesc_fwd, k > 0);
          mma_commit(&d_ready[pt]);
          if (pt == 3) mma_commit(&x_empty[s]);
        }
        __syncwarp();
      }
      STX_STAMP(1);
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {  // G1: D = h1 * W1, trailing E0
        mbar_wait(&chunk_done[j], 0, 5);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll 1
          for (int pt = 0; pt <= j; ++pt) {
#pragma unroll 1
            for (int g = (pt == j ? 0 : j); g <= j; ++g) {  // K steps 4g .. 4g+3
              const uint64_t b0 = adv(dW1f, pt * 32768 + g * 8192);
              const uint32_t a0 = tmem_a1 + g * 32;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) mma_ts(tmem_d + pt * 64, a0 + kk * 8, adv(b0, kk * 2048), idesc_fwd, (g | kk) != 0);
            }
            if (j == 3) mma_commit(&d_ready[pt]);
          }
        }
        __syncwarp();
      }
      STX_STAMP(3);
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {  // G2: D[:, :16] = h2 * W2, trailing E1 (columns 0..15 are free after its first part)
        mbar_wait(&chunk_done[j], 1, 6);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t b0 = adv(dW2f, j * 2048);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) mma_ts(tmem_d, tmem_a2 + j * 32 + kk * 8, adv(b0, kk * 512), idesc_head, (j | kk) != 0);
          if (j == 3) mma_commit(head_ready);
        }
        __syncwarp();
      }
      STX_STAMP(5);
      mbar_wait(head_done, it & 1, 7);
      tc_fence_after();
      if (elect_one()) {  // G3: D = dz (smem, K-major core matrices) * W2^T (same W2 image, K-major)
        mma_ss(tmem_d, smem_desc(sbase + kOffDz, 128, 256, SWIZZLE_NONE), smem_desc(sbase + kOffW2, 128, 256, SWIZZLE_NONE),
               idesc_bwd_full, 0);
#pragma unroll
        for (int pt = 0; pt < 4; ++pt) mma_commit(&d_ready[pt]);
      }
      __syncwarp();
      STX_STAMP(7);
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {  // G4: D = dh2 * W1^T, trailing E3  (W1 image as K-major SW128: 4 K-blocks of 64)
        mbar_wait(&chunk_done[j], 0, 8);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll 1
          for (int pt = 0; pt <= j; ++pt) {
#pragma unroll 1
            for (int g = (pt == j ? 0 : j); g <= j; ++g) {  // K steps 4g .. 4g+3 = the g-th 64-wide K block of the image
              const uint64_t b0 = adv(dW1b, g * 32768 + pt * 8192);
              const uint32_t a0 = tmem_a2 + g * 32;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) mma_ts(tmem_d + pt * 64, a0 + kk * 8, adv(b0, kk * 32), idesc_bwd, (g | kk) != 0);
            }
            if (j == 3) mma_commit(&d_ready[pt]);
          }
        }
        __syncwarp();
      }
      STX_STAMP(9);
    }
  } else {
    // ===================== epilogue warps 5..: lane quarter q = warp % 4, column position `sub` =====================
    // Time step cc of an epilogue handles the 32-column chunks cc*kSub + sub (kSub = NEPI/4 warps per lane quarter), so
    // the 64-column parts complete in order and are handed to the MMA warp one by one (NEPI=8: one part per step,
    // NEPI=16: two parts per step).
    constexpr int kSub = NEPI / 4, kSteps = 8 / kSub;
    const int q = warp & 3, sub = (warp - 5) >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t tmem_d = tmem + lane_addr, tmem_a1 = tmem + lane_addr + 256, tmem_a2 = tmem + lane_addr + 384;
    const float inv_m = 1.0f / (float)p.mb;
    float adv_mean = 0.f, adv_rstd = 1.f;
    if (p.adv_stats) adv_mean = p.adv_stats[0], adv_rstd = p.adv_stats[1];
    // source row of this thread's row in the NEXT tile: loaded one tile ahead so that the dependent per-row
    // loads below never wait on the index (two chained DRAM latencies would otherwise open every tile)
    // (kept RAW: a conversion right behind the load would stall this in-order warp for the full memory latency)
    int32_t idx_next = idx_first;  // (tile 0: requested in the prologue)
    // Deferred activation stores.  The SM's global-store path moves 32 B/clk (tools/store_bw.cu): the 2 KB a warp
    // produces per chunk occupy it for 64 cycles, and all 8 warps finish a chunk at about the same time.  Issued in one
    // burst, the later stores of a warp block it behind the other warps' bursts; so the four 512-byte stores of a chunk
    // are issued one at a time between the four column groups of the NEXT chunk this warp processes (whatever phase
    // that chunk belongs to), where the store path has long drained.
    long long prof_sig = 0;
    uint32_t pend[16];
    uint4* pend_ptr = nullptr;
#pragma unroll
    for (int j = 0; j < 16; ++j) pend[j] = 0u;
#define STX_FLUSH_PENDING(g)                                                                                           \
  do {                                                                                                                 \
    if (pend_ptr != nullptr)                                                                                           \
      st_hint(pend_ptr + (g) * 128, make_uint4(pend[4 * (g)], pend[4 * (g) + 1], pend[4 * (g) + 2], pend[4 * (g) + 3]), pol_stream); \
  } while (0)
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = cta_in_net + it * ncta;
      const int64_t mrow = (int64_t)tile * kTileM + q * 32 + lane;  // row inside the minibatch
      const int64_t srow = (int64_t)(tile % ring) * kTileM + q * 32 + lane;  // its row in the activation workspace (slot of the tile)
      // per-row loss inputs: issued now, consumed in E2 (their latency hides behind E0/E1)
      int pf_a = 0;
      float pf_0 = 0.f, pf_1 = 0.f;
      if (sub == 0) {
        const int64_t src = p.idx ? (int64_t)idx_next : p.row0 + mrow;
        if (net.is_actor) pf_a = p.action[src], pf_0 = p.logp_old[src], pf_1 = p.adv[src];
        else pf_0 = p.v_old[src], pf_1 = p.tgt[src];
        if (it + 1 < my_tiles && p.idx) idx_next = p.idx[mrow + (int64_t)ncta * kTileM];
      }
      // ---------------- E0 / E1: hidden layers ----------------
      STX_STAMP_AT(32 + 4 * it, warp == 5 && lane == 0 && it < 4);  // 32, 36, 40, 44: E0 start of tile it
#pragma unroll 1
      for (int layer = 0; layer < 2; ++layer) {
        if (warp == 5) STX_STAMP(16 + 2 * layer);
        const float* bias = layer == 0 ? s_b0 : s_b1;
        const uint32_t ta = layer == 0 ? tmem_a1 : tmem_a2;
        __nv_bfloat16* hout = layer == 0 ? net.h1 : net.h2;
#pragma unroll 1
        for (int cc = 0; cc < kSteps; ++cc) {
          const int c = cc * kSub + sub, part = c >> 1;
          mbar_wait(&d_ready[part], layer, 10 + layer);  // G0 / G1 are the 1st / 2nd completion of d_ready per tile
          tc_fence_after();
          if (warp == 5 && cc == 0) STX_STAMP(17 + 2 * layer);
          uint32_t r[32], pk[16];
          tmem_ld32(tmem_d + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int j = 4 * g; j < 4 * g + 4; ++j) {
              pk[j] = bias_relu_pack_bf16(r[2 * j], r[2 * j + 1], *reinterpret_cast<const float2*>(bias + c * 32 + 2 * j));
            }
            STX_FLUSH_PENDING(g);  // one 512-byte store of the previous chunk (32 lanes = 32 consecutive rows)
          }
          tmem_st16(ta + c * 16, pk);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&chunk_done[part]);  // this chunk of D consumed, its K-chunks of the next A operand written
          // the flushes above issued this warp's last stores (dh1) of the PREVIOUS tile: that tile's activations are complete
          if (layer == 0 && cc == 0 && it > 0 && p.flag_t != nullptr) flag_signal_timed(p.flag_t + which * num_tiles + (tile - ncta), lane, prof_sig);
#pragma unroll
          for (int j = 0; j < 16; ++j) pend[j] = pk[j];
          pend_ptr = tiled_ptr(hout, srow, c * 4, 32);
        }
      }
      // ---------------- E2: head + loss + d(head) ----------------
      if (warp == 5) STX_STAMP(20);
      STX_STAMP_AT(32 + 4 * it + 1, warp == 5 && lane == 0 && it < 4);  // E2 start of tile it
      if (sub == 0) {
        mbar_wait(head_ready, it & 1, 12);
        tc_fence_after();
        if (warp == 5) STX_STAMP(21);
        uint32_t r[16];
        tmem_ld16(tmem_d, r);
        tmem_ld_wait();
        float dz[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) dz[j] = 0.f;
        if (net.is_actor) {
          const float adv = (pf_1 - adv_mean) * adv_rstd;
          float loss, ent;
          if (net.A <= 8) actor_head<8>(r, s_b2, net.A, pf_a, pf_0, adv, p.clip_eps, p.ent_coef, inv_m, dz, loss, ent);
          else actor_head<16>(r, s_b2, net.A, pf_a, pf_0, adv, p.clip_eps, p.ent_coef, inv_m, dz, loss, ent);
          m_acc[0] += loss, m_acc[1] += ent, m_acc[3] += adv;
        } else {
          const float v = __uint_as_float(r[0]) + s_b2[0], vo = pf_0, tg = pf_1;
          const float diff = v - vo;
          const float vclip = vo + fminf(fmaxf(diff, -p.clip_eps), p.clip_eps);
          const float e1 = v - tg, e2 = vclip - tg, q1 = e1 * e1, q2 = e2 * e2;
          const float g2 = (fabsf(diff) < p.clip_eps) ? e2 : 0.f;
          const float dv = q1 > q2 ? e1 : (q1 < q2 ? g2 : 0.5f * (e1 + g2));
          dz[0] = p.vf_coef * dv * inv_m;
          m_acc[2] += 0.5f * fmaxf(q1, q2), m_acc[4] += v, m_acc[5] += tg;
        }
        if (warp == 5) STX_STAMP(41);
        // dz -> bf16: smem A operand (core-matrix K-major) + global (padded row of 64)
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(dz[2 * j], dz[2 * j + 1]);
        const int mr = q * 32 + lane;
        uint8_t* dzs = smem + kOffDz + (mr >> 3) * 256 + (mr & 7) * 16;
        *reinterpret_cast<uint4*>(dzs) = make_uint4(pk[0], pk[1], pk[2], pk[3]);        // k = 0..7
        *reinterpret_cast<uint4*>(dzs + 128) = make_uint4(pk[4], pk[5], pk[6], pk[7]);  // k = 8..15
        fence_async_proxy();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(head_done);
        if (warp == 5) STX_STAMP(44);
        st_hint(tiled_ptr(net.dz, srow, 0, 2), make_uint4(pk[0], pk[1], pk[2], pk[3]), pol_stream);
        st_hint(tiled_ptr(net.dz, srow, 1, 2), make_uint4(pk[4], pk[5], pk[6], pk[7]), pol_stream);
        // bias gradient of the head (off the critical path: G3 is already running): column sums over the 32 rows
        // of this warp; fold the two half-warps first (1 shuffle per column), then the 16-lane butterfly
        {
          float t[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) t[j] = dz[j] + __shfl_xor_sync(0xffffffffu, dz[j], 16);
#pragma unroll
          for (int s2 = 8; s2 >= 1; s2 >>= 1) {
#pragma unroll
            for (int i = 0; i < s2; ++i) {
              const bool hi = (lane & s2) != 0;
              const float send = hi ? t[i] : t[i + s2];
              const float keep = hi ? t[i + s2] : t[i];
              t[i] = keep + __shfl_xor_sync(0xffffffffu, send, s2);
            }
          }
          if (lane < 16) s_db[q * 528 + 512 + lane] += t[0];  // lanes 0..15 hold columns 0..15
        }
      }
      // ---------------- E3 / E4: dh2 = D * (h2 > 0) ; dh1 = D * (h1 > 0) ----------------
#pragma unroll 1
      for (int layer = 1; layer >= 0; --layer) {
        if (warp == 5) STX_STAMP(22 + 2 * (1 - layer));
        const uint32_t ta = layer == 1 ? tmem_a2 : tmem_a1;  // packed h of this layer (mask source)
        __nv_bfloat16* dout = layer == 1 ? net.dh2 : net.dh1;
        float* dbacc = s_db + q * 528 + (layer == 1 ? 256 : 0);
#pragma unroll 1
        for (int cc = 0; cc < kSteps; ++cc) {
          const int c = cc * kSub + sub, part = c >> 1;
          mbar_wait(&d_ready[part], 1 - layer, 13 + layer);  // G3 / G4 are the 3rd / 4th completion per tile
          tc_fence_after();
          uint32_t r[32], hm[16], pk[16];
          tmem_ld32(tmem_d + c * 32, r);
          tmem_ld16(ta + c * 16, hm);
          tmem_ld_wait();
          float dv[32];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int j = 4 * g; j < 4 * g + 4; ++j) {
              // h is post-relu (>= 0): positive <=> bf16 bit pattern non-zero (and not -0)
              const bool p0 = (hm[j] & 0x7FFFu) != 0u, p1 = (hm[j] & 0x7FFF0000u) != 0u;
              dv[2 * j] = p0 ? __uint_as_float(r[2 * j]) : 0.f;
              dv[2 * j + 1] = p1 ? __uint_as_float(r[2 * j + 1]) : 0.f;
              pk[j] = pack_bf16(dv[2 * j], dv[2 * j + 1]);
            }
            STX_FLUSH_PENDING(g);
          }
          if (layer == 1) {  // dh2 replaces h2 as the A operand of G4
            tmem_st16(ta + c * 16, pk);
            tmem_st_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&chunk_done[part]);  // E3: -> G4 ; E4: D columns free for the next tile's G0
#pragma unroll
          for (int j = 0; j < 16; ++j) pend[j] = pk[j];
          pend_ptr = tiled_ptr(dout, srow, c * 4, 32);
          if (layer == 1) {  // db1; db0 = column sums of dh1 comes out of K3b's dW0 job for free (ones-operand MMA)
            const float cs = warp_colsum32(dv, lane);
            dbacc[c * 32 + lane] += cs;  // this (quarter, column) is touched by this warp only
          }
        }
        if (warp == 5) STX_STAMP(26 + (1 - layer));
        STX_STAMP_AT(32 + 4 * it + 2 + (1 - layer), warp == 5 && lane == 0 && it < 4);  // E3 / E4 end of tile it
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) STX_FLUSH_PENDING(g);  // the last chunk of the last tile
    if (my_tiles > 0 && p.flag_t != nullptr) flag_signal_timed(p.flag_t + which * num_tiles + cta_in_net + (my_tiles - 1) * ncta, lane, prof_sig);
#undef STX_FLUSH_PENDING
    STX_STAMP_AT(61, warp == 5 && lane == 0);
    if (g_prof_buf != nullptr && warp == 5 && lane == 0) {
      long long* pb = g_prof_buf + (int64_t)blockIdx.x * 8;
      pb[0] = 1, pb[1] = clock64() - prof_t0, pb[2] = prof_sig, pb[3] = my_tiles;
    }
  }
  // ---- teardown: bias-gradient and metric partials of this CTA ----
  tc_fence_before();
  __syncthreads();
  const uint64_t pol_keep = l2_evict_last();  // read by the reduce kernel after K3b has streamed ~140 MB through L2
  for (int i = threadIdx.x; i < 528; i += fb_threads(NEPI))
    st_hint(&net.db_part[(int64_t)cta_in_net * 528 + i], (s_db[i] + s_db[528 + i]) + (s_db[2 * 528 + i] + s_db[3 * 528 + i]), pol_keep);
  {
    __shared__ float s_mw[fb_threads(NEPI) / 32][6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const float w = warp_sum(m_acc[k]);
      if (lane == 0) s_mw[warp][k] = w;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
      float acc = 0.f;
      if (threadIdx.x < 6)
        for (int w = 0; w < fb_threads(NEPI) / 32; ++w) acc += s_mw[w][threadIdx.x];
      st_hint(&p.metric_part[(int64_t)blockIdx.x * 8 + threadIdx.x], acc, pol_keep);
    }
  }
  if (warp == kFbMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
  STX_STAMP_AT(62, threadIdx.x == 0);
}

template <int NEPI>
__global__ void __launch_bounds__(fb_threads(NEPI), 1)
    tc_ppo_fwd_bwd_kernel(const __grid_constant__ CUtensorMap tmW0a, const __grid_constant__ CUtensorMap tmW1a,
                          const __grid_constant__ CUtensorMap tmW0c, const __grid_constant__ CUtensorMap tmW1c,
                          const FbParams p) {
  extern __shared__ uint8_t smem_raw[];
  fb_role<NEPI>(smem_raw, tmW0a, tmW1a, tmW0c, tmW1c, p);
}

// =============================== K3b: dW = A^T * B ===============================================
constexpr int kDwThreads = 192;  // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int kDwStages = 3;
constexpr uint32_t kDwStageBytes = 65536;  // A [64 x 256] 32 KB + B [64 x <=256] 32 KB
constexpr uint32_t kDwOffOnes = kDwStages * kDwStageBytes;  // 2 KB: B tile [2 column groups][64 rows][16 B] of bf16 1.0
constexpr uint32_t kDwOffBar = kDwOffOnes + 2048;
constexpr uint32_t kDwSmemBytes = kDwOffBar + 128 + 1024;
constexpr int kMaxJobs = 6;
constexpr int kMaxSubs = 2;

// One weight-gradient GEMM of a job: dW = A^T * B over the rows this CTA is given.
struct DwSub {
  float* part;     // [n_cta][256 x N] fp32 partial outputs
  float* colsum;   // nullable: [n_cta][256] per-CTA sums over the rows of every A column (A^T * 1): the bias gradient that
                   // belongs to A = dh1, obtained from one extra N=16 MMA per K step against an all-ones B tile
  int N;           // 16, 64 or 256 (= 8 * column groups of B)
  int map;         // index of the (A, B) tensor-map pair in DwMaps
  int tmem_col;    // accumulator columns [tmem_col, tmem_col + N (+16 with colsum)) of each 256-column half
  int wait_x;      // fused launch: 1 = the B operand is xg (written by the ACTOR CTAs: its own completion counter)
};
// A job = the CTAs [cta_begin, cta_begin + n_cta) running the same 1..2 GEMMs over an interleaved share of the 64-row chunks.
struct DwJob {
  DwSub sub[kMaxSubs];
  int n_sub;
  int cta_begin;   // first CTA of this job (relative to the first weight-gradient CTA of the launch)
  int n_cta;
  int num_chunks;  // mb / 64
  int net;
};
struct DwParams {
  DwJob job[kMaxJobs];
  int n_jobs;
  int num_tiles;
  // fused launch: completion counters written by the K3a CTAs of the same grid (nullptr: operands are complete at entry)
  const uint32_t* flag_t;  // [2][tiles]
  const uint32_t* flag_x;  // [tiles]
  int ring_tiles;  // see FbParams
  int act_policy;
};
struct DwMaps {
  CUtensorMap a[kMaxJobs];  // tiled [mb x 256]: 2D u64 view {256 per (tile, colgroup), tiles*32}, box {128, 32}
  CUtensorMap b[kMaxJobs];  // tiled [mb x N]:   box {128, N/8}
};

// The K3b CTA (warps 0..5 of the block work; further warps of a fused launch only join the block barriers).
// `cta` counts from the first weight-gradient CTA of the launch.
__device__ __forceinline__ void dw_role(uint8_t* smem_raw, const DwMaps& maps, const DwParams& p, const int cta_abs, const bool fused) {
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDwOffBar);
  uint64_t* full = bars;                 // [kDwStages]
  uint64_t* empty = bars + kDwStages;    // [kDwStages]
  uint64_t* acc_done = bars + 2 * kDwStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kDwStages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long prof_t0 = clock64();
  long long* const prof = g_prof_buf ? g_prof_buf + (int64_t)blockIdx.x * 8 : nullptr;

  int j = 0;
  while (j + 1 < p.n_jobs && cta_abs >= p.job[j + 1].cta_begin) ++j;
  const DwJob& job = p.job[j];
  const int cta = cta_abs - job.cta_begin;
  const int my_chunks = cta < job.num_chunks ? (job.num_chunks - cta + job.n_cta - 1) / job.n_cta : 0;
  const int n_sub = job.n_sub;
  const int my_iters = my_chunks * n_sub;  // pipeline slots: (chunk, sub-GEMM) pairs in order
  bool any_colsum = false;
  for (int sj = 0; sj < n_sub; ++sj) any_colsum |= job.sub[sj].colsum != nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kDwStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_done, 1);
    fence_barrier_init();
    for (int sj = 0; sj < n_sub; ++sj) {
      tma_prefetch_desc(&maps.a[job.sub[sj].map]);
      tma_prefetch_desc(&maps.b[job.sub[sj].map]);
    }
  }
  griddep_launch();
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if (any_colsum) {  // all-ones B operand (every layout of a constant tile is the same tile)
    for (int i = threadIdx.x; i < 512; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + kDwOffOnes)[i] = 0x3F803F80u;
    fence_async_proxy();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  griddep_wait();  // the activations are K3a's output; the partials written below were read by the previous reduce

  if (warp == 0) {
    if (elect_one()) {
      const uint64_t pol_stream = l2_policy(p.act_policy);  // default evict_first: every activation tile is read exactly once
      long long w_flag = 0, w_empty = 0;
      int k = 0;
      bool need_x = false;
      for (int sj = 0; sj < n_sub; ++sj) need_x |= job.sub[sj].wait_x != 0;
      // Completion of this CTA's chunks is polled a WINDOW at a time: one batch of independent acquire loads (one L2 round
      // trip) and ONE generic->async proxy fence cover up to kWin chunks, instead of a dependent round trip + fence in front
      // of every TMA issue (measured: ~900 cycles per pipeline slot, which made the loads, not the MMAs, the pace).
      constexpr int kWin = 8;
      int known = fused ? 0 : my_chunks;  // chunks [0, known) of this CTA's sequence have their operands in memory
      for (int it = 0; it < my_chunks; ++it) {
        // split launches: newest rows first (K3a wrote the last tiles most recently, so they are the likeliest L2 residents);
        // fused launch: in the order the K3a CTAs of this grid finish their tiles
        const int ci = cta + it * job.n_cta;
        const int chunk = fused ? ci : job.num_chunks - 1 - ci;  // 64 rows: half of a 128-row tile
        const int tile = chunk >> 1, r0 = (chunk & 1) * 64;
        if (it >= known) {
          const long long tq = clock64();
          uint32_t spins = 0;
          unsigned long long t0 = 0ull;
          for (;;) {
            uint32_t vt[kWin], vx[kWin];
#pragma unroll
            for (int w = 0; w < kWin; ++w) {
              const int tw = (cta + (it + w) * job.n_cta) >> 1;
              const bool in = it + w < my_chunks;
              vt[w] = in ? ld_acquire_gpu(p.flag_t + job.net * p.num_tiles + tw) : 0u;
              vx[w] = (in && need_x) ? ld_acquire_gpu(p.flag_x + tw) : kFlagGather;
            }
            int ok = 0;
#pragma unroll
            for (int w = 0; w < kWin; ++w)
              if (ok == w && it + w < my_chunks && vt[w] >= kFlagEpi && vx[w] >= kFlagGather) ok = w + 1;
            if (ok > 0) {
              known = it + ok;
              break;
            }
            spin_guard(spins, t0, "K3 activations of a tile");
          }
          fence_proxy_async_global();  // the TMA (async proxy) reads below must observe what the acquire loads observed
          w_flag += clock64() - tq;
        }
        for (int sj = 0; sj < n_sub; ++sj, ++k) {
          const DwSub& sub = job.sub[sj];
          const int s = k % kDwStages;
          const long long tq = clock64();
          if (k >= kDwStages) mbar_wait(&empty[s], ((k / kDwStages) & 1) ^ 1, 20);
          w_empty += clock64() - tq;
          uint8_t* st = smem + s * kDwStageBytes;
          mbar_arrive_expect_tx(&full[s], 32768 + (uint32_t)sub.N * 128);
          // one box = rows r0..r0+63 of every column group: smem image [colgroup][row][16 B]
          const int slot = p.ring_tiles > 0 ? tile % p.ring_tiles : tile;
          tma_load_2d_hint(st, &maps.a[sub.map], &full[s], r0 * 2, slot * 32, pol_stream);
          tma_load_2d_hint(st + 32768, &maps.b[sub.map], &full[s], r0 * 2, slot * (sub.N >> 3), pol_stream);
        }
      }
      if (prof) prof[0] = 2 + j, prof[2] = w_flag, prof[3] = w_empty, prof[5] = my_iters;
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_ones = idesc_bf16(128, 16, 1, 1);
    long long w_full = 0;
    int k = 0;
    for (int it = 0; it < my_chunks; ++it) {
      for (int sj = 0; sj < n_sub; ++sj, ++k) {
        const DwSub& sub = job.sub[sj];
        const uint32_t idesc = idesc_bf16(128, sub.N, 1, 1);  // both operands MN-major
        const bool colsum = sub.colsum != nullptr;
        const int s = k % kDwStages;
        const long long tq = clock64();
        mbar_wait(&full[s], (k / kDwStages) & 1, 21);
        w_full += clock64() - tq;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a0 = sbase + s * kDwStageBytes, b0 = a0 + 32768;
          const uint32_t acc = (it > 0) ? 1u : 0u;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {      // 16 rows (K) per step = two core-matrix row groups of 128 B
#pragma unroll
            for (int h = 0; h < 2; ++h)      // M halves: hidden units [128 h, 128 h + 128) = column groups 16 h ..
              mma_ss(tmem + h * 256 + sub.tmem_col, smem_desc(a0 + h * 16384 + ks * 256, 128, 1024, SWIZZLE_NONE),
                     smem_desc(b0 + ks * 256, 128, 1024, SWIZZLE_NONE), idesc, (acc | (uint32_t)(ks > 0)));
            if (colsum) {
#pragma unroll
              for (int h = 0; h < 2; ++h)  // columns [N, N+16) behind the GEMM's: A^T * ones (every column holds the same sums)
                mma_ss(tmem + h * 256 + sub.tmem_col + sub.N, smem_desc(a0 + h * 16384 + ks * 256, 128, 1024, SWIZZLE_NONE),
                       smem_desc(sbase + kDwOffOnes, 128, 1024, SWIZZLE_NONE), idesc_ones, (acc | (uint32_t)(ks > 0)));
            }
          }
          mma_commit(&empty[s]);
        }
        __syncwarp();
      }
    }
    if (elect_one()) mma_commit(acc_done);
    __syncwarp();
    if (prof && lane == 0) prof[4] = w_full;
  } else if (warp < 6) {
    const int q = warp & 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint64_t pol_keep = l2_evict_last();  // the reduce kernel reads these next: keep them in L2 under the activation stream
    if (my_iters > 0) {
      mbar_wait(acc_done, 0, 22);
      tc_fence_after();
    }
    for (int sj = 0; sj < n_sub; ++sj) {
      const DwSub& sub = job.sub[sj];
      // partial layout [N/4][256 rows][4]: the 32 lanes of a warp (= 32 consecutive rows) write 512 contiguous bytes
      float4* out = reinterpret_cast<float4*>(sub.part + (int64_t)cta * 256 * sub.N);
      for (int h = 0; h < 2; ++h) {
        const int m = h * 128 + q * 32 + lane;
        for (int c = 0; c < sub.N / 16; ++c) {
          uint32_t r[16];
          if (my_iters > 0) {
            tmem_ld16(tmem + lane_addr + h * 256 + sub.tmem_col + c * 16, r);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = 0u;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
            st_hint(&out[(int64_t)(c * 4 + i) * 256 + m],
                    make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])),
                    pol_keep);
        }
        if (sub.colsum != nullptr) {
          uint32_t r[16];
          r[0] = 0u;
          if (my_iters > 0) {
            tmem_ld16(tmem + lane_addr + h * 256 + sub.tmem_col + sub.N, r);
            tmem_ld_wait();
          }
          st_hint(&sub.colsum[(int64_t)cta * 256 + m], __uint_as_float(r[0]), pol_keep);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (prof && threadIdx.x == 0) prof[1] = clock64() - prof_t0;
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

__global__ void __launch_bounds__(kDwThreads, 1) tc_dw_kernel(const __grid_constant__ DwMaps maps, const DwParams p) {
  extern __shared__ uint8_t smem_raw[];
  dw_role(smem_raw, maps, p, (int)blockIdx.x, false);
}

// ONE launch for K3a + K3b: the first n_cta[0] + n_cta[1] CTAs are K3a CTAs (fb_role), the rest weight-gradient CTAs
// (dw_role) that consume the activation tiles as the K3a CTAs of the same grid complete them (flag_t / flag_x),
// through L2 instead of after a kernel boundary.  All CTAs are co-resident (grid <= 148, one CTA per SM); a consumer only
// ever waits for K3a CTAs, which wait for nobody outside their own CTA, so the launch cannot deadlock.  The assignment of
// tiles to CTAs is static => the partial sums keep a fixed association order (run-to-run deterministic gradients).
template <int NEPI>
__global__ void __launch_bounds__(fb_threads(NEPI), 1)
    tc_ppo_fused_kernel(const __grid_constant__ CUtensorMap tmW0a, const __grid_constant__ CUtensorMap tmW1a,
                        const __grid_constant__ CUtensorMap tmW0c, const __grid_constant__ CUtensorMap tmW1c,
                        const FbParams p, const __grid_constant__ DwMaps maps, const DwParams dp) {
  extern __shared__ uint8_t smem_raw[];
  const int n_fb = p.n_cta[0] + p.n_cta[1];
  if ((int)blockIdx.x < n_fb) fb_role<NEPI>(smem_raw, tmW0a, tmW1a, tmW0c, tmW1c, p);
  else dw_role(smem_raw, maps, dp, (int)blockIdx.x - n_fb, true);
}

// =============================== reduce partials -> gradient arena ================================
struct RedSeg {
  const float* part;    // partial base (source view: rows x cols with leading dimension src_ld)
  int64_t part_stride;  // floats between the partials of consecutive CTAs
  int n_part;
  int rows, cols;       // elements to take from the source view
  int src_ld;
  int transpose;        // 0: dst[r*dst_ld + c]   1: dst[c*dst_ld + r]
  int dst_ld;
  int64_t dst_off;
  int net;              // optimiser segment (0 actor, 1 critic) for the sum-of-squares side output
  int cg4;              // 1: source is a K3b partial in [cols/4][256 rows][4] order (items enumerate it linearly)
  int items;            // rows*cols/4 (vec) or rows*cols
};
struct RedParams {
  RedSeg seg[12];
  int n_seg;
  int total_items;
  const float* metric_part;  // [n_cta_total][8]
  int n_cta_total;
  float* metrics;            // [6] accumulated
  float weight;
  float inv_mb;
  int overwrite;
  double* sumsq;             // nullable: [2][gridDim] block partials of sum((weight*g)^2) per optimiser segment
  unsigned long long* sumsq_count;  // where the number of block partials per segment (= gridDim) is published
  uint32_t* flags;           // nullable: the fused launch's tile counters, cleared here for the next optimiser step
  int n_flags;
};

constexpr int kRedThreads = 288;  // x 592 blocks >= the ~168k gradient entries of the benchmark networks: one element per thread
constexpr int kRedMaxBlocks = 4 * kNumSMs;  // the optimiser scratch holds 8 x 148 doubles = 2 segments x 592 block partials

// Optimiser context of the fused variant (stx_ppo_minibatch_update): clip_by_global_norm + Adam of stx_adam.cu applied
// by the thread that reduced the gradient entry.  scratch layout = stx_adam.cu's AdamScratch.
struct FusedOpt {
  float* P;
  float* MU;
  float* NU;
  int32_t* counts;
  const StxAdamSeg* segs;
  StxAdamHyper h;
  __nv_bfloat16* P16;
  float* gnorm_out;
  unsigned long long* arrive;   // grid-barrier ticket (monotonic)
  unsigned long long* finish;   // last-block ticket (monotonic)
};

// Fixed-order reduction of the per-CTA partials into the gradient arena.  One launch, every segment
// concurrently; 4 independent accumulators per element for load-level parallelism; the association order
// is fixed => run-to-run deterministic gradients.  Optionally leaves the per-block sum of squares of the
// reduced gradient for the fused optimiser (which can then skip its own norm pass and grid barrier).
//
// FUSED: every thread owns at most ONE gradient entry (host-checked) and keeps it in a register; after the per-block
// sum-of-squares partials are published the blocks meet at a software grid barrier (all blocks are co-resident: the
// host checks the occupancy), every block re-reduces the partials in the fixed order of clip_adam_kernel<PRENORM>
// and the thread applies clip + Adam + the bf16 shadow update to its entry: K4 without its launch, its gradient
// re-read and its own ramp-up.
template <bool FUSED>
__global__ void __launch_bounds__(kRedThreads) tc_reduce_kernel(const RedParams p, float* __restrict__ grad, const FusedOpt f) {
  griddep_launch();
  griddep_wait();
  const uint64_t pol_keep = l2_evict_last();  // partials and gradients: small, reused every step, latency-critical
  if (p.flags != nullptr && blockIdx.x == gridDim.x - 1)  // the fused K3 launch that counted them up has completed
    for (int i = threadIdx.x; i < p.n_flags; i += blockDim.x) p.flags[i] = 0u;
  double sq0 = 0.0, sq1 = 0.0;
  int64_t my_dst = -1;  // FUSED: arena index, value and optimiser segment of this thread's entry
  float my_g = 0.f;
  int my_net = 0;
  for (int gi = blockIdx.x * blockDim.x + threadIdx.x; gi < p.total_items; gi += gridDim.x * blockDim.x) {
    int s = 0, i = gi;
    while (i >= p.seg[s].items) i -= p.seg[s].items, ++s;
    const RedSeg& g = p.seg[s];
    {
      int r, c;
      const float* src;
      if (g.cg4) {  // linear walk over the partial: coalesced reads; (r, c) = (row, column) of the K3b output
        r = (i & 1023) >> 2, c = ((i >> 10) << 2) | (i & 3);
        if (c >= g.cols) continue;
        src = g.part + i;
      } else {
        r = i / g.cols, c = i % g.cols;
        src = g.part + (int64_t)r * g.src_ld + c;
      }
      // latency-bound (the partials were evicted by the activation traffic: every round trip goes to DRAM): 8 independent
      // loads in flight per thread and round trip, 4 accumulation chains in a fixed order.  (16 per round trip with
      // predicated tails measured 3x SLOWER: the batch no longer fits the scoreboard/registers the compiler allots.)
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const int64_t S = g.part_stride;
      int k = 0;
      for (; k + 8 <= g.n_part; k += 8) {
        const float* q = src + (int64_t)k * S;
        const float v0 = ld_hint(q, pol_keep), v1 = ld_hint(q + S, pol_keep), v2 = ld_hint(q + 2 * S, pol_keep), v3 = ld_hint(q + 3 * S, pol_keep);
        const float v4 = ld_hint(q + 4 * S, pol_keep), v5 = ld_hint(q + 5 * S, pol_keep), v6 = ld_hint(q + 6 * S, pol_keep), v7 = ld_hint(q + 7 * S, pol_keep);
        a0 += v0, a1 += v1, a2 += v2, a3 += v3;
        a0 += v4, a1 += v5, a2 += v6, a3 += v7;
      }
      if (k < g.n_part) {  // up to 7 left: one more batch of independent (predicated) loads
        const float* q = src + (int64_t)k * S;
        const int n = g.n_part - k;
        const float v0 = ld_hint(q, pol_keep), v1 = n > 1 ? ld_hint(q + S, pol_keep) : 0.f, v2 = n > 2 ? ld_hint(q + 2 * S, pol_keep) : 0.f;
        const float v3 = n > 3 ? ld_hint(q + 3 * S, pol_keep) : 0.f, v4 = n > 4 ? ld_hint(q + 4 * S, pol_keep) : 0.f, v5 = n > 5 ? ld_hint(q + 5 * S, pol_keep) : 0.f;
        const float v6 = n > 6 ? ld_hint(q + 6 * S, pol_keep) : 0.f;
        a0 += v0, a1 += v1, a2 += v2, a3 += v3;
        a0 += v4, a1 += v5, a2 += v6;
      }
      float* dst = grad + g.dst_off + (g.transpose ? (int64_t)c * g.dst_ld + r : (int64_t)r * g.dst_ld + c);
      float o = p.weight * ((a0 + a1) + (a2 + a3));
      if (!p.overwrite) o += *dst;
      st_hint(dst, o, pol_keep);
      if (g.net) sq1 += (double)o * o;
      else sq0 += (double)o * o;
      if (FUSED) my_dst = dst - grad, my_g = o, my_net = g.net;
    }
  }
  if (p.sumsq != nullptr) {  // only meaningful with overwrite (then o is the whole gradient)
    __shared__ double sred[32];
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      const double b = block_sum<double>(n == 0 ? sq0 : sq1, sred);
      if (threadIdx.x == 0) p.sumsq[(int64_t)n * gridDim.x + blockIdx.x] = b;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *p.sumsq_count = gridDim.x;
  }
  if (blockIdx.x == 0 && threadIdx.x < 6) {
    float acc = 0.f;
    for (int k = 0; k < p.n_cta_total; ++k) acc += p.metric_part[(int64_t)k * 8 + threadIdx.x];
    p.metrics[threadIdx.x] += p.weight * acc * p.inv_mb;
  }
  if (FUSED) {
    __shared__ StxAdamSeg s_seg[2];
    __shared__ float s_gs[2], s_bc1[2], s_bc2[2], s_lr[2], s_gn[2];
    if (threadIdx.x < 2) s_seg[threadIdx.x] = f.segs[threadIdx.x];
    // ---- grid barrier (partials of every block are published above) ----
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned long long t = atomicAdd(f.arrive, 1ull);
      const unsigned long long target = (t / gridDim.x + 1ull) * gridDim.x;
      unsigned int spins = 0;
      unsigned long long t0 = 0ull;
      while (*reinterpret_cast<volatile unsigned long long*>(f.arrive) < target) spin_guard(spins, t0, "reduce+optimiser grid barrier");
      __threadfence();
    }
    __syncthreads();
    // ---- norms: warp s <-> segment s, same summation order as clip_adam_kernel<PRENORM> ----
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < 2) {
      const int s = warp;
      const double* vp = p.sumsq + (int64_t)s * gridDim.x;
      double part = 0.0;
      for (unsigned int b = lane; b < gridDim.x; b += 32) part += __ldcg(vp + b);
      part = warp_sum(part);
      if (lane == 0) {
        const double ss = part * (double)f.h.grad_scale * (double)f.h.grad_scale;
        const float g_norm = (float)sqrt(ss);
        const float clip = (g_norm < s_seg[s].max_grad_norm) ? 1.0f : s_seg[s].max_grad_norm / g_norm;
        const int32_t c = f.counts[2 * s] + 1;
        float lr = s_seg[s].init_lr;
        if (f.h.decay) {
          const int32_t k = f.counts[2 * s + 1] / f.h.steps_per_update;  // floor division, utils/training.py:25
          lr = s_seg[s].init_lr * (1.0f - (float)k / (float)f.h.num_updates);
        }
        s_gs[s] = f.h.grad_scale * clip;
        s_bc1[s] = 1.0f - powf(f.h.b1, (float)c);
        s_bc2[s] = 1.0f - powf(f.h.b2, (float)c);
        s_lr[s] = lr;
        s_gn[s] = g_norm;
      }
    }
    __syncthreads();
    if (my_dst >= 0) {
      const int s = my_net;
      const float ge = my_g * s_gs[s];
      const float me = f.h.b1 * f.MU[my_dst] + (1.0f - f.h.b1) * ge;
      const float ve = f.h.b2 * f.NU[my_dst] + (1.0f - f.h.b2) * ge * ge;
      const float u = (me / s_bc1[s]) / (sqrtf(ve / s_bc2[s]) + f.h.eps);
      const float pn = f.P[my_dst] - s_lr[s] * u;
      f.MU[my_dst] = me, f.NU[my_dst] = ve, f.P[my_dst] = pn;
      if (f.P16) f.P16[my_dst] = __float2bfloat16_rn(pn);
    }
    // the step counters are advanced by whichever block finishes last (every block has read them by then)
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned long long t = atomicAdd(f.finish, 1ull);
      if (t % gridDim.x == gridDim.x - 1) {
        for (int s = 0; s < 2; ++s) {
          f.counts[2 * s] += 1;
          f.counts[2 * s + 1] += 1;
          if (f.gnorm_out) f.gnorm_out[s] = s_gn[s];
        }
      }
    }
  }
}

// =============================== host side =======================================================
inline size_t al(size_t x) { return (x + 255) / 256 * 256; }

struct TcWs {
  __nv_bfloat16 *h1[2], *h2[2], *dh1[2], *dh2[2], *dz[2], *xg;
  float *part_w1[2], *part_w2[2], *part_w0[2], *db_part[2], *db0_part[2], *metric_part;
  uint32_t* flags;  // [2][tiles] tile counters, [tiles] xg counters  (fused launch)
  size_t bytes;
};

constexpr int kCtaPerNet = kNumSMs / 2;              // 74
// Launch plan, per network (x 2 networks = 148 CTAs).
//  split (default): K3a on 74 CTAs, then K3b with 35/18/21 split-K CTAs for dW1 / dW2 / dW0 -- proportional to the
//    bytes each job streams (dW1: h1 + dh2, dW2: h2 + dz, dW0: dh1 + x): K3b alone is HBM-bound, so the CTAs should finish together.
//  fused (STX_K3_FUSED=1, experimental): ONE launch; 64 K3a CTAs (256 tiles of the benchmark minibatch = exactly 4 each),
//    6 CTAs accumulate dW1 and 4 CTAs dW2 + dW0 + db0 behind per-tile flags.  Parity-green, but MEASURED SLOWER on B200
//    (profiles/r02_k3_fused_accounting.txt: 267 us per minibatch step at 64/6/4, 131 us at 40/21/13, split 113 us in the
//    same eager loop): with the full-size activation workspace the stream still goes through HBM (145 MB written + 145 MB
//    read per step), and a weight-gradient CTA with 192 KB of TMA loads in flight sustains only ~44 B/clk against that
//    saturated memory system (1.45k cycles per 64 KB stage against 1.0k of MMA; the 33-40 KB stages of the dW2/dW0 job are
//    pure latency: 1.6k cycles each).  It needs an L2-resident ring of tile slots to pay off; kept for that experiment.
//  Tuning switches: STX_DW_SPLIT="w1,w2,w0" (split), STX_K3_SPLIT="fb,w1,w02" (fused), each summing to 74.
struct K3Plan {
  int w1 = 35, w2 = 18, w0 = 21;   // split
  bool fused = false;
  int fb = 64, fw1 = 6, fw02 = 4;  // fused
  int act_policy = 0;              // L2 policy of the activation stream (STX_K3_POLICY: 0 evict_first, 1 normal, 2 evict_last)
  int ring = 0;                    // fused: tile slots per network of the activation ring (STX_K3_RING; 0 = one slot per tile)
  K3Plan() {
    int a, b, c;
    const char* e = getenv("STX_DW_SPLIT");
    if (e && sscanf(e, "%d,%d,%d", &a, &b, &c) == 3 && a > 0 && b > 0 && c > 0 && a + b + c == kCtaPerNet) w1 = a, w2 = b, w0 = c;
    e = getenv("STX_K3_SPLIT");
    if (e && sscanf(e, "%d,%d,%d", &a, &b, &c) == 3 && a > 0 && b > 0 && c > 0 && a + b + c == kCtaPerNet) fb = a, fw1 = b, fw02 = c;
    e = getenv("STX_K3_FUSED");
    if (e) fused = e[0] == '1';
    e = getenv("STX_K3_RING");
    if (e) ring = atoi(e);
    e = getenv("STX_K3_POLICY");
    if (e) act_policy = atoi(e);
  }
};
static const K3Plan g_plan;
#define kDwCtaW1 (g_plan.w1)
#define kDwCtaW2 (g_plan.w2)
#define kDwCtaW0 (g_plan.w0)

TcWs carve_tc(int64_t mb, char* base) {
  TcWs w{};
  size_t o = 0;
  auto take = [&](size_t bytes) {
    char* q = base ? base + o : nullptr;
    o += al(bytes);
    return q;
  };
  const int tiles = (int)(mb / 128);
  w.flags = (uint32_t*)take((size_t)3 * tiles * 4);  // first: the caller zero-fills the workspace once, the reduce kernel re-zeroes
  const int n_w1 = g_plan.fused ? g_plan.fw1 : g_plan.w1, n_w2 = g_plan.fused ? g_plan.fw02 : g_plan.w2,
            n_w0 = g_plan.fused ? g_plan.fw02 : g_plan.w0;
  for (int n = 0; n < 2; ++n) {
    w.h1[n] = (__nv_bfloat16*)take((size_t)mb * 256 * 2);
    w.h2[n] = (__nv_bfloat16*)take((size_t)mb * 256 * 2);
    w.dh1[n] = (__nv_bfloat16*)take((size_t)mb * 256 * 2);
    w.dh2[n] = (__nv_bfloat16*)take((size_t)mb * 256 * 2);
    w.dz[n] = (__nv_bfloat16*)take((size_t)mb * 16 * 2);
    w.part_w1[n] = (float*)take((size_t)n_w1 * 65536 * 4);
    w.part_w2[n] = (float*)take((size_t)n_w2 * 256 * 16 * 4);
    w.part_w0[n] = (float*)take((size_t)n_w0 * 256 * 64 * 4);
    w.db_part[n] = (float*)take((size_t)kCtaPerNet * 528 * 4);
    w.db0_part[n] = (float*)take((size_t)n_w0 * 256 * 4);
  }
  w.xg = (__nv_bfloat16*)take((size_t)mb * 64 * 2);
  w.metric_part = (float*)take((size_t)kNumSMs * 8 * 4);
  w.bytes = o;
  return w;
}

bool tc_ppo_shape_ok(const StxMlp* m) {
  return m->n_layers == 3 && m->sizes[1] == kH && m->sizes[2] == kH && m->sizes[0] <= 64 && m->sizes[0] % 8 == 0 &&
         m->sizes[3] >= 1 && m->sizes[3] <= 16 && m->activation == STX_ACT_RELU && !m->use_layer_norm;
}

// max dynamic shared memory opt-in, once per device and kernel
template <typename K>
static int opt_in_smem(K kernel, uint32_t bytes, int slot) {
  static unsigned long long done[4] = {0, 0, 0, 0};  // bit d of done[slot]: set on device d
  int dev = 0;
  STX_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 64 && (done[slot] >> dev) & 1ull) return STX_OK;
  STX_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  if (dev < 64) done[slot] |= 1ull << dev;
  return STX_OK;
}

}  // namespace tc

size_t tc_ppo_workspace_bytes(const StxMlp* a, const StxMlp* c, int64_t mb) {
  (void)a, (void)c;
  return tc::carve_tc(mb, nullptr).bytes;
}

int tc_ppo_minibatch_grads(const StxMlp* actor, const StxMlp* critic, const StxPpoBatch* b, int64_t mb_off, int64_t mb,
                           const StxPpoHyper* h, float grad_weight, float* grad_arena, float* metrics, void* ws_raw, size_t,
                           cudaStream_t st, const StxFusedAdam* opt) {
  using namespace tc;
  STX_REQUIRE(tc_ppo_shape_ok(actor) && tc_ppo_shape_ok(critic), STX_E_SHAPE,
              "STX_PREC_BF16 PPO kernels need a relu MLP without LayerNorm, sizes [D<=64 (mult of 8), 256, 256, head<=16]");
  STX_REQUIRE(mb % 128 == 0, STX_E_SHAPE, "STX_PREC_BF16 PPO kernels need a minibatch that is a multiple of 128 rows (got %lld)", (long long)mb);
  STX_REQUIRE(actor->params_bf16 && critic->params_bf16, STX_E_ARG, "STX_PREC_BF16 needs the bf16 shadow arenas");
  const int D = actor->sizes[0];
  const StxMlp* nets[2] = {actor, critic};
  TcWs ws = carve_tc(mb, reinterpret_cast<char*>(ws_raw));
  int64_t aoff, coff, total;
  stx_ppo_arena_offsets(actor, critic, &aoff, &coff, &total);
  const int64_t noff[2] = {aoff, coff};
  const bool fused = g_plan.fused;
  const int num_tiles = (int)(mb / 128);
  const int n_fb = fused ? g_plan.fb : kCtaPerNet;  // K3a CTAs per network

  // ---- K3a parameters ----
  FbParams fp{};
  CUtensorMap tmW0[2], tmW1[2];
  for (int n = 0; n < 2; ++n) {
    const StxMlp* m = nets[n];
    const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(m->params_bf16);
    const int A = m->sizes[3];
    const int64_t off_w1 = (int64_t)D * kH + kH, off_w2 = off_w1 + (int64_t)kH * kH + kH;
    if (int rc = make_map_2d_pub(&tmW0[n], w, (uint64_t)D, kH, kH, 64, 64)) return rc;
    if (int rc = make_map_2d_pub(&tmW1[n], w + off_w1, kH, kH, kH, 256, 64)) return rc;
    FbNet& fn = fp.net[n];
    fn.w2 = w + off_w2;
    fn.b0 = m->params + (int64_t)D * kH, fn.b1 = m->params + off_w1 + (int64_t)kH * kH, fn.b2 = m->params + off_w2 + (int64_t)kH * A;
    fn.h1 = ws.h1[n], fn.h2 = ws.h2[n], fn.dh1 = ws.dh1[n], fn.dh2 = ws.dh2[n], fn.dz = ws.dz[n];
    fn.db_part = ws.db_part[n], fn.A = A, fn.is_actor = (n == 0);
    fp.n_cta[n] = n_fb;
  }
  fp.obs = reinterpret_cast<const __nv_bfloat16*>(b->obs);
  fp.xg = ws.xg;
  fp.idx = b->perm ? b->perm + mb_off : nullptr;
  fp.row0 = mb_off;
  fp.action = b->action, fp.logp_old = b->log_prob, fp.v_old = b->value, fp.adv = b->advantages, fp.tgt = b->targets;
  fp.adv_stats = h->standardize_advantages ? b->adv_stats : nullptr;
  fp.metric_part = ws.metric_part;
  fp.D = D, fp.mb = (int)mb, fp.clip_eps = h->clip_eps, fp.ent_coef = h->ent_coef, fp.vf_coef = h->vf_coef;
  const int ring_tiles = (fused && g_plan.ring > 0 && g_plan.ring < num_tiles) ? g_plan.ring : 0;
  if (fused) fp.flag_t = ws.flags, fp.flag_x = ws.flags + 2 * num_tiles;
  fp.ring_tiles = ring_tiles, fp.act_policy = g_plan.act_policy;

  // ---- K3b parameters ----
  DwMaps maps;
  DwParams dp{};
  int cta = 0, jn = 0, mi = 0;
  const int chunks = (int)(mb / 64);
  const uint64_t tiles = (uint64_t)(ring_tiles > 0 ? ring_tiles : num_tiles);
  int n_w1 = 0, n_w2 = 0, n_w0 = 0;  // partials per network of each weight gradient (for the reduce below)
  for (int n = 0; n < 2; ++n) {
    // tensor-map pairs of this network: m1 (h1, dh2) -> dW1, m2 (h2, dz) -> dW2, m0 (dh1, xg) -> dW0^T (+ db0 = dh1^T * 1)
    const int m1 = mi++, m2 = mi++, m0 = mi++;
    if (int rc = make_map_tiled(&maps.a[m1], ws.h1[n], tiles, 32)) return rc;
    if (int rc = make_map_tiled(&maps.b[m1], ws.dh2[n], tiles, 32)) return rc;
    if (int rc = make_map_tiled(&maps.a[m2], ws.h2[n], tiles, 32)) return rc;
    if (int rc = make_map_tiled(&maps.b[m2], ws.dz[n], tiles, 2)) return rc;
    if (int rc = make_map_tiled(&maps.a[m0], ws.dh1[n], tiles, 32)) return rc;
    if (int rc = make_map_tiled(&maps.b[m0], ws.xg, tiles, 8)) return rc;
    const DwSub s1{ws.part_w1[n], nullptr, 256, m1, 0, 0};
    const DwSub s2{ws.part_w2[n], nullptr, 16, m2, fused ? 96 : 0, 0};
    const DwSub s0{ws.part_w0[n], ws.db0_part[n], 64, m0, 0, 1};
    auto add_job = [&](int n_cta, int n_sub, DwSub a, DwSub c) {
      DwJob& j = dp.job[jn++];
      j.sub[0] = a, j.sub[1] = c, j.n_sub = n_sub, j.cta_begin = cta, j.n_cta = n_cta, j.num_chunks = chunks, j.net = n;
      cta += n_cta;
    };
    if (fused) {
      add_job(g_plan.fw1, 1, s1, s1);
      add_job(g_plan.fw02, 2, s2, s0);  // per chunk: dW2 (operands complete after E3) first, then dW0 (needs dh1)
      n_w1 = g_plan.fw1, n_w2 = n_w0 = g_plan.fw02;
    } else {
      add_job(kDwCtaW1, 1, s1, s1);
      add_job(kDwCtaW2, 1, s2, s2);
      add_job(kDwCtaW0, 1, s0, s0);
      n_w1 = kDwCtaW1, n_w2 = kDwCtaW2, n_w0 = kDwCtaW0;
    }
  }
  dp.n_jobs = jn, dp.num_tiles = num_tiles, dp.ring_tiles = ring_tiles, dp.act_policy = g_plan.act_policy;
  if (fused) dp.flag_t = fp.flag_t, dp.flag_x = fp.flag_x;

  // ---- launches ----
  // 8 epilogue warps: 16 (two parts per step, 80 registers) measured 5 % slower (profiles/README.md)
  if (fused) {
    if (int rc = opt_in_smem(tc_ppo_fused_kernel<8>, kFbSmemBytes > kDwSmemBytes ? kFbSmemBytes : kDwSmemBytes, 2)) return rc;
    STX_CUDA_OK(launch_pdl(tc_ppo_fused_kernel<8>, dim3(2 * n_fb + cta), dim3(fb_threads(8)),
                           kFbSmemBytes > kDwSmemBytes ? kFbSmemBytes : kDwSmemBytes, st, tmW0[0], tmW1[0], tmW0[1], tmW1[1], fp, maps, dp));
    STX_LAUNCH_OK();
  } else {
    if (int rc = opt_in_smem(tc_ppo_fwd_bwd_kernel<8>, kFbSmemBytes, 0)) return rc;
    if (int rc = opt_in_smem(tc_dw_kernel, kDwSmemBytes, 1)) return rc;
    STX_CUDA_OK(launch_pdl(tc_ppo_fwd_bwd_kernel<8>, dim3(2 * n_fb), dim3(fb_threads(8)), kFbSmemBytes, st, tmW0[0], tmW1[0], tmW0[1],
                           tmW1[1], fp));
    STX_LAUNCH_OK();
    STX_CUDA_OK(launch_pdl(tc_dw_kernel, dim3(cta), dim3(kDwThreads), kDwSmemBytes, st, maps, dp));
    STX_LAUNCH_OK();
  }

  // ---- reduce ----
  RedParams rp{};
  int sidx = 0;
  auto add_seg = [&](const float* part, int64_t stride, int n_part, int rows, int cols, int src_ld, int transpose, int dst_ld,
                     int64_t dst_off, int net_id, int cg4) {
    RedSeg g{};
    g.part = part, g.part_stride = stride, g.n_part = n_part, g.rows = rows, g.cols = cols, g.src_ld = src_ld;
    g.transpose = transpose, g.dst_ld = dst_ld, g.dst_off = dst_off, g.net = net_id;
    g.cg4 = cg4;
    g.items = cg4 ? ((cols + 3) / 4) * 1024 : rows * cols;  // scalar items (4x the threads of a float4 version: measurably faster)
    rp.total_items += g.items;
    rp.seg[sidx++] = g;
  };
  for (int n = 0; n < 2; ++n) {
    const int A = nets[n]->sizes[3];
    const int64_t o_w0 = noff[n], o_b0 = o_w0 + (int64_t)D * kH, o_w1 = o_b0 + kH, o_b1 = o_w1 + (int64_t)kH * kH, o_w2 = o_b1 + kH,
                  o_b2 = o_w2 + (int64_t)kH * A;
    add_seg(ws.part_w1[n], 65536, n_w1, kH, kH, 256, 0, kH, o_w1, n, 1);           // dW1[in][out]
    add_seg(ws.part_w0[n], 256 * 64, n_w0, kH, D, 64, 1, kH, o_w0, n, 1);          // part(j, d) -> W0[d][j]
    add_seg(ws.part_w2[n], 256 * 16, n_w2, kH, A, 16, 0, A, o_w2, n, 1);           // dW2[j][a]
    add_seg(ws.db0_part[n], 256, n_w0, 1, kH, 256, 0, kH, o_b0, n, 0);             // db0 from K3b (column sums of dh1)
    add_seg(ws.db_part[n] + 256, 528, n_fb, 1, kH, 528, 0, kH, o_b1, n, 0);
    add_seg(ws.db_part[n] + 512, 528, n_fb, 1, A, 528, 0, A, o_b2, n, 0);
  }
  rp.n_seg = sidx;
  rp.metric_part = ws.metric_part, rp.n_cta_total = 2 * n_fb, rp.metrics = metrics;
  rp.weight = grad_weight, rp.inv_mb = 1.0f / (float)mb;
  rp.overwrite = opt ? 1 : h->overwrite_grads;
  if (fused) rp.flags = ws.flags, rp.n_flags = 3 * num_tiles;
  // side output for the fused optimiser: partials[seg][block] right after the 16-byte header of its scratch
  void* adam_scratch = opt ? opt->scratch : h->adam_scratch;
  rp.sumsq = (rp.overwrite && adam_scratch) ? reinterpret_cast<double*>(reinterpret_cast<char*>(adam_scratch) + 16) : nullptr;
  rp.sumsq_count = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(adam_scratch) + 8);
  int red_blocks = (rp.total_items + kRedThreads - 1) / kRedThreads;  // one item per thread, blocks spread over all SMs
  if (red_blocks > kRedMaxBlocks) red_blocks = kRedMaxBlocks;
  FusedOpt fo{};
  if (opt) {
    // one entry per thread and all blocks co-resident (grid barrier), else the caller must use the two-call path
    static int blocks_per_sm = -1;
    if (blocks_per_sm < 0) STX_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, tc_reduce_kernel<true>, kRedThreads, 0));
    STX_REQUIRE(rp.total_items <= red_blocks * kRedThreads && red_blocks <= blocks_per_sm * kNumSMs, STX_E_SHAPE,
                "stx_ppo_minibatch_update: %d gradient entries do not fit one co-resident wave (%d blocks x %d threads, %d blocks/SM)",
                rp.total_items, red_blocks, kRedThreads, blocks_per_sm);
    fo.P = opt->param_arena, fo.MU = opt->mu, fo.NU = opt->nu, fo.counts = opt->counts, fo.segs = opt->segs, fo.h = opt->hyper;
    fo.P16 = reinterpret_cast<__nv_bfloat16*>(opt->params_bf16), fo.gnorm_out = opt->gnorm_out;
    char* sc = reinterpret_cast<char*>(opt->scratch);
    // own ticket words (stx_adam.cu keeps its barrier / last-block tickets for a 148-block grid in the same scratch: the
    // modulo arithmetic of a ticket is only valid for one grid size)
    char* tail = sc + 16 + sizeof(double) * 8 * kNumSMs;  // = sizeof(AdamScratch) + partials; 64 spare bytes follow
    fo.arrive = reinterpret_cast<unsigned long long*>(tail + 24);
    fo.finish = reinterpret_cast<unsigned long long*>(tail + 32);
    STX_CUDA_OK(launch_pdl(tc_reduce_kernel<true>, dim3(red_blocks), dim3(kRedThreads), 0, st, rp, grad_arena, fo));
  } else {
    STX_CUDA_OK(launch_pdl(tc_reduce_kernel<false>, dim3(red_blocks), dim3(kRedThreads), 0, st, rp, grad_arena, fo));
  }
  STX_LAUNCH_OK();
  return STX_OK;
}

}  // namespace stx

extern "C" int stx_tc_debug_set_prof_buffer(long long* buf) {
  cudaError_t e = cudaMemcpyToSymbol(stx::tc::g_prof_buf, &buf, sizeof(buf));
  return e == cudaSuccess ? 0 : (int)e;
}

extern "C" int stx_tc_debug_set_clock_buffer(long long* buf) {
  cudaError_t e = cudaMemcpyToSymbol(stx::tc::g_clock_buf, &buf, sizeof(buf));
  return e == cudaSuccess ? 0 : (int)e;
}
