// Large-batch residual-TD3 learner on the 5th-generation tensor cores (tcgen05 + TMEM), TF32 operands (round-to-nearest),
// fp32 accumulate.  Opt-in throughput mode (TD3.precision = "tf32") for batches that fill the GPU with 64-row tiles;
// the fp32 FFMA kernels of rtd3_td3.cu stay the parity path and are the reference these are tested against
// (tests/test_tc_learner_gpu.py: losses / Q-values / targets / gradients within the TF32 tolerance).
//
// Reference behaviour: robot.py:312-366 (train_critic), :369-398 (train_actor); networks robot.py:128-206 with ONE
// hidden-to-hidden layer (layers == 2, the benchmark shape 2 x 256 of BASELINE.json configs[2]/[4]).
//
// One CTA = 64 batch rows through the whole chain; everything between the replay gather and the gradient reduction
// stays on chip:
//   * activations X live in shared memory in the UMMA no-swizzle canonical layout [H/4 k-chunks][64 rows][4 floats].
//     The SAME bytes serve as a K-major operand (rows = M, columns = K: forward  H1 = H0 W1^T, backward dH0 = dZ1 W1)
//     and as an MN-major operand (columns = M or N, rows = K: weight gradient dW1 = dZ1^T H0, K = the 64 batch rows);
//   * hidden products are tcgen05.mma.cta_group::1.kind::tf32, M=64 (forward / dX) or M=128 (dW), N=H, K=8 per
//     instruction, issued by one thread; accumulators in TMEM (512 columns: [0,H) forward / dX, [0,2H) dW);
//   * weights stream from L2 as 32-wide K slabs (one contiguous cp.async.bulk each) through a 2-stage mbarrier ring fed
//     by a free-running producer warp; forward reads the chunk-major copy Wu[(k/4)*H+n][k%4] = W[n][k], the input
//     gradient the copy of the transpose Wv[(n/4)*H+k][n%4] = W[n][k] (both kept by the optimiser kernel);
//   * the dW tile of a CTA (H x H fp32) goes TMEM -> registers -> swizzled shared staging -> ONE 2-D TMA reduce-add
//     (cp.reduce.async.bulk.tensor) per 32x32 block into the gradient arena: the batch reduction over CTAs happens in
//     L2, no atomics issued by the SMs, no partial-gradient round trip through HBM;
//   * first / output layers (fan-in <= 4, fan-out <= 2), bias / small-weight gradients: fp32 FFMA + warp shuffles,
//     reduced into the arena with 1-D bulk reduce-adds.
#include <cuda.h>

#include "rtd3_common.cuh"
#include "rtd3_mlp.cuh"
#include "rtd3_tc.cuh"
#include "rtd3_td3.cuh"

namespace rtd3 {

constexpr int kLtRows = 64;           // batch rows per CTA = UMMA M of the forward / dX products
constexpr int kLtEpiWarps = 16;       // 4 per scheduler: the fp32 phases between the products are issue / latency bound
constexpr int kLtEpi = kLtEpiWarps * 32;
constexpr int kLtThreads = kLtEpi + 64;   // warps 0-15: row / epilogue threads, warp 16: TMA producer, warp 17: MMA issuer
constexpr int kLtEpiMma = kLtEpi + 32;    // named barrier 2: epilogue warps + the MMA warp
constexpr int kLtColGroups = kLtEpi / kLtRows;   // thread t: row t % 64, column group t / 64 (first / output layer)

// Shared-memory plan in floats (the base is aligned to 1024 B by hand: the dW staging tiles are SWIZZLE_128B boxes).
template <int H>
struct Lt {
  static constexpr int kBuf = H * kLtRows;          // one activation buffer [H/4][64][4]
  static constexpr int kSlabK = 16;                 // K per weight slab = 2 MMAs
  static constexpr int kStages = 4;                 // slabs in flight: the TMA latency (~800 cycles) against 256 cycles of MMA per slab
  static constexpr int kStage = H * kSlabK;         // one K slab: [kSlabK/4 chunks][H][4]
  static constexpr int kSlabs = H / kSlabK;
  static constexpr int oBufA = 0;
  static constexpr int oBufB = kBuf;
  static constexpr int oStage = 2 * kBuf;           // 2 weight stages, aliased by the dW staging tiles (8 warps x 2 x 4 KB)
  static constexpr int kStageRegion = 16384;        // 64 KB
  static constexpr int oW0 = oStage + kStageRegion; // [H][4] first-layer weights, zero padded
  static constexpr int oB = oW0 + 4 * H;            // [2][H] biases of layer 0 / 1
  static constexpr int oWo = oB + 2 * H;            // [2][H] output weights, zero padded
  static constexpr int oBo = oWo + 2 * H;           // [4] output bias
  static constexpr int oG = oBo + 4;                // small-gradient staging: W0|b0 [5H] | b1 [H] | Wout [2H]
  static constexpr int oIn0 = oG + 8 * H;           // [64][4] network input
  static constexpr int oS = oIn0 + 256;             // [64][8] per-row scalars
  static constexpr int oOut = oS + 512;             // [64][2] network output
  static constexpr int oDout = oOut + 128;          // [64][2] gradient w.r.t. the output
  static constexpr int oPart = oDout + 128;         // [8][64][2] partial sums of the output layer / input gradient
  static constexpr int oBar = oPart + kLtColGroups * 128;          // full[2] empty[2] acc_ready stage_free (uint64) + tmem slot
  static constexpr int kFloats = oBar + 32;         // full[4] empty[4] acc_ready stage_free (uint64) + tmem slot
  static constexpr size_t kBytes = (size_t)kFloats * 4 + 1024;
  static_assert(kStages * kStage <= kStageRegion && 2 * H * 32 <= kStageRegion, "weight stages / re-laid dW operands exceed their region");   // = the two re-laid dW operands of lt_dw
};

// Development aid: phase timestamps of CTA 0 / thread 0 (rtd3_debug_lt_prof), off unless switched on.
__device__ long long g_lt_prof[128];
__device__ int g_lt_prof_on = 0;
#define LT_STAMP(i)                                                                                 \
  do {                                                                                              \
    if (g_lt_prof_on && blockIdx.x == 0 && threadIdx.x == 0) g_lt_prof[(i)] = clock64();             \
  } while (0)

struct LtCtx {
  uint32_t tmem;
  uint32_t it_p = 0, it_c = 0;        // slab counters of the producer / the MMA issuer
  uint32_t acc_phase = 0;             // parity of acc_ready the epilogue warps wait for next
  uint32_t gate_phase = 0;            // parity of stage_free the producer waits for next
};

struct LtMaps {
  alignas(64) CUtensorMap dw[2];      // gradient of the hidden-to-hidden weights of the trained net(s): [H n][H k] fp32
};

__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, int c0, int c1, const void* smem_src) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(c0), "r"(c1), "r"(smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void bulk_reduce_add_f32(float* gdst, const float* smem_src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---- small parameters of one network -> shared memory (epilogue threads) ------------------------------------------------
// Two halves so that the global-load latency hides behind a streamed product: lt_small_fetch issues the loads into
// registers (call it before waiting for an accumulator), lt_small_store writes them to shared memory once the previous
// network's parameters are no longer read.
template <int H>
struct LtSmall {
  static constexpr int kW0 = (4 * H + kLtEpi - 1) / kLtEpi, kB = (H + kLtEpi - 1) / kLtEpi, kWo = (2 * H + kLtEpi - 1) / kLtEpi;
  float w0[kW0], b0[kB], b1[kB], wo[kWo], bo;
};

template <int H>
__device__ __forceinline__ void lt_small_fetch(LtSmall<H>& r, const float* __restrict__ P, const NetShape& s) {
  using S = LtSmall<H>;
  const int t = threadIdx.x;
#pragma unroll
  for (int k = 0; k < S::kW0; ++k) {
    const int i = t + k * kLtEpi, c = i >> 2, j = i & 3;
    r.w0[k] = (i < 4 * H && j < s.in) ? __ldg(P + c * s.in + j) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < S::kB; ++k) {
    const int i = t + k * kLtEpi;
    r.b0[k] = i < H ? __ldg(P + net_b_off(s, 0) + i) : 0.f;
    r.b1[k] = i < H ? __ldg(P + net_b_off(s, 1) + i) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < S::kWo; ++k) {
    const int i = t + k * kLtEpi;
    r.wo[k] = (i < 2 * H && (i / H) < s.out) ? __ldg(P + net_w_off(s, 2) + i) : 0.f;
  }
  r.bo = (t < 4 && t < s.out) ? __ldg(P + net_b_off(s, 2) + t) : 0.f;
}

template <int H>
__device__ __forceinline__ void lt_small_store(float* sm, const LtSmall<H>& r) {
  using L = Lt<H>;
  using S = LtSmall<H>;
  const int t = threadIdx.x;
#pragma unroll
  for (int k = 0; k < S::kW0; ++k)
    if (t + k * kLtEpi < 4 * H) sm[L::oW0 + t + k * kLtEpi] = r.w0[k];
#pragma unroll
  for (int k = 0; k < S::kB; ++k) {
    const int i = t + k * kLtEpi;
    if (i < H) { sm[L::oB + i] = r.b0[k]; sm[L::oB + H + i] = r.b1[k]; }
  }
#pragma unroll
  for (int k = 0; k < S::kWo; ++k)
    if (t + k * kLtEpi < 2 * H) sm[L::oWo + t + k * kLtEpi] = r.wo[k];
  if (t < 4) sm[L::oBo + t] = r.bo;
  bar_sync(1, kLtEpi);
}

template <int H>
__device__ __forceinline__ void lt_load_small(float* sm, const float* __restrict__ P, const NetShape& s) {
  LtSmall<H> r;
  lt_small_fetch<H>(r, P, s);
  lt_small_store<H>(sm, r);
}

// ---- first layer: X[r][c] = tf32(relu(b0[c] + sum_j in0[r][j] W0[c][j])) in the chunk layout ------------------------------
// A warp takes every 8th 4-column chunk (weights broadcast from shared memory, held in registers for two rows per lane).
template <int H>
__device__ __forceinline__ void lt_layer0(float* sm, int buf) {
  using L = Lt<H>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 xa = ld4(sm + L::oIn0 + lane * 4), xb = ld4(sm + L::oIn0 + (lane + 32) * 4);
#pragma unroll 2
  for (int c = warp; c < H / 4; c += kLtEpiWarps) {
    const float4 b = ld4(sm + L::oB + 4 * c);
    const float bv[4] = {b.x, b.y, b.z, b.w};
    float ha[4], hb[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 w = ld4(sm + L::oW0 + (4 * c + q) * 4);
      float va = bv[q], vb = bv[q];
      va = fmaf(xa.x, w.x, va); va = fmaf(xa.y, w.y, va); va = fmaf(xa.z, w.z, va); va = fmaf(xa.w, w.w, va);
      vb = fmaf(xb.x, w.x, vb); vb = fmaf(xb.y, w.y, vb); vb = fmaf(xb.z, w.z, vb); vb = fmaf(xb.w, w.w, vb);
      ha[q] = tf32_rn(fmaxf(va, 0.f));
      hb[q] = tf32_rn(fmaxf(vb, 0.f));
    }
    st4(sm + buf + (c * kLtRows + lane) * 4, make_float4(ha[0], ha[1], ha[2], ha[3]));
    st4(sm + buf + (c * kLtRows + lane + 32) * 4, make_float4(hb[0], hb[1], hb[2], hb[3]));
  }
}

// Sum V values per lane over the 32 lanes of a warp with V-1 + log2(32/V) shuffles (log2(V) halving exchanges, then a butterfly
// over the remaining lane bits) instead of 5V.  Afterwards the lanes with (lane & (32/V - 1)) == 0 hold in v[0] the total of
// value number lane / (32/V).
template <int V>
__device__ __forceinline__ void warp_reduce_many(float (&v)[V]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int w = V / 2, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
    const bool hi = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = hi ? v[i] : v[i + w];
      const float keep = hi ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
#pragma unroll
  for (int bit = (32 / V) >> 1; bit >= 1; bit >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], bit);
}

// ---- streamed product: TMEM[64 x H] = X(a_buf)[64 x H] * Wg^T, Wg chunk-major [H/4][H][4] in global memory --------------
// Called by all warps; the operands written by the epilogue threads before the call are fenced here.
template <int H>
__device__ __forceinline__ void lt_gemm(float* sm, LtCtx& cx, const float* __restrict__ Wg, int a_buf, bool gate) {
  using L = Lt<H>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::oBar);
  uint64_t *full = bars, *empty = bars + L::kStages, *acc = bars + 2 * L::kStages, *sfree = bars + 2 * L::kStages + 1;
  constexpr uint32_t kBytes = (uint32_t)L::kStage * 4;
  if (warp == kLtEpiWarps) {
    if (lane == 0) {
      if (gate) {                                   // the stages are aliased by the dW staging tiles: wait for that drain
        mbar_wait(sfree, cx.gate_phase);
        cx.gate_phase ^= 1;
      }
      for (int ks = 0; ks < L::kSlabs; ++ks, ++cx.it_p) {
        const int st = cx.it_p % L::kStages;
        mbar_wait(empty + st, ((cx.it_p / L::kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(full + st, kBytes);
        bulk_g2s(sm + L::oStage + st * L::kStage, Wg + (size_t)ks * L::kStage, kBytes, full + st);
      }
    }
    __syncwarp();
  } else if (warp == kLtEpiWarps + 1) {
    bar_sync(2, kLtEpiMma);
    if (lane == 0) {
      tc_fence_after();
      // D = F32, A = B = TF32, both K-major, N = H, M = 64
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(kLtRows >> 4) << 24);
      for (int ks = 0; ks < L::kSlabs; ++ks, ++cx.it_c) {
        const int st = cx.it_c % L::kStages;
        mbar_wait(full + st, (cx.it_c / L::kStages) & 1);
        tc_fence_after();
#pragma unroll
        for (int k4 = 0; k4 < L::kSlabK / 8; ++k4) {
          const int kk = ks * (L::kSlabK / 8) + k4;  // K step of 8 = chunks 2kk, 2kk+1
          const uint64_t ad = umma_desc_kmajor(smem_u32(sm + a_buf + kk * (2 * kLtRows * 4)), kLtRows * 16, 128);
          const uint64_t bd = umma_desc_kmajor(smem_u32(sm + L::oStage + st * L::kStage + k4 * (2 * H * 4)), (uint32_t)H * 16, 128);
          umma_tf32(cx.tmem, ad, bd, idesc, kk != 0 ? 1u : 0u);
        }
        umma_commit(empty + st);
      }
      umma_commit(acc);
    }
    __syncwarp();
  } else {
    fence_proxy_async();
    tc_fence_before();
    bar_sync(2, kLtEpiMma);
  }
}

// ---- weight gradient: TMEM[H x H] = dZ(dz_buf)^T [H x 64] * Hprev(h_buf) [64 x H], both operands MN-major ------------------
// TF32 operands can only be MN-major in the SWIZZLE_128B_BASE32B layout (the no-swizzle layout the activations live in
// yields zeros - tools/umma/mn_test.cu), so the two tiles are re-laid, 32 batch rows per round, into the weight-stage
// region (idle here):  [g = column/32][kg = row/4][row%4][32 floats], the 32 B units of a 128 B row XOR-ed with row%4
// (Swizzle<2,5,2>);  SBO = 512 B between 4-row groups, LBO = 4096 B between 32-column groups.  Two rounds accumulate.
template <int H>
__device__ __forceinline__ void lt_dw(float* sm, LtCtx& cx, int dz_buf, int h_buf) {
  using L = Lt<H>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* acc = reinterpret_cast<uint64_t*>(sm + L::oBar) + 2 * L::kStages;
  constexpr int kOp = H * 32;                        // floats of one re-laid operand (32 rows x H)
#pragma unroll 1
  for (int round = 0; round < 2; ++round) {
    if (warp == kLtEpiWarps) {
    } else if (warp == kLtEpiWarps + 1) {
      bar_sync(2, kLtEpiMma);
      if (lane == 0) {
        tc_fence_after();
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint64_t kSw128Base32 = 1ull << 61;
#pragma unroll 1
        for (int hf = 0; hf < H / 128; ++hf) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = umma_desc_kmajor(smem_u32(sm + L::oStage + hf * (4 * 8 * 128) + ks * 256), 4096, 512) | kSw128Base32;
            const uint64_t bd = umma_desc_kmajor(smem_u32(sm + L::oStage + kOp + ks * 256), 4096, 512) | kSw128Base32;
            umma_tf32(cx.tmem + (uint32_t)(hf * H), ad, bd, idesc, (round | ks) != 0 ? 1u : 0u);
          }
        }
        umma_commit(acc);
      }
      __syncwarp();
    } else {
      const int t = threadIdx.x;
      if (round == 0) bar_sync(1, kLtEpi);           // the tiles were just written by other epilogue threads
#pragma unroll 2
      for (int i = t; i < 32 * (H / 4); i += kLtEpi) {
        const int c4 = i >> 5, bl = i & 31;          // 16 B column chunk, row inside the round
        const int row = bl & 3, kg = bl >> 2;
        const int dst = (((c4 >> 3) * 8 + kg) * 4 + row) * 32 + (((((c4 & 7) >> 1) ^ row)) << 3) + ((c4 & 1) << 2);
        const int src = (c4 * kLtRows + 32 * round + bl) * 4;
        st4(sm + L::oStage + dst, ld4(sm + dz_buf + src));
        st4(sm + L::oStage + kOp + dst, ld4(sm + h_buf + src));
      }
      fence_proxy_async();
      tc_fence_before();
      bar_sync(2, kLtEpiMma);
      if (round == 0) {                              // the region is rewritten: wait until round 0's MMAs have read it
        mbar_wait(acc, cx.acc_phase);
        cx.acc_phase ^= 1;
        tc_fence_after();
      }
    }
  }
}

// ---- dW tile: TMEM -> registers -> swizzled staging -> 2-D TMA reduce-add into the gradient arena ------------------------
template <int H>
__device__ __forceinline__ void lt_drain(float* sm, LtCtx& cx, const CUtensorMap* tm) {
  using L = Lt<H>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::oBar);
  uint64_t *acc = bars + 2 * L::kStages, *sfree = bars + 2 * L::kStages + 1;
  if (warp < kLtEpiWarps) {
    const int q = warp & 3, ch = warp >> 2;
    mbar_wait(acc, cx.acc_phase);
    cx.acc_phase ^= 1;
    tc_fence_after();
    float* stg = sm + L::oStage + warp * 1024;       // one 32 x 32 fp32 tile per warp (4 KB, 1024 B aligned)
    // Every CTA reduces into the same H x H block of the arena, and the L2 serialises reduce-adds to one address.  The four warps
    // of a TMEM sub-partition (ch = 0..3) share its kTiles = (row halves) x (32-column blocks) tiles, kIter each; which tile a
    // warp takes at step `it` is rotated by the CTA index, so that at any moment the CTAs are spread over all kTiles positions
    // (with a rotation over the warp's own kIter tiles only, a quarter of the CTAs hit the same addresses at the same time).
    constexpr int kTiles = (H / 128) * (H / 32), kIter = kTiles / 4;
#pragma unroll 1
    for (int it = 0; it < kIter; ++it) {
      {
        const int e = (ch * kIter + it + (int)blockIdx.x) % kTiles;
        const int hf = e / (H / 32), cb = (e % (H / 32)) * 32;
        float v[32];
        tmem_ld32(cx.tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(hf * H + cb), v);
        float* dst = stg;
        if (it >= 1) {
          if (lane == 0) bulk_wait_read<0>();        // the previous reduce has consumed the tile (4 warps per scheduler interleave)
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) st4(dst + lane * 32 + ((j ^ (lane & 7)) << 2), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(tm, cb, hf * 128 + 32 * q, dst);     // rows n (outer), columns k (inner)
          bulk_commit();
        }
      }
    }
    if (lane == 0) {
      bulk_wait_read<0>();
      mbar_arrive(sfree);
    }
    __syncwarp();
    tc_fence_before();
  }
}

// ---- epilogues of a streamed product ---------------------------------------------------------------------------------------
// kMode 0: dst = tf32(relu(acc + bias))                   (forward)
// kMode 1: dst = relu'(dst) * tf32(acc)   in place        (input gradient; dst holds the layer's forward output)
template <int H, int kMode>
__device__ __forceinline__ void lt_epilogue(float* sm, LtCtx& cx, int bias_off, int dst_buf) {
  using L = Lt<H>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* acc = reinterpret_cast<uint64_t*>(sm + L::oBar) + 2 * L::kStages;
  const int q = warp & 3, ch = warp >> 2;
  const int r = 16 * q + (lane & 15);               // M = 64: row i sits in TMEM lane (i % 16) + 32 * (i / 16)
  mbar_wait(acc, cx.acc_phase);
  cx.acc_phase ^= 1;
  tc_fence_after();
#pragma unroll 1
  for (int cb = ch * (H / 4); cb < (ch + 1) * (H / 4); cb += 32) {
    float v[32];
    tmem_ld32(cx.tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)cb, v);
    if (lane < 16) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float* p = sm + dst_buf + (((cb >> 2) + j) * kLtRows + r) * 4;
        float4 o;
        if (kMode == 0) {
          const float4 b = ld4(sm + bias_off + cb + 4 * j);
          o.x = tf32_rn(fmaxf(v[4 * j + 0] + b.x, 0.f)); o.y = tf32_rn(fmaxf(v[4 * j + 1] + b.y, 0.f));
          o.z = tf32_rn(fmaxf(v[4 * j + 2] + b.z, 0.f)); o.w = tf32_rn(fmaxf(v[4 * j + 3] + b.w, 0.f));
        } else {
          const float4 m = ld4(p);
          o.x = m.x > 0.f ? tf32_rn(v[4 * j + 0]) : 0.f; o.y = m.y > 0.f ? tf32_rn(v[4 * j + 1]) : 0.f;
          o.z = m.z > 0.f ? tf32_rn(v[4 * j + 2]) : 0.f; o.w = m.w > 0.f ? tf32_rn(v[4 * j + 3]) : 0.f;
        }
        st4(p, o);
      }
    }
  }
  tc_fence_before();
}

// Input-gradient epilogue of the critic inside the actor step: nothing is stored; the first-layer mask is recomputed from
// the critic input and the gradient w.r.t. the action columns is reduced on the fly:
//   da[r][o] = sum_c relu'(h0[r][c]) * acc[r][c] * W0[c][2+o]            (robot.py:386-390 through critic 1)
template <int H>
__device__ __forceinline__ void lt_epilogue_din(float* sm, LtCtx& cx) {
  using L = Lt<H>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* acc = reinterpret_cast<uint64_t*>(sm + L::oBar) + 2 * L::kStages;
  const int q = warp & 3, ch = warp >> 2;
  const int r = 16 * q + (lane & 15);
  mbar_wait(acc, cx.acc_phase);
  cx.acc_phase ^= 1;
  tc_fence_after();
  const float4 x = ld4(sm + L::oIn0 + r * 4);
  float d0 = 0.f, d1 = 0.f;
#pragma unroll 1
  for (int cb = ch * (H / 4); cb < (ch + 1) * (H / 4); cb += 32) {
    float v[32];
    tmem_ld32(cx.tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)cb, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float4 w = ld4(sm + L::oW0 + (cb + j) * 4);
      float h = sm[L::oB + cb + j];
      h = fmaf(x.x, w.x, h); h = fmaf(x.y, w.y, h); h = fmaf(x.z, w.z, h); h = fmaf(x.w, w.w, h);
      const float g = h > 0.f ? v[j] : 0.f;
      d0 = fmaf(g, w.z, d0);
      d1 = fmaf(g, w.w, d1);
    }
  }
  if (lane < 16) {
    sm[L::oPart + (ch * kLtRows + r) * 2] = d0;
    sm[L::oPart + (ch * kLtRows + r) * 2 + 1] = d1;
  }
  tc_fence_before();
}

// ---- output layer: out[r][o] = bo[o] + sum_c X[r][c] Wo[o][c] ------------------------------------------------------------
template <int H>
__device__ __forceinline__ void lt_out_layer(float* sm, int buf) {
  using L = Lt<H>;
  const int t = threadIdx.x, r = t & 63, cq = t >> 6;
  float p0 = 0.f, p1 = 0.f;
#pragma unroll 4
  for (int c = cq * (H / kLtColGroups); c < (cq + 1) * (H / kLtColGroups); c += 4) {
    const float4 h = ld4(sm + buf + ((c >> 2) * kLtRows + r) * 4);
    const float4 w0 = ld4(sm + L::oWo + c), w1 = ld4(sm + L::oWo + H + c);
    p0 = fmaf(h.x, w0.x, p0); p0 = fmaf(h.y, w0.y, p0); p0 = fmaf(h.z, w0.z, p0); p0 = fmaf(h.w, w0.w, p0);
    p1 = fmaf(h.x, w1.x, p1); p1 = fmaf(h.y, w1.y, p1); p1 = fmaf(h.z, w1.z, p1); p1 = fmaf(h.w, w1.w, p1);
  }
  sm[L::oPart + (cq * kLtRows + r) * 2] = p0;
  sm[L::oPart + (cq * kLtRows + r) * 2 + 1] = p1;
  bar_sync(1, kLtEpi);
  if (t < kLtRows) {
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      float v = sm[L::oBo + o];
#pragma unroll
      for (int g = 0; g < kLtColGroups; ++g) v += sm[L::oPart + (g * kLtRows + t) * 2 + o];
      sm[L::oOut + t * 2 + o] = v;
    }
  }
}

// ---- backward through the output layer, in place on buf (h1 -> dz1), with the sums the small gradients need ----------------
//   gWout[o][c] = sum_r dout[r][o] h1[r][c];   dz1[r][c] = relu'(h1) * sum_o dout[r][o] Wo[o][c];   gb1[c] = sum_r dz1[r][c]
template <int H, int NOUT>
__device__ __forceinline__ void lt_bwd_out(float* sm, int buf) {
  using L = Lt<H>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float2 da = *reinterpret_cast<const float2*>(sm + L::oDout + lane * 2);
  const float2 db = *reinterpret_cast<const float2*>(sm + L::oDout + (lane + 32) * 2);
  constexpr int V = NOUT == 1 ? 8 : 16;              // [gWout row 0 (4) | gb1 (4) | gWout row 1 (4) | pad]
#pragma unroll 2
  for (int c = warp; c < H / 4; c += kLtEpiWarps) {
    float* pa = sm + buf + (c * kLtRows + lane) * 4;
    float* pb = pa + 32 * 4;
    const float4 ha = ld4(pa), hb = ld4(pb);
    const float4 w0 = ld4(sm + L::oWo + 4 * c), w1 = ld4(sm + L::oWo + H + 4 * c);
    const float hav[4] = {ha.x, ha.y, ha.z, ha.w}, hbv[4] = {hb.x, hb.y, hb.z, hb.w};
    const float w0v[4] = {w0.x, w0.y, w0.z, w0.w}, w1v[4] = {w1.x, w1.y, w1.z, w1.w};
    float za[4], zb[4], v[V];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      za[j] = hav[j] > 0.f ? tf32_rn(fmaf(da.x, w0v[j], da.y * w1v[j])) : 0.f;
      zb[j] = hbv[j] > 0.f ? tf32_rn(fmaf(db.x, w0v[j], db.y * w1v[j])) : 0.f;
      v[j] = fmaf(da.x, hav[j], db.x * hbv[j]);
      v[4 + j] = za[j] + zb[j];
      if (NOUT > 1) { v[8 + j] = fmaf(da.y, hav[j], db.y * hbv[j]); v[12 + j] = 0.f; }
    }
    st4(pa, make_float4(za[0], za[1], za[2], za[3]));
    st4(pb, make_float4(zb[0], zb[1], zb[2], zb[3]));
    warp_reduce_many<V>(v);
    if ((lane & (32 / V - 1)) == 0) {
      const int j = lane / (32 / V);                 // value number: 0-3 gWout[0], 4-7 gb1, 8-11 gWout[1]
      if (j < 4) sm[L::oG + 6 * H + 4 * c + j] = v[0];
      else if (j < 8) sm[L::oG + 5 * H + 4 * c + (j - 4)] = v[0];
      else if (j < 12) sm[L::oG + 7 * H + 4 * c + (j - 8)] = v[0];
    }
  }
}

// ---- first-layer gradients from dz0 (buf): gb0[c] = sum_r dz0[r][c];  gW0[c][j] = sum_r dz0[r][c] in0[r][j] -----------------
template <int H, int NIN>
__device__ __forceinline__ void lt_colsum_in(float* sm, int buf) {
  using L = Lt<H>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 xa4 = ld4(sm + L::oIn0 + lane * 4), xb4 = ld4(sm + L::oIn0 + (lane + 32) * 4);
  const float xa[4] = {xa4.x, xa4.y, xa4.z, xa4.w}, xb[4] = {xb4.x, xb4.y, xb4.z, xb4.w};
  constexpr int V = NIN == 4 ? 32 : 16;              // value q * (1 + NIN) + k: k = 0 bias, k >= 1 input k-1, of column 4c + q
#pragma unroll 2
  for (int c = warp; c < H / 4; c += kLtEpiWarps) {
    const float4 za4 = ld4(sm + buf + (c * kLtRows + lane) * 4), zb4 = ld4(sm + buf + (c * kLtRows + lane + 32) * 4);
    const float za[4] = {za4.x, za4.y, za4.z, za4.w}, zb[4] = {zb4.x, zb4.y, zb4.z, zb4.w};
    float v[V];
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v[q * (1 + NIN)] = za[q] + zb[q];
#pragma unroll
      for (int j = 0; j < NIN; ++j) v[q * (1 + NIN) + 1 + j] = fmaf(za[q], xa[j], zb[q] * xb[j]);
    }
    warp_reduce_many<V>(v);
    if ((lane & (32 / V - 1)) == 0) {
      const int i = lane / (32 / V);
      if (i < 4 * (1 + NIN)) {
        const int q = i / (1 + NIN), k = i - q * (1 + NIN), col = 4 * c + q;
        if (k == 0) sm[L::oG + H * NIN + col] = v[0];
        else sm[L::oG + col * NIN + k - 1] = v[0];
      }
    }
  }
}

// ---- small gradients: staging -> gradient arena (1-D bulk reduce-adds; the two output-bias floats by atomicAdd) -------------
template <int H>
__device__ __forceinline__ void lt_reduce_small(float* sm, float* __restrict__ G, const NetShape& s) {
  using L = Lt<H>;
  const int t = threadIdx.x;
  fence_proxy_async();
  bar_sync(1, kLtEpi);
  if (t == 0) {
    bulk_reduce_add_f32(G, sm + L::oG, (uint32_t)(H * (s.in + 1) * 4));
    bulk_reduce_add_f32(G + net_b_off(s, 1), sm + L::oG + 5 * H, (uint32_t)(H * 4));
    bulk_reduce_add_f32(G + net_w_off(s, 2), sm + L::oG + 6 * H, (uint32_t)(s.out * H * 4));
    bulk_commit();
    bulk_wait_read<0>();
  } else if (t >= 32 && t < 32 + s.out) {
    const int o = t - 32;
    float v = 0.f;
    for (int r = 0; r < kLtRows; ++r) v += sm[L::oDout + r * 2 + o];
    atomicAdd(G + net_b_off(s, 2) + o, v);
  }
  bar_sync(1, kLtEpi);
}

// ---- common prologue / epilogue of the two kernels ---------------------------------------------------------------------------
template <int H>
__device__ __forceinline__ float* lt_setup(LtCtx& cx) {
  using L = Lt<H>;
  extern __shared__ unsigned char lt_raw[];
  float* sm = reinterpret_cast<float*>(lt_raw + ((1024u - (smem_u32(lt_raw) & 1023u)) & 1023u));   // stays in the shared window
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::oBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * L::kStages + 2);
  const int t = threadIdx.x;
  if (t == 0) {
    for (int i = 0; i < 2 * L::kStages; ++i) mbar_init(bars + i, 1);   // full[kStages], empty[kStages]
    mbar_init(bars + 2 * L::kStages, 1);                // acc_ready
    mbar_init(bars + 2 * L::kStages + 1, kLtEpiWarps);  // stage_free: one arrival per epilogue warp
    fence_mbar_init();
  }
  if (t < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cx.tmem = *tmem_slot;
  return sm;
}

__device__ __forceinline__ void lt_teardown(const LtCtx& cx) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(cx.tmem), "n"(512) : "memory");
}

// ================================================================================================================================
// critic step (robot.py:312-366 up to the optimiser steps), see td3_critic_kernel in rtd3_td3.cu for the fp32 twin
// ================================================================================================================================
template <int H>
__global__ void __launch_bounds__(kLtThreads, 1)
td3_critic_tc_kernel(Arena ar, const float* __restrict__ params, const float* __restrict__ params_uv, float* __restrict__ grads, ReplayView rp,
                     const int32_t* __restrict__ idx, const float* __restrict__ noise, int B, Td3Hyper hp, float* __restrict__ loss,
                     float* __restrict__ q_out, float* __restrict__ y_out, int32_t* __restrict__ steps, double* __restrict__ beta_pows,
                     const __grid_constant__ LtMaps maps) {
  using L = Lt<H>;
  LtCtx cx;
  float* sm = lt_setup<H>(cx);
  const int t = threadIdx.x, warp = t >> 5;
  const int r0 = blockIdx.x * kLtRows;
  if (blockIdx.x == 0 && t == 0) advance_adam_clock(steps, beta_pows, 1);
  float* S = sm + L::oS;          // [64][8]: 0 s.x 1 s.y 2 a.x 3 a.y 4 reward 5 notdone 6 y 7 valid
  float* in0 = sm + L::oIn0;
  if (t < kLtRows) {
    const int row = r0 + t;
    const bool valid = row < B;
    const int j = valid ? idx[row] : 0;
    const float2 s = rp.s[j], a = rp.a[j], s2 = rp.s2[j];
    S[t * 8 + 0] = s.x; S[t * 8 + 1] = s.y; S[t * 8 + 2] = a.x; S[t * 8 + 3] = a.y;
    S[t * 8 + 4] = rp.r[j]; S[t * 8 + 5] = rp.notdone[j]; S[t * 8 + 7] = valid ? 1.f : 0.f;
    in0[t * 4 + 0] = s2.x; in0[t * 4 + 1] = s2.y; in0[t * 4 + 2] = 0.f; in0[t * 4 + 3] = 0.f;
  }
  __syncthreads();
  LT_STAMP(0);
  const float* Pu = params_uv;                    // forward operand order
  const float* Pv = params_uv + ar.total();       // input-gradient operand order
  LtSmall<H> sp;                                  // small parameters of the next pass, fetched one product ahead
  if (warp < kLtEpiWarps) lt_small_fetch<H>(sp, params + ar.off(3), ar.actor);

  // pass 0: target actor(s2); 1, 2: target critics(s2, a'); 3, 4: critics(s, a) forward + backward
#pragma unroll 1
  for (int pass = 0; pass < 5; ++pass) {
    const int net = pass == 0 ? 3 : (pass == 1 ? 4 : (pass == 2 ? 5 : pass - 2));
    const NetShape shape = pass == 0 ? ar.actor : ar.critic;
    const bool train = pass >= 3;
    const int64_t w1 = ar.off(net) + net_w_off(shape, 1);
    const int hbuf = train ? L::oBufB : L::oBufA;
    if (warp < kLtEpiWarps) {
      lt_small_store<H>(sm, sp);
      LT_STAMP(1 + pass * 16 + 0);
      lt_layer0<H>(sm, L::oBufA);
      LT_STAMP(1 + pass * 16 + 1);
    }
    lt_gemm<H>(sm, cx, Pu + w1, L::oBufA, false);
    if (warp < kLtEpiWarps) {
      if (pass < 4) lt_small_fetch<H>(sp, params + ar.off(pass == 0 ? 4 : (pass == 1 ? 5 : pass - 1)), ar.critic);   // next pass's network
      LT_STAMP(1 + pass * 16 + 2);
      lt_epilogue<H, 0>(sm, cx, L::oB + H, hbuf);
      LT_STAMP(1 + pass * 16 + 3);
      bar_sync(1, kLtEpi);
      lt_out_layer<H>(sm, hbuf);
      if (t < kLtRows) {
        const float* out = sm + L::oOut;
        float* dout = sm + L::oDout;
        if (pass == 0) {                                   // smoothing noise and clips (robot.py:338-339)
          const int row = min(r0 + t, B - 1);
          const float2 zn = target_noise(noise, hp, row);
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            float e = (o == 0 ? zn.x : zn.y) * hp.policy_noise;
            e = fminf(fmaxf(e, -hp.noise_clip), hp.noise_clip);
            in0[t * 4 + 2 + o] = fminf(fmaxf(out[t * 2 + o] + e, -hp.max_action), hp.max_action);
          }
        } else if (pass == 1) {
          S[t * 8 + 6] = out[t * 2];
        } else if (pass == 2) {                            // clipped double-Q target (robot.py:342-345)
          const float qmin = fminf(S[t * 8 + 6], out[t * 2]);
          const float y = S[t * 8 + 4] + hp.gamma * qmin * S[t * 8 + 5];
          S[t * 8 + 6] = y;
          if (y_out && r0 + t < B) y_out[r0 + t] = y;
          in0[t * 4 + 0] = S[t * 8 + 0]; in0[t * 4 + 1] = S[t * 8 + 1];
          in0[t * 4 + 2] = S[t * 8 + 2]; in0[t * 4 + 3] = S[t * 8 + 3];
        } else {                                           // MSE loss and its gradient (robot.py:348-353)
          const int c = pass - 3;
          const float valid = S[t * 8 + 7];
          const float q = out[t * 2];
          const float diff = (q - S[t * 8 + 6]) * valid;
          dout[t * 2] = 2.0f * diff / (float)B;
          dout[t * 2 + 1] = 0.f;
          if (q_out && valid != 0.f) q_out[c * B + r0 + t] = q;
          const float l = warp_sum(diff * diff / (float)B);
          if ((t & 31) == 0) atomicAdd(loss + c, l);
        }
      }
      bar_sync(1, kLtEpi);
      LT_STAMP(1 + pass * 16 + 4);
      if (train) lt_bwd_out<H, 1>(sm, L::oBufB);
      LT_STAMP(1 + pass * 16 + 5);
    }
    if (train) {
      lt_dw<H>(sm, cx, L::oBufB, L::oBufA);
      LT_STAMP(1 + pass * 16 + 6);
      lt_drain<H>(sm, cx, &maps.dw[pass - 3]);
      LT_STAMP(1 + pass * 16 + 7);
      lt_gemm<H>(sm, cx, Pv + w1, L::oBufB, true);
      if (warp < kLtEpiWarps) {
        LT_STAMP(1 + pass * 16 + 8);
        lt_epilogue<H, 1>(sm, cx, 0, L::oBufA);
        LT_STAMP(1 + pass * 16 + 9);
        bar_sync(1, kLtEpi);
        lt_colsum_in<H, 4>(sm, L::oBufA);
        LT_STAMP(1 + pass * 16 + 10);
        lt_reduce_small<H>(sm, grads + ar.off(net), shape);
        LT_STAMP(1 + pass * 16 + 11);
      }
    }
  }
  LT_STAMP(90);
  lt_teardown(cx);
  LT_STAMP(91);
}

// ================================================================================================================================
// actor step (robot.py:369-398 up to the optimiser step): L = -mean(Q1(s, pi(s))), gradient w.r.t. the actor only
// ================================================================================================================================
template <int H>
__global__ void __launch_bounds__(kLtThreads, 1)
td3_actor_tc_kernel(Arena ar, const float* __restrict__ params, const float* __restrict__ params_uv, float* __restrict__ grads, ReplayView rp,
                    const int32_t* __restrict__ idx, int B, float* __restrict__ loss, int32_t* __restrict__ steps, double* __restrict__ beta_pows,
                    const __grid_constant__ LtMaps maps) {
  using L = Lt<H>;
  LtCtx cx;
  float* sm = lt_setup<H>(cx);
  const int t = threadIdx.x, warp = t >> 5;
  const int r0 = blockIdx.x * kLtRows;
  if (blockIdx.x == 0 && t == 0) advance_adam_clock(steps, beta_pows, 0);
  float* S = sm + L::oS;          // [64][8]: 0 s.x 1 s.y 2 a.x 3 a.y 7 valid
  float* in0 = sm + L::oIn0;
  float* dout = sm + L::oDout;
  if (t < kLtRows) {
    const int row = r0 + t;
    const bool valid = row < B;
    const float2 s = rp.s[valid ? idx[row] : 0];
    S[t * 8 + 0] = s.x; S[t * 8 + 1] = s.y; S[t * 8 + 7] = valid ? 1.f : 0.f;
    in0[t * 4 + 0] = s.x; in0[t * 4 + 1] = s.y; in0[t * 4 + 2] = 0.f; in0[t * 4 + 3] = 0.f;
  }
  __syncthreads();
  const float* Pu = params_uv;
  const float* Pv = params_uv + ar.total();
  const int64_t wa = ar.off(0) + net_w_off(ar.actor, 1), wc = ar.off(1) + net_w_off(ar.critic, 1);

  // ---- a = pi(s) (fed the raw replay state, robot.py:386): h0a -> A, h1a -> B (kept for the backward pass) ----
  LtSmall<H> sp;
  if (warp < kLtEpiWarps) {
    lt_load_small<H>(sm, params + ar.off(0), ar.actor);
    lt_layer0<H>(sm, L::oBufA);
  }
  lt_gemm<H>(sm, cx, Pu + wa, L::oBufA, false);
  if (warp < kLtEpiWarps) {
    lt_small_fetch<H>(sp, params + ar.off(1), ar.critic);
    lt_epilogue<H, 0>(sm, cx, L::oB + H, L::oBufB);
    bar_sync(1, kLtEpi);
    lt_out_layer<H>(sm, L::oBufB);
    if (t < kLtRows) {
      S[t * 8 + 2] = sm[L::oOut + t * 2]; S[t * 8 + 3] = sm[L::oOut + t * 2 + 1];
      in0[t * 4 + 2] = S[t * 8 + 2]; in0[t * 4 + 3] = S[t * 8 + 3];
    }
    bar_sync(1, kLtEpi);
    // ---- Q1(s, a): h0c -> A (h0a is recomputed later), h1c in place ----
    lt_small_store<H>(sm, sp);
    lt_layer0<H>(sm, L::oBufA);
  }
  lt_gemm<H>(sm, cx, Pu + wc, L::oBufA, false);
  if (warp < kLtEpiWarps) {
    lt_small_fetch<H>(sp, params + ar.off(0), ar.actor);     // for the backward pass through the actor
    lt_epilogue<H, 0>(sm, cx, L::oB + H, L::oBufA);
    bar_sync(1, kLtEpi);
    lt_out_layer<H>(sm, L::oBufA);
    if (t < kLtRows) {
      const float valid = S[t * 8 + 7];
      dout[t * 2] = -valid / (float)B;
      dout[t * 2 + 1] = 0.f;
      const float l = warp_sum(-sm[L::oOut + t * 2] * valid / (float)B);
      if ((t & 31) == 0) atomicAdd(loss, l);
    }
    bar_sync(1, kLtEpi);
    lt_bwd_out<H, 1>(sm, L::oBufA);          // dz1c in place (the critic's small-gradient sums land in the staging and are ignored)
  }
  // ---- dQ/da through critic 1: only the input gradient, nothing stored ----
  lt_gemm<H>(sm, cx, Pv + wc, L::oBufA, false);
  if (warp < kLtEpiWarps) {
    lt_epilogue_din<H>(sm, cx);
    bar_sync(1, kLtEpi);
    if (t < kLtRows) {
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) { d0 += sm[L::oPart + (g * kLtRows + t) * 2]; d1 += sm[L::oPart + (g * kLtRows + t) * 2 + 1]; }
      dout[t * 2] = d0;
      dout[t * 2 + 1] = d1;
      in0[t * 4 + 2] = 0.f; in0[t * 4 + 3] = 0.f;      // back to the actor's input
    }
    bar_sync(1, kLtEpi);
    // ---- backward through the actor ----
    lt_small_store<H>(sm, sp);
    lt_bwd_out<H, 2>(sm, L::oBufB);          // h1a -> dz1a, gWout, gb1
    lt_layer0<H>(sm, L::oBufA);              // recompute h0a
  }
  lt_dw<H>(sm, cx, L::oBufB, L::oBufA);
  lt_drain<H>(sm, cx, &maps.dw[0]);
  lt_gemm<H>(sm, cx, Pv + wa, L::oBufB, true);
  if (warp < kLtEpiWarps) {
    lt_epilogue<H, 1>(sm, cx, 0, L::oBufA);
    bar_sync(1, kLtEpi);
    lt_colsum_in<H, 2>(sm, L::oBufA);
    lt_reduce_small<H>(sm, grads + ar.off(0), ar.actor);
  }
  lt_teardown(cx);
}

}  // namespace rtd3

using namespace rtd3;

typedef CUresult (*LtEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
static LtEncodeTiledFn g_lt_encode = nullptr;

// [H n][H k] fp32 gradient block of a hidden-to-hidden layer, box 32 x 32, SWIZZLE_128B (the staging tile layout of lt_drain)
static int32_t make_dw_map(CUtensorMap* tm, float* base, int H) {
  if (!g_lt_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      rtd3::set_error("cuTensorMapEncodeTiled is not available from this driver");
      return RTD3_ERR_STATE;
    }
    g_lt_encode = (LtEncodeTiledFn)fn;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)H};
  const cuuint64_t strides[1] = {(cuuint64_t)H * 4};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = g_lt_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    rtd3::set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return RTD3_ERR_STATE;
  }
  return 0;
}

static bool lt_shape_ok(const rtd3_td3* h) { return h->ar.actor.layers == 2 && (h->ar.actor.hid == 128 || h->ar.actor.hid == 256); }


extern "C" {

/* Development aid: switch the phase timestamps of the tf32 critic kernel on/off; out (nullable) receives the 128 clock64 stamps. */
int32_t rtd3_debug_lt_prof(int32_t on, long long* out) {
  RTD3_CUDA(cudaMemcpyToSymbol(g_lt_prof_on, &on, sizeof(int)));
  if (out) RTD3_CUDA(cudaMemcpyFromSymbol(out, g_lt_prof, sizeof(long long) * 128));
  return 0;
}

int32_t rtd3_td3_tf32_supported(const rtd3_td3* h) { return (h && lt_shape_ok(h)) ? 1 : 0; }

int32_t rtd3_td3_critic_step_tf32(rtd3_td3* h, const float* params, const float* params_uv, float* grads, const float* rp_s, const float* rp_a,
                                  const float* rp_r, const float* rp_s2, const float* rp_notdone, const int32_t* idx, const float* noise,
                                  int32_t batch, float gamma, float policy_noise, float noise_clip, float max_action, float* loss2, float* q_out,
                                  float* y_out, int32_t* steps, double* beta_pows, void* stream) {
  RTD3_CHECK_ARG(h && params && params_uv && grads && rp_s && rp_a && rp_r && rp_s2 && rp_notdone && idx && noise && loss2 && steps && beta_pows,
                 "null argument");
  RTD3_CHECK_ARG(batch > 0, "batch must be positive");
  RTD3_CHECK_ARG(lt_shape_ok(h), "the tf32 learner needs layers == 2 and hidden in {128, 256}");
  ReplayView rp{(const float2*)rp_s, (const float2*)rp_a, rp_r, (const float2*)rp_s2, rp_notdone};
  Td3Hyper hp{gamma, policy_noise, noise_clip, max_action, 0ull, nullptr, 0ull};
  return critic_step_tc_launch(h, params, params_uv, grads, rp, idx, noise, batch, hp, loss2, q_out, y_out, steps, beta_pows, (cudaStream_t)stream);
}

}  // extern "C"

int32_t rtd3::critic_step_tc_launch(rtd3_td3* h, const float* params, const float* params_uv, float* grads, const ReplayView& rp, const int32_t* idx,
                                    const float* noise, int32_t batch, const Td3Hyper& hp, float* loss2, float* q_out, float* y_out, int32_t* steps,
                                    double* beta_pows, cudaStream_t st) {
  RTD3_CHECK_ARG(lt_shape_ok(h), "the tf32 learner needs layers == 2 and hidden in {128, 256}");
  const int H = h->ar.critic.hid;
  LtMaps maps;
  for (int c = 0; c < 2; ++c) {
    const int32_t rc = make_dw_map(&maps.dw[c], grads + h->ar.off(1 + c) + net_w_off(h->ar.critic, 1), H);
    if (rc) return rc;
  }
  const int grid = (batch + kLtRows - 1) / kLtRows;
  if (H == 128) {
    RTD3_CUDA(ensure_dyn_smem((const void*)td3_critic_tc_kernel<128>, Lt<128>::kBytes));
    td3_critic_tc_kernel<128><<<grid, kLtThreads, Lt<128>::kBytes, st>>>(h->ar, params, params_uv, grads, rp, idx, noise, batch, hp, loss2, q_out, y_out,
                                                                          steps, beta_pows, maps);
  } else {
    RTD3_CUDA(ensure_dyn_smem((const void*)td3_critic_tc_kernel<256>, Lt<256>::kBytes));
    td3_critic_tc_kernel<256><<<grid, kLtThreads, Lt<256>::kBytes, st>>>(h->ar, params, params_uv, grads, rp, idx, noise, batch, hp, loss2, q_out, y_out,
                                                                          steps, beta_pows, maps);
  }
  RTD3_LAUNCHED();
  return 0;
}

extern "C" {

int32_t rtd3_td3_actor_step_tf32(rtd3_td3* h, const float* params, const float* params_uv, float* grads, const float* rp_s, const int32_t* idx,
                                 int32_t batch, float* loss1, int32_t* steps, double* beta_pows, void* stream) {
  RTD3_CHECK_ARG(h && params && params_uv && grads && rp_s && idx && loss1 && steps && beta_pows, "null argument");
  RTD3_CHECK_ARG(batch > 0, "batch must be positive");
  RTD3_CHECK_ARG(lt_shape_ok(h), "the tf32 learner needs layers == 2 and hidden in {128, 256}");
  const int H = h->ar.actor.hid;
  LtMaps maps;
  for (int c = 0; c < 2; ++c) {
    const int32_t rc = make_dw_map(&maps.dw[c], grads + h->ar.off(0) + net_w_off(h->ar.actor, 1), H);
    if (rc) return rc;
  }
  ReplayView rp{(const float2*)rp_s, nullptr, nullptr, nullptr, nullptr};
  const int grid = (batch + kLtRows - 1) / kLtRows;
  cudaStream_t st = (cudaStream_t)stream;
  if (H == 128) {
    RTD3_CUDA(ensure_dyn_smem((const void*)td3_actor_tc_kernel<128>, Lt<128>::kBytes));
    td3_actor_tc_kernel<128><<<grid, kLtThreads, Lt<128>::kBytes, st>>>(h->ar, params, params_uv, grads, rp, idx, batch, loss1, steps, beta_pows, maps);
  } else {
    RTD3_CUDA(ensure_dyn_smem((const void*)td3_actor_tc_kernel<256>, Lt<256>::kBytes));
    td3_actor_tc_kernel<256><<<grid, kLtThreads, Lt<256>::kBytes, st>>>(h->ar, params, params_uv, grads, rp, idx, batch, loss1, steps, beta_pows, maps);
  }
  RTD3_LAUNCHED();
  return 0;
}

}  // extern "C"
