// Building blocks of the f16 resident-weight network forward (layers == 2), shared by mlp_forward_f16_kernel (rtd3_tc.cu) and
// the multi-tick kernel (rtd3_tick.cu).  One CTA = 16 row warps + one MMA-issuing warp; a tile = 128 batch rows = UMMA M:
//   first layer (in <= 4) in fp32 FFMA -> X [H/8][128][8] half in shared memory (UMMA no-swizzle K-major chunk layout)
//   hidden layer: H/16 tcgen05.mma.kind::f16 against W2, RESIDENT in shared memory as Wr [H/8][H][8] half, fp32 accumulator in TMEM
//   epilogue: bias, ReLU and the output layer straight from the accumulator (row thread x column part), partial sums through `part`.
// fp16 carries the same 11-bit significand as TF32 in half the bytes: 128 KB of W2 (H = 256) fit next to a 64 KB X tile, so nothing
// is streamed per tile (the TF32 kernel's tile time is its 256 KB weight stream).  Both kernels run exactly this code per tile, so
// their outputs are bit-identical.
#pragma once
#include <cuda_fp16.h>

#include "rtd3_common.cuh"
#include "rtd3_mlp.cuh"
#include "rtd3_tc.cuh"

namespace rtd3 {

constexpr int kHfRows = 128;                          // batch rows per tile = UMMA M
constexpr int kHfRowWarps = 16;                       // four per TMEM sub-partition: warp w serves rows 32*(w%4).. and column part w/4
constexpr int kHfColParts = kHfRowWarps / 4;
constexpr int kHfThreads = (kHfRowWarps + 1) * 32;    // + the MMA warp
constexpr int kHfRowMma = kHfThreads;                 // named barrier 2: row warps + MMA warp
constexpr int kHfRowThreads = kHfRowWarps * 32;       // named barrier 1: row warps

struct HfSmem {
  unsigned char* Xs;        // [H/8][128 rows][8] half
  unsigned char* Wr;        // [H/8][H rows n][8] half, resident
  float* F1;                // [H][8]: in <= 2: {w0, w1, b1, 0, ...}; else {w0..w3, b1, 0, 0, 0}   (one 16 B broadcast load per column)
  float* E2;                // [H][4]: {b2, Wo[0][c], Wo[1][c], 0}
  float* bo;                // [2] (+2 pad)
  uint64_t* acc_ready;      // [2]
  uint64_t* w_full;
  uint32_t* tmem_slot;
  float* part;              // [2 buffers][column parts][128 rows][2] output-layer partial sums
  float* extra;             // kernel-specific tail (hf_smem_bytes(H) + what the kernel adds)
};

__host__ __device__ inline size_t hf_smem_bytes(int hid) {
  return (size_t)hid * kHfRows * 2 + (size_t)hid * hid * 2 + ((size_t)hid * 8 + (size_t)hid * 4 + 4) * 4 + 64 + 2 * kHfColParts * kHfRows * 2 * 4;
}

__device__ __forceinline__ HfSmem hf_carve(unsigned char* smem_raw, int H) {
  HfSmem m;
  m.Xs = smem_raw;
  m.Wr = m.Xs + (size_t)H * kHfRows * 2;
  m.F1 = reinterpret_cast<float*>(m.Wr + (size_t)H * H * 2);
  m.E2 = m.F1 + H * 8;
  m.bo = m.E2 + H * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(m.bo + 4);
  m.acc_ready = bars;
  m.w_full = bars + 2;
  m.tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  m.part = reinterpret_cast<float*>(bars + 8);
  m.extra = m.part + 2 * kHfColParts * kHfRows * 2;
  return m;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}

__device__ __forceinline__ uint32_t pack_half2_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // first source -> upper half
  return r;
}

// Barriers, TMEM allocation, the resident weight (issued as 8 bulk copies on w_full, in flight during the rest of the setup) and
// the per-column parameter rows.  All threads of the CTA call it; returns the TMEM base address.
__device__ __forceinline__ uint32_t hf_setup(const HfSmem& m, const NetShape& s, const float* __restrict__ P, const uint16_t* __restrict__ Wh,
                                             uint32_t tmem_cols) {
  const int H = s.hid, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  if (t == 0) {
    mbar_init(m.acc_ready, 1);
    mbar_init(m.acc_ready + 1, 1);
    mbar_init(m.w_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(m.tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();                                                    // barriers initialised before the weight copies are issued
  if (warp == kHfRowWarps && lane == 0) {
    const uint32_t bytes = (uint32_t)H * H * 2, chunk = bytes / 8;
    mbar_arrive_expect_tx(m.w_full, bytes);
    for (int c = 0; c < 8; ++c) bulk_g2s(m.Wr + (size_t)c * chunk, reinterpret_cast<const unsigned char*>(Wh) + (size_t)c * chunk, chunk, m.w_full);
  }
  const bool in2 = s.in <= 2;
  for (int i = t; i < H * 8; i += kHfThreads) {
    const int c = i >> 3, j = i & 7;
    float v = 0.f;
    if (j < s.in) v = __ldg(P + net_w_off(s, 0) + c * s.in + j);
    else if (j == (in2 ? 2 : 4)) v = __ldg(P + net_b_off(s, 0) + c);
    m.F1[i] = v;
  }
  for (int i = t; i < H * 4; i += kHfThreads) {
    const int c = i >> 2, j = i & 3;
    float v = 0.f;
    if (j == 0) v = __ldg(P + net_b_off(s, 1) + c);
    else if (j - 1 < s.out) v = __ldg(P + net_w_off(s, 2) + (j - 1) * H + c);
    m.E2[i] = v;
  }
  if (t < 2) m.bo[t] = t < s.out ? __ldg(P + net_b_off(s, 2) + t) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *m.tmem_slot;
}

// instruction descriptor: D = F32, A = B = F16 (format 0), both K-major, N = H, M = 128
__device__ __forceinline__ uint32_t hf_idesc(int H) { return (1u << 4) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(kHfRows >> 4) << 24); }

// One thread: the H/16 products of a tile into accumulator `acc`, then the commit that signals `ready`.
__device__ __forceinline__ void hf_issue_tile(const HfSmem& m, int H, uint32_t acc, uint32_t idesc, uint64_t* ready) {
  tc_fence_after();
  for (int k16 = 0; k16 < H / 16; ++k16) {
    const uint64_t ad = umma_desc_kmajor(smem_u32(m.Xs + (size_t)(2 * k16) * (kHfRows * 16)), kHfRows * 16, 128);
    const uint64_t bd = umma_desc_kmajor(smem_u32(m.Wr + (size_t)(2 * k16) * ((size_t)H * 16)), (uint32_t)H * 16, 128);
    umma_f16(acc, ad, bd, idesc, k16 != 0 ? 1u : 0u);
  }
  umma_commit(ready);
}

// Row thread: X[rt][c_lo..c_hi) = half(relu(b1 + x0 W1^T)) in the chunk layout, made visible to the tensor core.  The caller then
// meets the MMA warp at named barrier 2.
__device__ __forceinline__ void hf_first_layer(const HfSmem& m, const float (&x0)[4], bool in2, int rt, int c_lo, int c_hi) {
  for (int c = c_lo; c < c_hi; c += 8) {
    float h[8];
    if (in2) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(m.F1 + (c + q) * 8);
        h[q] = fmaxf(fmaf(x0[1], w.y, fmaf(x0[0], w.x, w.z)), 0.f);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(m.F1 + (c + q) * 8);
        float v = m.F1[(c + q) * 8 + 4];
        v = fmaf(x0[0], w.x, v); v = fmaf(x0[1], w.y, v); v = fmaf(x0[2], w.z, v); v = fmaf(x0[3], w.w, v);
        h[q] = fmaxf(v, 0.f);
      }
    }
    uint4 pk;
    pk.x = pack_half2_sat(h[0], h[1]); pk.y = pack_half2_sat(h[2], h[3]); pk.z = pack_half2_sat(h[4], h[5]); pk.w = pack_half2_sat(h[6], h[7]);
    *reinterpret_cast<uint4*>(m.Xs + ((size_t)(c >> 3) * kHfRows + rt) * 16) = pk;
  }
  fence_proxy_async();                            // generic-proxy writes of X -> visible to the tensor core (async proxy)
  tc_fence_before();
}

// Row thread: partial output-layer sums of its column part, y_part = Wo[:, c_lo..c_hi) relu(acc + b2), written to part buffer `buf`.
// `acc` already carries this warp's TMEM lane offset.
__device__ __forceinline__ void hf_epilogue(const HfSmem& m, uint32_t acc, int rt, int cpart, int c_lo, int c_hi, int buf) {
  float o0 = 0.f, o1 = 0.f;
  for (int cb = c_lo; cb < c_hi; cb += 32) {
    float v[32];
    tmem_ld32(acc + (uint32_t)cb, v);
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      const float4 e = *reinterpret_cast<const float4*>(m.E2 + (cb + q) * 4);
      const float h = fmaxf(v[q] + e.x, 0.f);
      o0 = fmaf(h, e.y, o0);
      o1 = fmaf(h, e.z, o1);
    }
  }
  tc_fence_before();
  float* pt = m.part + buf * (kHfColParts * kHfRows * 2);
  pt[(cpart * kHfRows + rt) * 2] = o0;             // warps without columns contribute zeros
  pt[(cpart * kHfRows + rt) * 2 + 1] = o1;
}

// After named barrier 1: the network output of row rt.
__device__ __forceinline__ float2 hf_output(const HfSmem& m, int rt, int buf) {
  const float* pt = m.part + buf * (kHfColParts * kHfRows * 2);
  float y0 = m.bo[0], y1 = m.bo[1];
#pragma unroll
  for (int c = 0; c < kHfColParts; ++c) { y0 += pt[(c * kHfRows + rt) * 2]; y1 += pt[(c * kHfRows + rt) * 2 + 1]; }
  return make_float2(y0, y1);
}

// Column range of a row warp: quarters when every quarter is a whole number of 32-column TMEM loads, else halves or the whole row
// (the surplus warps of each sub-partition then idle through the barriers with an empty range).
__device__ __forceinline__ void hf_columns(int H, int cpart, int& c_lo, int& c_hi) {
  const int parts = (H % (32 * kHfColParts) == 0) ? kHfColParts : ((H % 64 == 0) ? 2 : 1);
  c_lo = cpart < parts ? cpart * (H / parts) : 0;
  c_hi = cpart < parts ? c_lo + H / parts : 0;
}

}  // namespace rtd3
