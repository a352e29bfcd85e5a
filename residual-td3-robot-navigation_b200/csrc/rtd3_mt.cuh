// numpy-legacy MT19937 on device: one stream per env, state word-major [624][n] so that a warp
// touching word k of 32 consecutive streams is one coalesced 128 B access.
// Algorithm: numpy/random/src/mt19937/mt19937.c (mt19937_seed, mt19937_gen) and
// legacy-distributions.c (legacy_gauss), distributions.c (random_interval) - restated, not copied.
#pragma once
#include "rtd3_common.cuh"

namespace rtd3 {

struct MtStream {
  uint32_t* w;      // word 0 of this stream; word k lives at w[k * stride]
  int64_t stride;   // = number of streams
  int pos;

  __device__ __forceinline__ uint32_t& at(int k) { return w[(int64_t)k * stride]; }

  __device__ void seed(uint32_t s) {
    for (int k = 0; k < RTD3_MT_N; ++k) {
      at(k) = s;
      s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)(k + 1);
    }
    pos = RTD3_MT_N;
  }

  __device__ void regenerate() {
    constexpr int N = RTD3_MT_N, M = 397;
    uint32_t first = at(0);
    uint32_t cur = first;
    for (int k = 0; k < N; ++k) {
      uint32_t nxt = (k + 1 < N) ? at(k + 1) : first;   // word 0 is read before it is overwritten (k = 0)
      uint32_t far = at(k + M < N ? k + M : k + M - N);
      uint32_t yv = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
      uint32_t v = far ^ (yv >> 1) ^ ((yv & 1u) ? 0x9908b0dfu : 0u);
      at(k) = v;
      if (k == 0) first = v;   // k = N-1 pairs with the NEW word 0 (mt19937_gen's last statement)
      cur = nxt;
    }
    pos = 0;
  }

  __device__ __forceinline__ uint32_t next_u32() {
    if (pos >= RTD3_MT_N) regenerate();
    uint32_t yv = at(pos++);
    yv ^= yv >> 11;
    yv ^= (yv << 7) & 0x9d2c5680u;
    yv ^= (yv << 15) & 0xefc60000u;
    yv ^= yv >> 18;
    return yv;
  }

  // random_double: 53-bit, exact in float64 whatever the contraction
  __device__ __forceinline__ double next_double() {
    uint32_t a = next_u32() >> 5, b = next_u32() >> 6;
    return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
  }

  // lo + (hi-lo)*u with numpy's two roundings (no fma contraction)
  __device__ __forceinline__ double uniform(double lo, double hi) {
    return __dadd_rn(lo, __dmul_rn(__dsub_rn(hi, lo), next_double()));
  }

  // random_interval, 32-bit masked rejection
  __device__ __forceinline__ uint32_t interval(uint32_t maxv) {
    if (maxv == 0) return 0;
    uint32_t mask = maxv;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    do { v = next_u32() & mask; } while (v > maxv);
    return v;
  }
};

// The twist of one stream by a whole warp (all 32 lanes must call it): the serial loop above costs ~200 us when a single lane of
// a warp runs it (624 dependent iterations of uncoalesced loads - it made EVERY masked reset of 65 536 envs take 200 us, because
// the streams' positions are spread over the whole cycle after the goal rejection loop and ~0.6 % of them wrap per call).
// Here the 624 words are held in registers, word k = 32 j + lane in v[j] of that lane: ONE round of 20 independent loads, the
// recurrence on warp shuffles, one round of stores (the first warp-wide version walked 32 words at a time through memory:
// 20 dependent L2 round trips, ~20 us per twist).  Word k needs the OLD words k, k+1 and word (k+397) % 624, which is old for
// k < 227 and already new for k >= 227 (k-227 lies 7-8 groups behind), and word 623 pairs with the NEW word 0: walking j upwards
// with every shuffle of a group taken before the group is overwritten reproduces the serial order exactly.
__device__ __forceinline__ void mt_regenerate_warp(uint32_t* w, int64_t stride) {
  constexpr int N = RTD3_MT_N, J = (RTD3_MT_N + 31) / 32;            // 20 groups; the last one holds words 608..623 in lanes 0..15
  constexpr uint32_t kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  uint32_t v[J];
#pragma unroll
  for (int j = 0; j < J; ++j) v[j] = (32 * j + lane < N) ? w[(int64_t)(32 * j + lane) * stride] : 0u;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int k = 32 * j + lane;
    uint32_t nxt = __shfl_down_sync(kFull, v[j], 1);                 // word k+1 for lanes 0..30 (old: taken before the write below)
    const uint32_t first_of_next = __shfl_sync(kFull, j + 1 < J ? v[j + 1] : 0u, 0);
    if (lane == 31) nxt = first_of_next;
    if (j == J - 1) {
      const uint32_t w0 = __shfl_sync(kFull, v[0], 0);               // word 623 pairs with the NEW word 0
      if (k == N - 1) nxt = w0;
    }
    uint32_t far;
    if (j <= 6) {                                                    // k + 397 < 624 for the whole group: old words (j+12, lane+13) / (j+13, lane-19)
      const int src = (lane + 13) & 31;
      const uint32_t a = __shfl_sync(kFull, v[j + 12], src), b = __shfl_sync(kFull, v[j + 13], src);
      far = lane + 13 < 32 ? a : b;
    } else if (j >= 8) {                                             // new words k - 227 = (j-8, lane+29) / (j-7, lane-3)
      const int src = (lane + 29) & 31;
      const uint32_t a = __shfl_sync(kFull, v[j - 8], src), b = __shfl_sync(kFull, v[j - 7], src);
      far = lane < 3 ? a : b;
    } else {                                                         // j == 7: words 224..226 still pair with old 621..623, the rest with new 0..28
      const uint32_t a = __shfl_sync(kFull, v[J - 1], (lane + 13) & 31), b = __shfl_sync(kFull, v[0], (lane + 29) & 31);
      far = lane < 3 ? a : b;
    }
    const uint32_t yv = (v[j] & 0x80000000u) | (nxt & 0x7fffffffu);
    v[j] = far ^ (yv >> 1) ^ ((yv & 1u) ? 0x9908b0dfu : 0u);
  }
#pragma unroll
  for (int j = 0; j < J; ++j)
    if (32 * j + lane < N) w[(int64_t)(32 * j + lane) * stride] = v[j];
  __syncwarp();
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t yv) {
  yv ^= yv >> 11;
  yv ^= (yv << 7) & 0x9d2c5680u;
  yv ^= (yv << 15) & 0xefc60000u;
  yv ^= yv >> 18;
  return yv;
}

// next_u32 for a warp whose lanes hold independent streams (lane-private `s`, `active` lanes draw): streams that have to wrap are
// twisted one after the other by the whole warp.  All 32 lanes must call it.
__device__ __forceinline__ uint32_t mt_next_u32_warp(MtStream& s, bool active) {
  uint32_t need = __ballot_sync(0xffffffffu, active && s.pos >= RTD3_MT_N);
  while (need) {
    const int src = __ffs(need) - 1;
    need &= need - 1;
    const unsigned long long wp = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)s.w, src);
    mt_regenerate_warp(reinterpret_cast<uint32_t*>((uintptr_t)wp), s.stride);
    if ((int)(threadIdx.x & 31) == src) s.pos = 0;
  }
  return active ? mt_temper(s.at(s.pos++)) : 0u;
}
// K consecutive words per active lane through ONE inlined copy of the twist (the loop is kept rolled: every inlined copy of
// mt_regenerate_warp costs ~1.5 KB of code and its 20 state registers at the call site)
template <int K>
__device__ __forceinline__ void mt_next_words_warp(MtStream& s, bool active, uint32_t (&out)[K]) {
#pragma unroll 1
  for (int h = 0; h < K; ++h) {
    const uint32_t t = mt_next_u32_warp(s, active);
#pragma unroll
    for (int q = 0; q < K; ++q)
      if (q == h) out[q] = t;
  }
}
// The same, out of line: for callers whose common path never wraps (reset_finish_warp) and should not carry the twist's registers.
static __device__ __noinline__ void mt_next_words4_warp_slow(uint32_t* w, int64_t stride, int* pos, bool active, uint32_t* out4) {
  MtStream s{w, stride, *pos};
  uint32_t o[4] = {0u, 0u, 0u, 0u};
  mt_next_words_warp<4>(s, active, o);
  *pos = s.pos;
#pragma unroll
  for (int q = 0; q < 4; ++q) out4[q] = o[q];
}
__device__ __forceinline__ double mt_double_from_words(uint32_t w0, uint32_t w1) {
  return ((double)(w0 >> 5) * 67108864.0 + (double)(w1 >> 6)) / 9007199254740992.0;
}
__device__ __forceinline__ double mt_next_double_warp(MtStream& s, bool active) {
  uint32_t w[2] = {0u, 0u};
  mt_next_words_warp<2>(s, active, w);
  return mt_double_from_words(w[0], w[1]);
}
// two doubles in draw order (uniform(lo, hi, 2): x first, then y)
__device__ __forceinline__ void mt_next_double2_warp(MtStream& s, bool active, double& u0, double& u1) {
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  mt_next_words_warp<4>(s, active, w);
  u0 = mt_double_from_words(w[0], w[1]);
  u1 = mt_double_from_words(w[2], w[3]);
}

// legacy_gauss with the spare kept in the bank. log/sqrt are float64 device libm (<= 1 ulp of glibc).
__device__ inline double mt_gauss(MtStream& s, int& has_gauss, double& spare) {
  if (has_gauss) {
    double t = spare;
    spare = 0.0;
    has_gauss = 0;
    return t;
  }
  double x1, x2, r2;
  do {
    x1 = __dsub_rn(__dmul_rn(2.0, s.next_double()), 1.0);
    x2 = __dsub_rn(__dmul_rn(2.0, s.next_double()), 1.0);
    r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
  } while (r2 >= 1.0 || r2 == 0.0);
  double f = sqrt(__ddiv_rn(__dmul_rn(-2.0, log(r2)), r2));
  spare = __dmul_rn(f, x1);
  has_gauss = 1;
  return __dmul_rn(f, x2);
}

// legacy_gauss for the lanes of a warp holding independent streams (all 32 lanes must call it; `active` lanes draw): the rejection
// loop runs until every active lane has its pair, wrapping streams are twisted by the whole warp.  hg / sp: the lane's spare.
__device__ __forceinline__ double mt_gauss_warp(MtStream& s, bool active, int& hg, double& sp) {
  double v = 0.0;
  bool searching = active && !hg;
  if (active && hg) { v = sp; sp = 0.0; hg = 0; }
  while (__any_sync(0xffffffffu, searching)) {
    double u1, u2;
    mt_next_double2_warp(s, searching, u1, u2);
    if (searching) {
      const double x1 = __dsub_rn(__dmul_rn(2.0, u1), 1.0), x2 = __dsub_rn(__dmul_rn(2.0, u2), 1.0);
      const double r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
      if (!(r2 >= 1.0 || r2 == 0.0)) {
        const double f = sqrt(__ddiv_rn(__dmul_rn(-2.0, log(r2)), r2));
        sp = __dmul_rn(f, x1);
        hg = 1;
        v = __dmul_rn(f, x2);
        searching = false;
      }
    }
  }
  return v;
}

// np.linalg.norm of a 2-vector as numpy evaluates it (sqrt(dot)): measured here to be
// sqrt(fma(dy,dy, dx*dx)) - 0 mismatches in 2e5 random pairs ("threshold compares").
__device__ __forceinline__ double norm2_np(double dx, double dy) { return sqrt(fma(dy, dy, __dmul_rn(dx, dx))); }

}  // namespace rtd3
