// ReplayBuffer.sample's index draw (robot.py:111): np.random.choice(len, B, replace=False) on a numpy-legacy
// MT19937 stream = the first B entries of a legacy Fisher-Yates shuffle of arange(len)
// (RandomState.shuffle: for i = len-1 .. 1: j = random_interval(i); swap(x[i], x[j])).  Bit-exact with numpy;
// `count` consecutive samples are drawn per call (a TD3 update needs 150 of them: robot.py:272-285).
//
// The shuffle looks serial, but only two small parts of it are:
//   1. which raw 32-bit draws are accepted by the masked rejection loop of random_interval - a chain through the
//      stream because the bound i drops by one per accepted draw.  One warp resolves 32 draws at a time: every lane
//      assumes a set of accepted lower lanes, derives its own bound, and the ballot is iterated to its fixed point
//      (lane L's decision depends only on lanes < L, so the fixed point is the sequential answer).
//   2. nothing else: the B outputs do not need the permutation.  Output p is found by walking the swap list
//      BACKWARDS from position p (pos==i -> j_i, pos==j_i -> i), independently per output and per sample.
// Kernel A (1 CTA) writes the swap lists J[s][i]; kernel B (one CTA per sample) back-traces the B outputs.
#include "rtd3_common.cuh"
#include "rtd3_mt.cuh"

namespace rtd3 {

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

__device__ __forceinline__ uint32_t mt_twist(uint32_t cur, uint32_t nxt, uint32_t far) {
  const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
  return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// mt19937_gen by one warp, in place, 32 words at a time: read, __syncwarp, write.  The three ranges differ only in
// where word kk+397 (mod 624) comes from; word 623 pairs with the NEW word 0, as in the reference implementation.
__device__ void mt_regenerate_warp(uint32_t* mt, int lane) {
  constexpr int N = RTD3_MT_N, M = 397;
  for (int base = 0; base < N; base += 32) {
    const int kk = base + lane;
    uint32_t v = 0;
    if (kk < N) {
      const uint32_t cur = mt[kk];
      const uint32_t nxt = mt[kk + 1 < N ? kk + 1 : 0];      // kk = 623 reads word 0, already regenerated
      const uint32_t far = mt[kk + M < N ? kk + M : kk + M - N];
      v = mt_twist(cur, nxt, far);
    }
    __syncwarp();
    if (kk < N) mt[kk] = v;
    __syncwarp();
  }
}

// Kernel A: swap lists.  J[s*n + i] = j_i for i = 1..n-1 (entry 0 unused).
__global__ void __launch_bounds__(32)
sample_swaps_kernel(rtd3_mt_bank b, int64_t stream_id, int32_t n, int32_t count, int32_t* __restrict__ J) {
  __shared__ uint32_t mt[RTD3_MT_N];
  const int lane = threadIdx.x;
  for (int k = lane; k < RTD3_MT_N; k += 32) mt[k] = b.mt[(int64_t)k * b.n + stream_id];
  int pos = b.pos[stream_id];
  __syncwarp();
  const uint32_t lt = (1u << lane) - 1u;
  for (int s = 0; s < count; ++s) {
    int32_t* Js = J + (int64_t)s * n;
    int i = n - 1;                                  // swaps still to draw: indices i, i-1, ..., 1
    while (i >= 1) {
      if (pos >= RTD3_MT_N) {
        mt_regenerate_warp(mt, lane);
        pos = 0;
      }
      const int g = min(32, RTD3_MT_N - pos);
      const bool active = lane < g;
      const uint32_t raw = active ? mt_temper(mt[pos + lane]) : 0u;
      // accepted(L) = (raw_L & mask(i_L)) <= i_L with i_L = i - #accepted lanes below L (i_L >= 1).
      // Fast path: evaluate it under the two extreme assumptions (no lower lane accepted / every lower lane accepted);
      // the truth lies between them, so if both ballots agree that is the answer (all but ~1 % of the groups).
      // Otherwise iterate the ballot to its fixed point.
      auto decide = [&](uint32_t assumed, int& bound, uint32_t& val) {
        bound = i - __popc(assumed & lt);
        const uint32_t mask = bound >= 1 ? (0xffffffffu >> __clz(bound)) : 0u;
        val = raw & mask;
        return __ballot_sync(0xffffffffu, active && bound >= 1 && val <= (uint32_t)bound);
      };
      int my_i, bi;
      uint32_t v, bv;
      const uint32_t acc_hi = decide(0u, my_i, v);
      const uint32_t acc_lo = decide(0xffffffffu, bi, bv);
      uint32_t acc = acc_hi;
      // (the sandwich argument needs one mask for the whole group: no power-of-two crossing within reach of i)
      const bool one_mask = __clz(i) == __clz(max(i - 31, 1));
      if (acc_lo != acc_hi || !one_mask) {
        uint32_t prev;
        do {
          prev = acc;
          acc = decide(prev, my_i, v);
        } while (acc != prev);
      } else {
        acc = decide(acc_hi, my_i, v);             // bounds / values under the agreed set
      }
      if (acc & (1u << lane)) Js[my_i] = (int32_t)v;
      const int n_acc = __popc(acc);
      if (n_acc >= i) {
        // the sample completes inside this group: the accepted lane with bound 1 is its last draw
        const uint32_t last = __ballot_sync(0xffffffffu, (acc & (1u << lane)) && my_i == 1);
        pos += (31 - __clz(last)) + 1;              // later draws of the group belong to the next sample
        i = 0;
      } else {
        pos += g;
        i -= n_acc;
      }
    }
  }
  __syncwarp();
  for (int k = lane; k < RTD3_MT_N; k += 32) b.mt[(int64_t)k * b.n + stream_id] = mt[k];
  if (lane == 0) b.pos[stream_id] = pos;
}

// Kernel B: out[s][p] = x[p] after the shuffle, by unwinding the swaps from position p.
__global__ void __launch_bounds__(256)
sample_trace_kernel(const int32_t* __restrict__ J, int32_t n, int32_t batch, int32_t* __restrict__ out) {
  extern __shared__ __align__(16) int32_t js[];     // swap list of this sample (n entries, padded to a multiple of 4)
  const int s = blockIdx.x;
  const int32_t* Js = J + (int64_t)s * n;
  const int n4 = (n + 3) & ~3;
  for (int k = threadIdx.x; k < n4; k += blockDim.x) js[k] = (k >= 1 && k < n) ? Js[k] : k;   // entries 0 and >= n are self-swaps
  __syncthreads();
  for (int p = threadIdx.x; p < batch; p += blockDim.x) {
    int pos = p;
    for (int i = 0; i < n4; i += 4) {
      const int4 j = *reinterpret_cast<const int4*>(js + i);
      pos = (pos == i + 0) ? j.x : ((pos == j.x) ? i + 0 : pos);
      pos = (pos == i + 1) ? j.y : ((pos == j.y) ? i + 1 : pos);
      pos = (pos == i + 2) ? j.z : ((pos == j.z) ? i + 2 : pos);
      pos = (pos == i + 3) ? j.w : ((pos == j.w) ? i + 3 : pos);
    }
    out[(int64_t)s * batch + p] = pos;
  }
}

// Philox4x32-10 (Salmon et al. 2011), the counter-based generator used for the non-parity index draw.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

__global__ void sample_philox_kernel(uint64_t seed, uint64_t offset, uint32_t n, int64_t total, int32_t* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (4 * t >= total) return;
  const uint64_t c = offset + (uint64_t)t;
  const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (4 * t + k < total) out[4 * t + k] = (int32_t)__umulhi(v[k], n);   // multiply-shift map to [0, n)
}

}  // namespace rtd3

using namespace rtd3;

extern "C" int32_t rtd3_sample_indices_philox(uint64_t seed, uint64_t offset, int64_t n, int32_t batch, int32_t count, int32_t* out,
                                              void* stream) {
  RTD3_CHECK_ARG(out && n >= 1 && n < (1ll << 31) && batch >= 1 && count >= 0, "bad argument");
  const int64_t total = (int64_t)batch * count;
  if (total == 0) return 0;
  sample_philox_kernel<<<(int)ceil_div(ceil_div(total, 4), 256), 256, 0, (cudaStream_t)stream>>>(seed, offset, (uint32_t)n, total, out);
  RTD3_LAUNCHED();
  return 0;
}


extern "C" int32_t rtd3_sample_indices_mt19937(const rtd3_mt_bank* bank, int64_t stream_id, int32_t n, int32_t batch, int32_t count,
                                               int32_t* out, int32_t* scratch, void* stream) {
  RTD3_CHECK_ARG(bank && bank->mt && bank->pos && out, "null argument");
  RTD3_CHECK_ARG(stream_id >= 0 && stream_id < bank->n, "stream id out of range");
  RTD3_CHECK_ARG(n >= 1 && batch >= 1 && batch <= n && count >= 0, "need 1 <= batch <= n");
  RTD3_CHECK_ARG(n <= 56000, "replay sizes above 56000 rows are not supported by the exact sampler");
  if (count == 0) return 0;
  RTD3_CHECK_ARG(scratch, "scratch (count * n int32) is required");
  static bool attr_set = false;
  if (!attr_set) {
    RTD3_CUDA(cudaFuncSetAttribute(sample_trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 56000 * 4));
    attr_set = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  sample_swaps_kernel<<<1, 32, 0, st>>>(*bank, stream_id, n, count, scratch);
  RTD3_LAUNCHED();
  const size_t smem = (size_t)((n + 3) & ~3) * 4;
  sample_trace_kernel<<<count, 256, smem, st>>>(scratch, n, batch, out);
  RTD3_LAUNCHED();
  return 0;
}
