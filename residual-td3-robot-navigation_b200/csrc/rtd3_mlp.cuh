// CTA-level fp32 building blocks for the residual-TD3 networks (robot.py:128-206):
//   actor  2 -> H -> ... -> H -> 2      critic  4 -> H -> ... -> H -> 1      (L hidden layers, ReLU, linear head)
// A CTA of kThreads threads owns R consecutive batch rows and walks the whole layer chain for them; activations
// stay in shared memory, weights ([out][in] row-major, torch layout) stream from L2 through a double-buffered
// cp.async tile ring.  Thread t owns output column t of every hidden layer (H <= kThreads) and keeps R
// accumulators in registers.
#pragma once
#include "rtd3_common.cuh"

namespace rtd3 {

constexpr int kThreads = 256;
constexpr int kMaxHidden = 256;
constexpr int kKC = 32;              // reduction chunk staged per cp.async stage
constexpr int kLdW = kKC + 4;        // padded row stride (floats) of a row-pattern weight tile: conflict-free LDS.128
constexpr int kStageFloats = kMaxHidden * kLdW;   // 9216 floats = 36 KB per stage (>= kKC * kMaxHidden for the column pattern)

struct NetShape {
  int in, hid, layers, out;          // layers = number of hidden layers L (>= 1)
};

// flat parameter layout = torch's parameters() order: W1 [hid][in], b1, W2 [hid][hid], b2, ..., Wout [out][hid], bout
__host__ __device__ inline int64_t net_w_off(const NetShape& s, int l) {   // l = 0..layers (layers = output layer)
  if (l == 0) return 0;
  int64_t off = (int64_t)s.in * s.hid + s.hid;
  off += (int64_t)(l - 1) * ((int64_t)s.hid * s.hid + s.hid);
  return off;
}
__host__ __device__ inline int64_t net_b_off(const NetShape& s, int l) {
  const int fan_in = (l == 0) ? s.in : s.hid;
  const int fan_out = (l == s.layers) ? s.out : s.hid;
  return net_w_off(s, l) + (int64_t)fan_in * fan_out;
}
__host__ __device__ inline int64_t net_param_count(const NetShape& s) { return net_b_off(s, s.layers) + s.out; }
__host__ __device__ inline int64_t net_stride(const NetShape& s) { return (net_param_count(s) + 3) / 4 * 4; }   // 16 B aligned slots

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Shared-memory working set of one CTA (R rows).  ld = hid + 4 keeps rows 16 B aligned and skews banks.
template <int R>
struct MlpSmem {
  float* wst;        // [2][kStageFloats] weight tile ring
  float* act[4];     // act[l]: output of hidden layer l (R x ld); at most 4 hidden layers kept for backward
  float* dz[2];      // gradient w.r.t. pre-activations, ping-pong (R x ld)
  float* dzt;        // transposed copy [hid][R] of the current dz (feeds the weight-gradient reduction)
  float* in0;        // [R][4] network input
  float* out;        // [R][2] network output
  float* dout;       // [R][2] gradient w.r.t. network output
  float* din;        // [R][4] gradient w.r.t. network input
  float* scratch;    // [R][8] per-row scalars (reward, notdone, y, ...)
  int ld;

  __device__ static size_t bytes(int hid, int keep_layers) {
    const int ld = hid + 4;
    size_t f = 2 * (size_t)kStageFloats + (size_t)keep_layers * R * ld + 2 * (size_t)R * ld + (size_t)hid * R + R * 4 + R * 2 + R * 2 +
               R * 4 + R * 8;
    return f * sizeof(float);
  }
  __device__ void carve(float* base, int hid, int keep_layers) {
    ld = hid + 4;
    float* p = base;
    wst = p; p += 2 * kStageFloats;
    for (int l = 0; l < 4; ++l) { act[l] = (l < keep_layers) ? p : nullptr; if (l < keep_layers) p += R * ld; }
    dz[0] = p; p += R * ld;
    dz[1] = p; p += R * ld;
    dzt = p; p += hid * R;
    in0 = p; p += R * 4;
    out = p; p += R * 2;
    dout = p; p += R * 2;
    din = p; p += R * 4;
    scratch = p;
  }
};

__host__ inline size_t mlp_smem_bytes(int R, int hid, int keep_layers) {
  const int ld = hid + 4;
  size_t f = 2 * (size_t)kStageFloats + (size_t)keep_layers * R * ld + 2 * (size_t)R * ld + (size_t)hid * R + R * 4 + R * 2 + R * 2 + R * 4 +
             R * 8;
  return f * sizeof(float);
}

// ---- first layer: h[r][c] = relu(b[c] + sum_j in0[r][j] * W[c][j]),  in <= 4 ----------------------------------
template <int R>
__device__ __forceinline__ void fwd_first(const float* __restrict__ W, const float* __restrict__ b, const float* in0, int in_dim,
                                          float* Y, int ld, int N) {
  const int c = threadIdx.x;
  if (c < N) {
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < in_dim; ++j) w[j] = __ldg(W + c * in_dim + j);
    const float bias = __ldg(b + c);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 x = *reinterpret_cast<const float4*>(in0 + r * 4);
      float v = bias;
      v = fmaf(x.x, w[0], v); v = fmaf(x.y, w[1], v); v = fmaf(x.z, w[2], v); v = fmaf(x.w, w[3], v);
      Y[r * ld + c] = fmaxf(v, 0.f);
    }
  }
  __syncthreads();
}

// stage rows [0,N) x cols [k0,k0+kc) of a row-major [N][K] matrix into wst as [N][kLdW]
__device__ __forceinline__ void stage_rowpat(float* wst, const float* __restrict__ W, int N, int K, int k0, int kc) {
  const int q = kc >> 2;                         // float4 per row
  for (int idx = threadIdx.x; idx < N * q; idx += kThreads) {
    const int row = idx / q, j = idx - row * q;
    cp_async16(wst + row * kLdW + 4 * j, W + (int64_t)row * K + k0 + 4 * j);
  }
}
// stage rows [n0,n0+nc) x all K cols of a row-major [N][K] matrix into wst as [nc][K]
__device__ __forceinline__ void stage_colpat(float* wst, const float* __restrict__ W, int K, int n0, int nc) {
  const int q = K >> 2;
  for (int idx = threadIdx.x; idx < nc * q; idx += kThreads) {
    const int row = idx / q, j = idx - row * q;
    cp_async16(wst + row * K + 4 * j, W + (int64_t)(n0 + row) * K + 4 * j);
  }
}

// ---- hidden layer forward: Y[r][c] = relu(b[c] + sum_k X[r][k] * W[c][k]) -------------------------------------
template <int R>
__device__ __forceinline__ void fwd_hidden(const float* __restrict__ W, const float* __restrict__ b, const float* X, float* Y, int ld,
                                           int N, int K, float* wst) {
  const int c = threadIdx.x;
  float acc[R];
  const float bias = (c < N) ? __ldg(b + c) : 0.f;
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = bias;
  const int nchunks = (K + kKC - 1) / kKC;
  stage_rowpat(wst, W, N, K, 0, min(kKC, K));
  cp_commit();
  for (int ch = 0; ch < nchunks; ++ch) {
    const int k0 = ch * kKC, kc = min(kKC, K - k0);
    float* cur = wst + (ch & 1) * kStageFloats;
    if (ch + 1 < nchunks) {
      stage_rowpat(wst + ((ch + 1) & 1) * kStageFloats, W, N, K, k0 + kKC, min(kKC, K - k0 - kKC));
      cp_commit();
      cp_wait_one();
    } else {
      cp_wait_all();
    }
    __syncthreads();
    if (c < N) {
      const float* wrow = cur + c * kLdW;
      for (int kk = 0; kk < kc; kk += 4) {
        const float4 w = *reinterpret_cast<const float4*>(wrow + kk);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 x = *reinterpret_cast<const float4*>(X + r * ld + k0 + kk);   // warp-broadcast
          acc[r] = fmaf(x.x, w.x, acc[r]); acc[r] = fmaf(x.y, w.y, acc[r]);
          acc[r] = fmaf(x.z, w.z, acc[r]); acc[r] = fmaf(x.w, w.w, acc[r]);
        }
      }
    }
    __syncthreads();
  }
  if (c < N) {
#pragma unroll
    for (int r = 0; r < R; ++r) Y[r * ld + c] = fmaxf(acc[r], 0.f);
  }
  __syncthreads();
}

// ---- output layer: out[r][o] = b[o] + sum_k X[r][k] * W[o][k]  (out_dim <= 2): one warp per (r,o) pair ----------
template <int R>
__device__ __forceinline__ void fwd_out(const float* __restrict__ W, const float* __restrict__ b, const float* X, int ld, int K,
                                        int out_dim, float* out /*[R][2]*/) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int p = warp; p < R * out_dim; p += kThreads / 32) {
    const int r = p / out_dim, o = p - r * out_dim;
    float v = 0.f;
    for (int k = lane; k < K; k += 32) v = fmaf(X[r * ld + k], __ldg(W + o * K + k), v);
    v = warp_sum(v);
    if (lane == 0) out[r * 2 + o] = v + __ldg(b + o);
  }
  __syncthreads();
}

// whole network forward for the CTA's R rows.  keep=true stores hidden layer l in sm.act[l] (needed by backward);
// keep=false ping-pongs between sm.dz[0] / sm.dz[1] (free during a forward pass).  Result in sm.out.
template <int R>
__device__ __forceinline__ const float* mlp_forward(const float* __restrict__ P, const NetShape& s, MlpSmem<R>& sm, bool keep) {
  float* y0 = keep ? sm.act[0] : sm.dz[0];
  fwd_first<R>(P + net_w_off(s, 0), P + net_b_off(s, 0), sm.in0, s.in, y0, sm.ld, s.hid);
  const float* x = y0;
  for (int l = 1; l < s.layers; ++l) {
    float* y = keep ? sm.act[l] : sm.dz[l & 1];
    fwd_hidden<R>(P + net_w_off(s, l), P + net_b_off(s, l), x, y, sm.ld, s.hid, s.hid, sm.wst);
    x = y;
  }
  fwd_out<R>(P + net_w_off(s, s.layers), P + net_b_off(s, s.layers), x, sm.ld, s.hid, s.out, sm.out);
  return x;
}

// ---- backward -------------------------------------------------------------------------------------------------
// Output layer: dz_L[r][k] = relu'(h_L[r][k]) * sum_o dout[r][o] * Wout[o][k];  gWout[o][k] += sum_r dout[r][o] * h_L[r][k]
template <int R>
__device__ __forceinline__ void bwd_out(const float* __restrict__ W, float* __restrict__ gW, float* __restrict__ gb, const float* H,
                                        int ld, int K, int out_dim, const float* dout, float* dz, float* dzt) {
  const int k = threadIdx.x;
  if (k < K) {
    float w[2] = {__ldg(W + k), out_dim > 1 ? __ldg(W + K + k) : 0.f};
    float g0 = 0.f, g1 = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float h = H[r * ld + k];
      const float d0 = dout[r * 2], d1 = dout[r * 2 + 1];
      g0 = fmaf(d0, h, g0);
      g1 = fmaf(d1, h, g1);
      float d = fmaf(d0, w[0], d1 * w[1]);
      d = h > 0.f ? d : 0.f;
      dz[r * ld + k] = d;
      dzt[k * R + r] = d;
    }
    if (gW) {
      atomicAdd(gW + k, g0);
      if (out_dim > 1) atomicAdd(gW + K + k, g1);
    }
  }
  if (gb && k < out_dim) {
    float g = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) g += dout[r * 2 + k];
    atomicAdd(gb + k, g);
  }
  __syncthreads();
}

// Weight gradient of a hidden layer: gW[n][k] += sum_r dz[r][n] * X[r][k];  gb[n] += sum_r dz[r][n]
template <int R>
__device__ __forceinline__ void bwd_weights(float* __restrict__ gW, float* __restrict__ gb, const float* X, int ld, int N, int K,
                                            const float* dzt) {
  const int k = threadIdx.x;
  if (k < K) {
    float x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = X[r * ld + k];
    for (int n = 0; n < N; ++n) {
      float v = 0.f;
#pragma unroll
      for (int r4 = 0; r4 < R; r4 += 4) {
        const float4 d = *reinterpret_cast<const float4*>(dzt + n * R + r4);   // warp-broadcast
        v = fmaf(d.x, x[r4], v); v = fmaf(d.y, x[r4 + 1], v); v = fmaf(d.z, x[r4 + 2], v); v = fmaf(d.w, x[r4 + 3], v);
      }
      atomicAdd(gW + (int64_t)n * K + k, v);     // RED.ADD.F32, coalesced across the warp
    }
  }
  const int n = threadIdx.x;
  if (n < N) {
    float g = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) g += dzt[n * R + r];
    atomicAdd(gb + n, g);
  }
}

// Input gradient of a hidden layer: dzp[r][k] = relu'(Hprev[r][k]) * sum_n dz[r][n] * W[n][k]
template <int R>
__device__ __forceinline__ void bwd_input(const float* __restrict__ W, const float* dz, const float* Hprev, float* dzp, float* dzt,
                                          int ld, int N, int K, float* wst) {
  const int k = threadIdx.x;
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  const int nchunks = (N + kKC - 1) / kKC;
  stage_colpat(wst, W, K, 0, min(kKC, N));
  cp_commit();
  for (int ch = 0; ch < nchunks; ++ch) {
    const int n0 = ch * kKC, nc = min(kKC, N - n0);
    float* cur = wst + (ch & 1) * kStageFloats;
    if (ch + 1 < nchunks) {
      stage_colpat(wst + ((ch + 1) & 1) * kStageFloats, W, K, n0 + kKC, min(kKC, N - n0 - kKC));
      cp_commit();
      cp_wait_one();
    } else {
      cp_wait_all();
    }
    __syncthreads();
    if (k < K) {
      for (int nn = 0; nn < nc; nn += 4) {
        const float w0 = cur[(nn + 0) * K + k], w1 = cur[(nn + 1) * K + k], w2 = cur[(nn + 2) * K + k], w3 = cur[(nn + 3) * K + k];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 d = *reinterpret_cast<const float4*>(dz + r * ld + n0 + nn);   // warp-broadcast
          acc[r] = fmaf(d.x, w0, acc[r]); acc[r] = fmaf(d.y, w1, acc[r]);
          acc[r] = fmaf(d.z, w2, acc[r]); acc[r] = fmaf(d.w, w3, acc[r]);
        }
      }
    }
    __syncthreads();
  }
  if (k < K) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float d = Hprev[r * ld + k] > 0.f ? acc[r] : 0.f;
      dzp[r * ld + k] = d;
      dzt[k * R + r] = d;
    }
  }
  __syncthreads();
}

// First layer: gW1[c][j] += sum_r dz[r][c] * in0[r][j]; gb1[c] += sum_r dz[r][c];
// optionally din[r][j] = sum_c dz[r][c] * W1[c][j]  (one warp per (r,j) pair)
template <int R>
__device__ __forceinline__ void bwd_first(const float* __restrict__ W, float* __restrict__ gW, float* __restrict__ gb, const float* in0,
                                          int in_dim, const float* dz, int ld, int N, float* din) {
  const int c = threadIdx.x;
  if (gW && c < N) {
    float g[4] = {0.f, 0.f, 0.f, 0.f}, gbias = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float d = dz[r * ld + c];
      const float4 x = *reinterpret_cast<const float4*>(in0 + r * 4);
      g[0] = fmaf(d, x.x, g[0]); g[1] = fmaf(d, x.y, g[1]); g[2] = fmaf(d, x.z, g[2]); g[3] = fmaf(d, x.w, g[3]);
      gbias += d;
    }
    for (int j = 0; j < in_dim; ++j) atomicAdd(gW + c * in_dim + j, g[j]);
    atomicAdd(gb + c, gbias);
  }
  if (din) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int p = warp; p < R * in_dim; p += kThreads / 32) {
      const int r = p / in_dim, j = p - r * in_dim;
      float v = 0.f;
      for (int cc = lane; cc < N; cc += 32) v = fmaf(dz[r * ld + cc], __ldg(W + cc * in_dim + j), v);
      v = warp_sum(v);
      if (lane == 0) din[r * 4 + j] = v;
    }
  }
  __syncthreads();
}

// Whole-network backward from sm.dout (gradient w.r.t. the output, [R][2]).  Needs the activations kept by
// mlp_forward(keep=true).  G == nullptr skips parameter gradients (the actor loss only needs dQ/da through critic 1).
template <int R>
__device__ __forceinline__ void mlp_backward(const float* __restrict__ P, float* __restrict__ G, const NetShape& s, MlpSmem<R>& sm,
                                             bool want_din) {
  const int L = s.layers;
  int cur = 0;
  bwd_out<R>(P + net_w_off(s, L), G ? G + net_w_off(s, L) : nullptr, G ? G + net_b_off(s, L) : nullptr, sm.act[L - 1], sm.ld, s.hid,
             s.out, sm.dout, sm.dz[cur], sm.dzt);
  for (int l = L - 1; l >= 1; --l) {
    if (G) bwd_weights<R>(G + net_w_off(s, l), G + net_b_off(s, l), sm.act[l - 1], sm.ld, s.hid, s.hid, sm.dzt);
    __syncthreads();     // everyone is done with dzt before bwd_input rewrites it
    bwd_input<R>(P + net_w_off(s, l), sm.dz[cur], sm.act[l - 1], sm.dz[cur ^ 1], sm.dzt, sm.ld, s.hid, s.hid, sm.wst);
    cur ^= 1;
  }
  bwd_first<R>(P + net_w_off(s, 0), G ? G + net_w_off(s, 0) : nullptr, G ? G + net_b_off(s, 0) : nullptr, sm.in0, s.in, sm.dz[cur], sm.ld,
               s.hid, want_din ? sm.din : nullptr);
}

}  // namespace rtd3
