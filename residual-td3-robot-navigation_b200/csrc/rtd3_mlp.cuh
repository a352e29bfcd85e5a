// CTA-level fp32 building blocks for the residual-TD3 networks (robot.py:128-206):
//   actor  2 -> H -> ... -> H -> 2      critic  4 -> H -> ... -> H -> 1      (L hidden layers, ReLU, linear head)
// A CTA of kThreads threads owns R consecutive batch rows and walks the whole layer chain for them; activations
// stay in shared memory.  Hidden-layer products are out[r][c] = sum_j X[r][j] * M[j][c] with M row-major and c
// contiguous: the forward pass reads the TRANSPOSED weight copy Wt [in][out] (kept next to the torch-layout
// parameters by the optimiser kernel), the backward pass reads the torch layout W [out][in] itself.  A thread owns 4
// consecutive columns (one coalesced 16 B load per weight row, straight from L2 - no shared-memory staging, which
// made LDGSTS + LDS traffic the bottleneck of the first version) and a quarter of the j range; the four partial
// sums meet in shared memory once per layer.  Weight gradients are NOT formed here: forward/backward store the per-row layer
// inputs and pre-activation gradients to an L2-resident scratch, and wgrad_kernel (rtd3_td3.cu) reduces them over
// the batch tile by tile without atomics.
#pragma once
#include "rtd3_common.cuh"

namespace rtd3 {

constexpr int kThreads = 256;
constexpr int kMaxHidden = 256;
constexpr int kMaxLayers = 4;
constexpr int kGroups = 4;           // thread groups splitting the reduction range of a hidden layer
constexpr int kColThreads = kThreads / kGroups;   // 64 column groups x 4 columns = 256 columns

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

// Per-row scratch of one trained network (global memory, L2 resident): what the weight-gradient pass needs.
//   in0 [B][4] | dout [B][2] | h[l] [B][hid] (output of hidden layer l) | dz[l] [B][hid] (grad of its pre-activation)
struct RowScratch {
  float* base;
  int B, hid, layers;
  // sections start at multiples of 4 rows so that every row stays 16 B aligned for float4 access
  __host__ __device__ static int64_t pad(int B) { return ((int64_t)B + 3) / 4 * 4; }
  __host__ __device__ static int64_t floats(int B, int hid, int layers) { return pad(B) * (8 + 2 * (int64_t)layers * hid); }
  __host__ __device__ float* in0() const { return base; }
  __host__ __device__ float* dout() const { return base + pad(B) * 4; }
  __host__ __device__ float* h(int l) const { return base + pad(B) * 8 + (int64_t)l * pad(B) * hid; }
  __host__ __device__ float* dz(int l) const { return base + pad(B) * 8 + (int64_t)(layers + l) * pad(B) * hid; }
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kN>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kN) : "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Shared-memory working set of one CTA (R rows), addressed as OFFSETS (in floats) into the one dynamic shared array
// `smem_f`: every access below is smem_f[offset + i], so the compiler always emits LDS/STS (an earlier version kept
// generic pointers in a struct, which ended up in local memory and turned every access into a generic LD).
extern __shared__ __align__(16) float smem_f[];

template <int R>
struct MlpSmem {
  int wst;        // [kGroups][R][kMaxHidden] partial sums of the reduction groups
  int act0;       // act(l) = act0 + l*R*ld: output of hidden layer l (R x ld), kept for backward
  int dz0, dz1;   // gradient w.r.t. pre-activations / forward ping-pong (R x ld)
  int in0;        // [R][4] network input
  int out;        // [R][2] network output
  int dout;       // [R][2] gradient w.r.t. network output
  int din;        // [R][4] gradient w.r.t. network input
  int scratch;    // [R][8] per-row scalars (reward, notdone, y, ...)
  int ld;         // hid + 4: rows stay 16 B aligned, banks are skewed

  __host__ __device__ static size_t floats(int hid, int keep_layers) {
    const int ld = hid + 4;
    return (size_t)kGroups * R * kMaxHidden + (size_t)keep_layers * R * ld + 2 * (size_t)R * ld + R * 4 + R * 2 + R * 2 + R * 4 + R * 8;
  }
  __device__ __forceinline__ void carve(int base, int hid, int keep_layers) {
    ld = hid + 4;
    int p = base;
    wst = p; p += kGroups * R * kMaxHidden;
    act0 = p; p += keep_layers * R * ld;
    dz0 = p; p += R * ld;
    dz1 = p; p += R * ld;
    in0 = p; p += R * 4;
    out = p; p += R * 2;
    dout = p; p += R * 2;
    din = p; p += R * 4;
    scratch = p;
  }
  __device__ __forceinline__ int act(int l) const { return act0 + l * R * ld; }
  __device__ __forceinline__ int dz(int i) const { return i ? dz1 : dz0; }
};

__host__ inline size_t mlp_smem_bytes(int R, int hid, int keep_layers) {
  const int ld = hid + 4;
  size_t f = (size_t)kGroups * R * kMaxHidden + (size_t)keep_layers * R * ld + 2 * (size_t)R * ld + R * 4 + R * 2 + R * 2 + R * 4 + R * 8;
  return f * sizeof(float);
}

__device__ __forceinline__ float4 lds4(int off) { return *reinterpret_cast<const float4*>(smem_f + off); }

// copy the CTA's R rows of a [R][ld] shared tile to rows r0.. of a [B][hid] global array (float4, coalesced)
template <int R>
__device__ __forceinline__ void store_rows(float* __restrict__ g, int s_off, int ld, int hid, int r0, int B) {
  const int q = hid >> 2;
  for (int idx = threadIdx.x; idx < R * q; idx += kThreads) {
    const int r = idx / q, j = idx - r * q;
    if (r0 + r < B) *reinterpret_cast<float4*>(g + (int64_t)(r0 + r) * hid + 4 * j) = lds4(s_off + r * ld + 4 * j);
  }
}

// ---- first layer: h[r][c] = relu(b[c] + sum_j in0[r][j] * W[c][j]),  in <= 4 ----------------------------------
template <int R>
__device__ __forceinline__ void fwd_first(const float* __restrict__ W, const float* __restrict__ b, int in0, int in_dim, int Y, int ld, int N) {
  const int c = threadIdx.x;
  if (c < N) {
    float w[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < in_dim) w[j] = __ldg(W + c * in_dim + j);
    const float bias = __ldg(b + c);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 x = lds4(in0 + r * 4);
      float v = bias;
      v = fmaf(x.x, w[0], v); v = fmaf(x.y, w[1], v); v = fmaf(x.z, w[2], v); v = fmaf(x.w, w[3], v);
      smem_f[Y + r * ld + c] = fmaxf(v, 0.f);
    }
  }
  __syncthreads();
}

// ---- hidden-layer product: part[g][r][c] = sum_{j in range(g)} X[r][j] * M[j][c],  M row-major [J][C] in global ----
// Thread t: column group cg = t % 64 (columns 4cg..4cg+3), reduction group g = t / 64 (a contiguous quarter of j, in
// multiples of 4).  Per 4 j's: four coalesced 16 B weight loads (software-pipelined one iteration ahead), R broadcast
// LDS.128 of X, 16*R FFMA.  Partials are left in smem_f[red + (g*R + r)*kMaxHidden + c].
// kJ = weight rows per register buffer (8 for row tiles R <= 8, 4 for R = 16 where the accumulators need the registers): two buffers alternate, so 8 coalesced 16 B loads (one per row)
                         // are in flight per thread while the previous 8 rows are consumed - with 2 warps per scheduler a
                         // one-iteration (4-row) pipeline left the loop 52 % stalled on these loads (ncu r1c)

template <int R, int kJ>
__device__ __forceinline__ void consume_rows(const float4 (&w)[kJ], int X, int ld, int j, int cnt, float (&acc)[R][4]) {
#pragma unroll
  for (int jj = 0; jj < kJ; jj += 4) {
    if (jj < cnt) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 x = lds4(X + r * ld + j + jj);        // warp-broadcast
        acc[r][0] = fmaf(x.x, w[jj].x, acc[r][0]); acc[r][1] = fmaf(x.x, w[jj].y, acc[r][1]);
        acc[r][2] = fmaf(x.x, w[jj].z, acc[r][2]); acc[r][3] = fmaf(x.x, w[jj].w, acc[r][3]);
        acc[r][0] = fmaf(x.y, w[jj + 1].x, acc[r][0]); acc[r][1] = fmaf(x.y, w[jj + 1].y, acc[r][1]);
        acc[r][2] = fmaf(x.y, w[jj + 1].z, acc[r][2]); acc[r][3] = fmaf(x.y, w[jj + 1].w, acc[r][3]);
        acc[r][0] = fmaf(x.z, w[jj + 2].x, acc[r][0]); acc[r][1] = fmaf(x.z, w[jj + 2].y, acc[r][1]);
        acc[r][2] = fmaf(x.z, w[jj + 2].z, acc[r][2]); acc[r][3] = fmaf(x.z, w[jj + 2].w, acc[r][3]);
        acc[r][0] = fmaf(x.w, w[jj + 3].x, acc[r][0]); acc[r][1] = fmaf(x.w, w[jj + 3].y, acc[r][1]);
        acc[r][2] = fmaf(x.w, w[jj + 3].z, acc[r][2]); acc[r][3] = fmaf(x.w, w[jj + 3].w, acc[r][3]);
      }
    }
  }
}

template <int kJ>
__device__ __forceinline__ void load_rows(float4 (&w)[kJ], const float4* __restrict__ mp, int stride4, int cnt) {
#pragma unroll
  for (int i = 0; i < kJ; ++i)
    if (i < cnt) w[i] = __ldg(mp + i * stride4);
}

template <int R>
__device__ __forceinline__ void rows_times_matrix(const float* __restrict__ M, int X, int ld, int J, int C, int red) {
  constexpr int kJ = (R >= 16) ? 4 : 8;
  const int cg = threadIdx.x & (kColThreads - 1), g = threadIdx.x / kColThreads;
  const int c0 = cg * 4;
  const int part = ((J + 4 * kGroups - 1) / (4 * kGroups)) * 4;
  const int jlo = g * part, jhi = min(J, jlo + part);
  float acc[R][4];
#pragma unroll
  for (int r = 0; r < R; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; acc[r][3] = 0.f; }
  if (c0 < C && jlo < jhi) {
    const float4* mp = reinterpret_cast<const float4*>(M + jlo * C + c0);
    const int stride4 = C >> 2;                       // float4 per weight row
    float4 wa[kJ], wb[kJ];
    load_rows<kJ>(wa, mp, stride4, jhi - jlo);
    for (int j = jlo; j < jhi; j += 2 * kJ) {
      if (j + kJ < jhi) load_rows<kJ>(wb, mp + kJ * stride4, stride4, jhi - j - kJ);
      consume_rows<R, kJ>(wa, X, ld, j, jhi - j, acc);
      if (j + 2 * kJ < jhi) load_rows<kJ>(wa, mp + 2 * kJ * stride4, stride4, jhi - j - 2 * kJ);
      if (j + kJ < jhi) consume_rows<R, kJ>(wb, X, ld, j + kJ, jhi - j - kJ, acc);
      mp += 2 * kJ * stride4;
    }
  }
  if (c0 < C) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      *reinterpret_cast<float4*>(smem_f + red + (g * R + r) * kMaxHidden + c0) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  }
  __syncthreads();
}

// ---- hidden layer forward: Y[r][c] = relu(b[c] + sum_k X[r][k] * Wt[k][c]) ------------------------------------
template <int R>
__device__ __forceinline__ void fwd_hidden(const float* __restrict__ Wt, const float* __restrict__ b, int X, int Y, int ld, int N, int K, int red) {
  rows_times_matrix<R>(Wt, X, ld, K, N, red);
  const int c = threadIdx.x;
  if (c < N) {
    const float bias = __ldg(b + c);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float v = bias;
#pragma unroll
      for (int g = 0; g < kGroups; ++g) v += smem_f[red + (g * R + r) * kMaxHidden + c];
      smem_f[Y + r * ld + c] = fmaxf(v, 0.f);
    }
  }
  __syncthreads();
}

// ---- output layer: out[r][o] = b[o] + sum_k X[r][k] * W[o][k]  (out_dim <= 2): one warp per (r,o) pair ----------
template <int R>
__device__ __forceinline__ void fwd_out(const float* __restrict__ W, const float* __restrict__ b, int X, int ld, int K, int out_dim, int out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int p = warp; p < R * out_dim; p += kThreads / 32) {
    const int r = p / out_dim, o = p - r * out_dim;
    float v = 0.f;
    for (int k = lane; k < K; k += 32) v = fmaf(smem_f[X + r * ld + k], __ldg(W + o * K + k), v);
    v = warp_sum(v);
    if (lane == 0) smem_f[out + r * 2 + o] = v + __ldg(b + o);
  }
  __syncthreads();
}

// Whole-network forward for the CTA's R rows.  keep=true stores hidden layer l in sm.act(l) (needed by backward) and,
// when `rs` is given, also in the global row scratch; keep=false ping-pongs between sm.dz0 / sm.dz1.
template <int R>
__device__ __forceinline__ void mlp_forward(const float* __restrict__ P, const float* __restrict__ Pt, const NetShape& s, const MlpSmem<R>& sm,
                                            bool keep, const RowScratch* rs, int r0) {
  const int y0 = keep ? sm.act(0) : sm.dz0;
  fwd_first<R>(P + net_w_off(s, 0), P + net_b_off(s, 0), sm.in0, s.in, y0, sm.ld, s.hid);
  if (rs) store_rows<R>(rs->h(0), y0, sm.ld, s.hid, r0, rs->B);
  int x = y0;
  for (int l = 1; l < s.layers; ++l) {
    const int y = keep ? sm.act(l) : sm.dz(l & 1);
    fwd_hidden<R>(Pt + net_w_off(s, l), P + net_b_off(s, l), x, y, sm.ld, s.hid, s.hid, sm.wst);
    if (rs) store_rows<R>(rs->h(l), y, sm.ld, s.hid, r0, rs->B);
    x = y;
  }
  fwd_out<R>(P + net_w_off(s, s.layers), P + net_b_off(s, s.layers), x, sm.ld, s.hid, s.out, sm.out);
}

// ---- backward (gradients w.r.t. pre-activations only; see wgrad_kernel for the parameters) -----------------------
// Output layer: dz_L[r][k] = relu'(h_L[r][k]) * sum_o dout[r][o] * Wout[o][k]
template <int R>
__device__ __forceinline__ void bwd_out(const float* __restrict__ W, int H, int ld, int K, int out_dim, int dout, int dz) {
  const int k = threadIdx.x;
  if (k < K) {
    const float w0 = __ldg(W + k), w1 = out_dim > 1 ? __ldg(W + K + k) : 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float d = fmaf(smem_f[dout + r * 2], w0, smem_f[dout + r * 2 + 1] * w1);
      smem_f[dz + r * ld + k] = smem_f[H + r * ld + k] > 0.f ? d : 0.f;
    }
  }
  __syncthreads();
}

// Input gradient of a hidden layer: dzp[r][k] = relu'(Hprev[r][k]) * sum_n dz[r][n] * W[n][k]   (torch layout W [N][K])
template <int R>
__device__ __forceinline__ void bwd_input(const float* __restrict__ W, int dz, int Hprev, int dzp, int ld, int N, int K, int red) {
  rows_times_matrix<R>(W, dz, ld, N, K, red);
  const int k = threadIdx.x;
  if (k < K) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float v = 0.f;
#pragma unroll
      for (int g = 0; g < kGroups; ++g) v += smem_f[red + (g * R + r) * kMaxHidden + k];
      smem_f[dzp + r * ld + k] = smem_f[Hprev + r * ld + k] > 0.f ? v : 0.f;
    }
  }
  __syncthreads();
}

// First layer input gradient (only the actor loss needs it): din[r][j] = sum_c dz[r][c] * W1[c][j]
template <int R>
__device__ __forceinline__ void bwd_first_input(const float* __restrict__ W, int in_dim, int dz, int ld, int N, int din) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int p = warp; p < R * in_dim; p += kThreads / 32) {
    const int r = p / in_dim, j = p - r * in_dim;
    float v = 0.f;
    for (int cc = lane; cc < N; cc += 32) v = fmaf(smem_f[dz + r * ld + cc], __ldg(W + cc * in_dim + j), v);
    v = warp_sum(v);
    if (lane == 0) smem_f[din + r * 4 + j] = v;
  }
  __syncthreads();
}

// Whole-network backward from sm.dout ([R][2]).  Needs the activations kept by mlp_forward(keep=true).
// rs != nullptr: every layer's dz (and dout, in0) goes to the row scratch for the weight-gradient pass.
template <int R>
__device__ __forceinline__ void mlp_backward(const float* __restrict__ P, const NetShape& s, const MlpSmem<R>& sm, const RowScratch* rs,
                                             int r0, bool want_din) {
  const int L = s.layers;
  int cur = 0;
  if (rs && threadIdx.x < R && r0 + threadIdx.x < rs->B) {
    const int r = threadIdx.x;
    *reinterpret_cast<float4*>(rs->in0() + (int64_t)(r0 + r) * 4) = lds4(sm.in0 + r * 4);
    *reinterpret_cast<float2*>(rs->dout() + (int64_t)(r0 + r) * 2) = *reinterpret_cast<const float2*>(smem_f + sm.dout + r * 2);
  }
  bwd_out<R>(P + net_w_off(s, L), sm.act(L - 1), sm.ld, s.hid, s.out, sm.dout, sm.dz(cur));
  if (rs) store_rows<R>(rs->dz(L - 1), sm.dz(cur), sm.ld, s.hid, r0, rs->B);
  for (int l = L - 1; l >= 1; --l) {
    bwd_input<R>(P + net_w_off(s, l), sm.dz(cur), sm.act(l - 1), sm.dz(cur ^ 1), sm.ld, s.hid, s.hid, sm.wst);
    cur ^= 1;
    if (rs) store_rows<R>(rs->dz(l - 1), sm.dz(cur), sm.ld, s.hid, r0, rs->B);
  }
  if (want_din) bwd_first_input<R>(P + net_w_off(s, 0), s.in, sm.dz(cur), sm.ld, s.hid, sm.din);
}

}  // namespace rtd3
