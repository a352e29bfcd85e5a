// Small-batch residual-TD3 steps on thread-block clusters (fp32, the parity path): robot.py:312-366 (train_critic) and
// robot.py:369-398 (train_actor) up to the optimiser steps, for batches that cannot fill the GPU row by row.
//
// The row-tile kernels of rtd3_td3.cu give every CTA 2-16 batch rows and let it stream the weights of all five networks from L2:
// at B = 256 that is 128 CTAs x 1.8 MB = 240 MB of L2 reads for 3.2 MB of parameters, and the step runs at the L2 bandwidth.  Here a
// CLUSTER of CS CTAs owns R batch rows and splits the COLUMNS of every layer: CTA c computes output columns [c*Wc, (c+1)*Wc) of each
// layer for all R rows, so it needs 1/CS of every weight matrix (a [H][Wc] column slice, staged by cp.async through a ring that runs
// ahead of the arithmetic - the weights do not depend on the activations), and pushes its slice of the layer output into the shared
// memory of all CTAs of the cluster so that each of them holds the full-width input of the next layer.  The pushes are st.async
// stores that complete transaction bytes on an mbarrier of the RECEIVING CTA: a CTA goes on as soon as ITS input is complete - no
// cluster-wide barrier per layer (measured: barrier.cluster after a push costs 1 300-2 900 cycles, the products 1 500).  Two
// mbarriers alternate by stage; consecutive stages write different tiles, and a CTA can only be one stage ahead of the slowest CTA
// of its cluster (it needs that CTA's output to go on), which is what makes the reuse of tiles and barriers two stages later safe.
// Network heads (1-2 outputs) are evaluated redundantly by every CTA.  The per-row layer inputs / pre-activation gradients go to the
// same row scratch as before; wgrad_kernel (+ fused Adam) is unchanged.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "rtd3_common.cuh"
#include "rtd3_mlp.cuh"
#include "rtd3_td3.cuh"

namespace rtd3 {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// 16 / 8 bytes into the shared memory of a CTA of the cluster; the bytes are counted on that CTA's mbarrier when they have landed
__device__ __forceinline__ void st_async4(uint32_t addr, const float4& v, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void st_async2(uint32_t addr, float a, float b, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(addr), "f"(a), "f"(b), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITC_%=:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONEC_%=;\n"
      "bra WAITC_%=;\n"
      "DONEC_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// Shared-memory plan of one CTA (offsets in floats into smem_f), its place in the cluster and the running stage state.
struct Cl {
  int ring;          // [ns][Wc / 4][4 H + 4] weight slices: groups of 4 rows, row n = the H weights that feed own output column n
  int red;           // [KG][R][Wc] partial sums of the reduction groups
  int small;         // per network: W0 slice [Wc][4] | biases [L][Wc] | head [2][H] | head bias [4]
  int act0;          // kept activations, nkeep full-width [R][ld] tiles
  int dz0, dz1;      // two more full-width tiles (forward of the target networks, backward)
  int in0, out, dout, din, S, dinp;   // [R][4], [R][2], [R][2], [R][4], [R][8], [CS][R][2]
  int bar;           // mbarriers (uint64): [0,1] stages, [2..4] weight ring slots, [5] small parameters
  int ld, Wc, H, L, ns, tile_floats, small_stride;
  int ncg_sh;        // log2 of the column groups a reduction group spans (>= Wc / 4)
  int kgn, part;     // reduction groups and their length
  int rank, c_lo, ncols;
  int stg;           // stages completed so far (the same number in every thread of the cluster)
  int dsel;          // which dz tile the next un-kept output goes to
  long long* prof;   // development: stage stamps of thread 0 of CTA 0 ((code << 48) | clock64), nullptr = off
  int pn;
};
__device__ __forceinline__ void stamp(Cl& c, int code) {
  if (c.prof && threadIdx.x == 0 && c.pn < 254) c.prof[1 + c.pn++] = ((long long)code << 48) | (clock64() & 0xffffffffffffll);
}

struct ClPlan {      // host side of the same plan
  int ns;
  size_t bytes;
};

__host__ __device__ inline int cl_wc(int H, int CS) { return ((H + 4 * CS - 1) / (4 * CS)) * 4; }
__host__ __device__ inline int cl_ncg_sh(int Wc) {
  int sh = 0;
  while ((4 << sh) < Wc) ++sh;
  return sh;
}
constexpr size_t kClSmemLimit = 226 * 1024;

static ClPlan make_plan(int R, int CS, int H, int L, int nets, int nkeep) {
  const int Wc = cl_wc(H, CS), ld = H + 4;
  const int kgn = kThreads / ((1 << cl_ncg_sh(Wc)) * (R / 4));
  const size_t tile = (size_t)(Wc / 4) * (4 * H + 4);
  const size_t small_stride = (size_t)(4 + L) * Wc + 2 * H + 4;
  const size_t fixed = (size_t)kgn * R * Wc + nets * small_stride + (size_t)(nkeep + 2) * R * ld + R * (4 + 2 + 2 + 4 + 8) + (size_t)CS * R * 2 + 16;
  ClPlan p;
  p.ns = 3;
  p.bytes = (fixed + p.ns * tile) * sizeof(float);
  if (p.bytes > kClSmemLimit) {
    p.ns = 2;
    p.bytes = (fixed + p.ns * tile) * sizeof(float);
  }
  return p;
}

__device__ __forceinline__ void cl_carve(Cl& c, int R, int CS, int H, int L, int ns, int nets, int nkeep) {
  c.H = H; c.L = L; c.ns = ns;
  c.Wc = cl_wc(H, CS);
  c.ld = H + 4;
  c.tile_floats = (c.Wc / 4) * (4 * H + 4);
  c.small_stride = (4 + L) * c.Wc + 2 * H + 4;
  c.ncg_sh = cl_ncg_sh(c.Wc);
  c.kgn = kThreads / ((1 << c.ncg_sh) * (R / 4));
  c.part = ((H + 4 * c.kgn - 1) / (4 * c.kgn)) * 4;
  int p = 0;
  c.ring = p; p += ns * c.tile_floats;
  c.red = p; p += c.kgn * R * c.Wc;
  c.small = p; p += nets * c.small_stride;
  c.act0 = p; p += nkeep * R * c.ld;
  c.dz0 = p; p += R * c.ld;
  c.dz1 = p; p += R * c.ld;
  c.in0 = p; p += R * 4;
  c.out = p; p += R * 2;
  c.dout = p; p += R * 2;
  c.din = p; p += R * 4;
  c.S = p; p += R * 8;
  c.dinp = p; p += CS * R * 2;
  c.bar = p;
  c.rank = (int)cluster_ctarank();
  c.c_lo = c.rank * c.Wc;
  c.ncols = max(0, min(c.Wc, H - c.c_lo));
  c.stg = 0; c.dsel = 0;
  c.prof = nullptr; c.pn = 0;
}

// ---- weight ring ------------------------------------------------------------------------------------------------------------------
// tile = the rows of one H x H matrix that produce the own output columns: forward W_l[n][:] (torch layout, n = own column),
// backward Wt_l[k][:] (the transposed copy, k = own column) - Wc contiguous rows of H floats, fetched in groups of 4 rows (one bulk
// copy of 16 H bytes per group: per-row copies made the issue, ~65 cycles per copy, the longest part of a stage) into groups of
// pitch 4 H + 4 floats: the product's lane cl reads the rows of group cl, and the skew keeps the lanes on different banks.
// Warp 0 issues, the copies complete on the slot's mbarrier.
__device__ __forceinline__ uint64_t* cl_bar(const Cl& c, int i) { return reinterpret_cast<uint64_t*>(smem_f + c.bar) + i; }
__device__ __forceinline__ void tile_issue(const Cl& c, const float* __restrict__ base, int slot) {
  if (base && threadIdx.x < 32 && c.ncols > 0) {
    uint64_t* bar = cl_bar(c, 2 + slot);
    if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, (uint32_t)(c.ncols * c.H * 4));
    __syncwarp();
    const int g = threadIdx.x;
    if (g < (c.ncols >> 2))
      bulk_g2s(smem_f + c.ring + slot * c.tile_floats + g * (4 * c.H + 4), base + (int64_t)(c.c_lo + 4 * g) * c.H, (uint32_t)(16 * c.H), bar);
  }
}
// make tile i resident; the slot of tile i-1 (all threads have left its product) is refilled with tile i + ns - 1
__device__ __forceinline__ int tile_acquire(const Cl& c, const float* __restrict__ next_base, int i) {
  __syncthreads();          // the previous epilogue has left the partial sums (the stage waits in between are not CTA barriers)
  tile_issue(c, next_base, (i + c.ns - 1) % c.ns);
  if (c.ncols > 0) mbar_wait(cl_bar(c, 2 + i % c.ns), (uint32_t)((i / c.ns) & 1));
  return c.ring + (i % c.ns) * c.tile_floats;
}

// small parameters of one network (own column slice) by bulk copies on mbarrier 5: W0 slice [ncols][in] | biases [L][Wc] | head
// [2][H] | head bias [4].  One lane per copy; small_bytes() is what the arming thread expects.
__device__ __forceinline__ uint32_t small_bytes(const Cl& c, const NetShape& s) {
  return (uint32_t)((c.ncols * s.in + s.layers * c.ncols + s.out * c.H + 4) * 4);
}
__device__ __forceinline__ void small_issue(const Cl& c, int slot, const float* __restrict__ P, const NetShape& s, int item) {
  float* sp = smem_f + c.small + slot * c.small_stride;
  uint64_t* bar = cl_bar(c, 5);
  if (item == 0) {
    if (c.ncols > 0) bulk_g2s(sp, P + net_w_off(s, 0) + c.c_lo * s.in, (uint32_t)(c.ncols * s.in * 4), bar);
  } else if (item <= s.layers) {
    const int l = item - 1;
    if (c.ncols > 0) bulk_g2s(sp + (4 + l) * c.Wc, P + net_b_off(s, l) + c.c_lo, (uint32_t)(c.ncols * 4), bar);
  } else if (item == kMaxLayers + 1) {
    bulk_g2s(sp + (4 + s.layers) * c.Wc, P + net_w_off(s, s.layers), (uint32_t)(s.out * c.H * 4), bar);
  } else if (item == kMaxLayers + 2) {
    bulk_g2s(sp + (4 + s.layers) * c.Wc + 2 * c.H, P + net_b_off(s, s.layers), 16u, bar);
  }
}
// the second head row of a one-output network is read (times a zero gradient) by the backward pass: it must be finite
__device__ __forceinline__ void small_zero(const Cl& c, int slot, const NetShape& s) {
  if (s.out < 2) {
    float* head = smem_f + c.small + slot * c.small_stride + (4 + s.layers) * c.Wc;
    for (int i = threadIdx.x; i < c.H; i += kThreads) head[c.H + i] = 0.f;
  }
}
__device__ __forceinline__ int sp_w0(const Cl& c, int slot) { return c.small + slot * c.small_stride; }
__device__ __forceinline__ int sp_bias(const Cl& c, int slot, int l) { return c.small + slot * c.small_stride + (4 + l) * c.Wc; }
__device__ __forceinline__ int sp_head(const Cl& c, int slot) { return c.small + slot * c.small_stride + (4 + c.L) * c.Wc; }

// the dz tile the next un-kept output goes to (they alternate: consecutive stages never write the same tile)
__device__ __forceinline__ int next_dz(Cl& c) {
  const int r = c.dsel ? c.dz1 : c.dz0;
  c.dsel ^= 1;
  return r;
}

// ---- stages ------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t stage_bar(const Cl& c) { return smem_u32(smem_f + c.bar) + 8u * (uint32_t)(c.stg & 1); }
// four own values (columns c_lo + cc .. +3 of row r) into tile Y of every CTA of the cluster
template <int CS>
__device__ __forceinline__ void push4(const Cl& c, int off, const float4& v) {
  const uint32_t own = smem_u32(smem_f + off), bar = stage_bar(c);
#pragma unroll
  for (int k = 0; k < CS; ++k) st_async4(map_to_rank(own, (uint32_t)k), v, map_to_rank(bar, (uint32_t)k));
}
// wait until the `bytes` all CTAs push to this one in the current stage have landed
__device__ __forceinline__ void stage_wait(Cl& c, uint32_t bytes) {
  const uint32_t bar = stage_bar(c);
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  mbar_wait_cluster(bar, (uint32_t)((c.stg >> 1) & 1));
  c.stg += 1;
}

// ---- first layer, own columns: h[r][c] = relu(b[c] + sum_j in0[r][j] * W0[c][j]) ---------------------------------------------------
template <int R, int CS>
__device__ __forceinline__ void cl_first(Cl& c, int slot, int in_dim, int Y, float* __restrict__ gh /*nullable [B][H]*/, int r0, int B) {
  const int q = c.ncols >> 2;
  const int w0 = sp_w0(c, slot), bb = sp_bias(c, slot, 0);
  for (int idx = threadIdx.x; idx < R * q; idx += kThreads) {
    const int r = idx / q, cc = (idx - r * q) * 4;
    const float4 x = lds4(c.in0 + r * 4);
    const float4 b = lds4(bb + cc);
    float v[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (in_dim == 4) {
        const float4 w = lds4(w0 + (cc + e) * 4);
        v[e] = fmaf(x.x, w.x, v[e]); v[e] = fmaf(x.y, w.y, v[e]); v[e] = fmaf(x.z, w.z, v[e]); v[e] = fmaf(x.w, w.w, v[e]);
      } else {
        const float2 w = *reinterpret_cast<const float2*>(smem_f + w0 + (cc + e) * 2);
        v[e] = fmaf(x.x, w.x, v[e]); v[e] = fmaf(x.y, w.y, v[e]);
      }
      v[e] = fmaxf(v[e], 0.f);
    }
    const float4 o = make_float4(v[0], v[1], v[2], v[3]);
    push4<CS>(c, Y + r * c.ld + c.c_lo + cc, o);
    if (gh && r0 + r < B) *reinterpret_cast<float4*>(gh + (int64_t)(r0 + r) * c.H + c.c_lo + cc) = o;
  }
  stamp(c, 10);
  stage_wait(c, (uint32_t)(R * c.H * 4));
  stamp(c, 11);
}

// ---- hidden product, own columns: red[g][r][c] = sum_{j in group g} X[r][j] * T[c][j] ------------------------------------------------
// thread = (reduction group, row group rg, column group cl); it owns rows rg, rg + RG, rg + 2 RG, rg + 3 RG (neighbouring row groups
// read neighbouring rows: with ld = H + 4 their float4 fall into different banks) and the 4 columns of tile group cl
template <int R>
__device__ __forceinline__ void cl_product(const Cl& c, int T, int X) {
  constexpr int RG = R / 4;
  const int t = threadIdx.x;
  const int cl = t & ((1 << c.ncg_sh) - 1);
  const int rg = (t >> c.ncg_sh) % RG;
  const int kg = (t >> c.ncg_sh) / RG;
  const int jlo = kg * c.part, jhi = min(c.H, jlo + c.part);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
  if (4 * cl < c.ncols) {
    const int tw = T + cl * (4 * c.H + 4), xr = X + rg * c.ld;
#pragma unroll 2
    for (int j = jlo; j < jhi; j += 4) {
      float4 w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) w[e] = lds4(tw + e * c.H + j);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 x = lds4(xr + i * RG * c.ld + j);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[i][e] = fmaf(x.x, w[e].x, acc[i][e]); acc[i][e] = fmaf(x.y, w[e].y, acc[i][e]);
          acc[i][e] = fmaf(x.z, w[e].z, acc[i][e]); acc[i][e] = fmaf(x.w, w[e].w, acc[i][e]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(smem_f + c.red + (kg * R + rg + i * RG) * c.Wc + 4 * cl) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
  __syncthreads();
}

// sum of the partials of the four values (r, cc..cc+3)
template <int R>
__device__ __forceinline__ float4 cl_reduce4(const Cl& c, int r, int cc) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int g = 0; g < c.kgn; ++g) {
    const float4 p = lds4(c.red + (g * R + r) * c.Wc + cc);
    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
  }
  return v;
}

// ---- hidden layer forward: Y[r][c] = relu(b[c] + sum_k X[r][k] * Wt[k][c]), own columns, pushed to the whole cluster ----------------
template <int R, int CS>
__device__ __forceinline__ void cl_fwd_hidden(Cl& c, int T, int slot, int l, int X, int Y, float* __restrict__ gh, int r0, int B) {
  stamp(c, 20);
  cl_product<R>(c, T, X);
  stamp(c, 21);
  const int q = c.ncols >> 2, bb = sp_bias(c, slot, l);
  for (int idx = threadIdx.x; idx < R * q; idx += kThreads) {
    const int r = idx / q, cc = (idx - r * q) * 4;
    const float4 p = cl_reduce4<R>(c, r, cc);
    const float4 b = lds4(bb + cc);
    const float4 o = make_float4(fmaxf(b.x + p.x, 0.f), fmaxf(b.y + p.y, 0.f), fmaxf(b.z + p.z, 0.f), fmaxf(b.w + p.w, 0.f));
    push4<CS>(c, Y + r * c.ld + c.c_lo + cc, o);
    if (gh && r0 + r < B) *reinterpret_cast<float4*>(gh + (int64_t)(r0 + r) * c.H + c.c_lo + cc) = o;
  }
  stamp(c, 22);
  stage_wait(c, (uint32_t)(R * c.H * 4));
  stamp(c, 23);
}

// ---- head, evaluated by every CTA for all R rows: out[r][o] = b[o] + sum_k X[r][k] * Wout[o][k]; 16 threads per dot product -----------
template <int R>
__device__ __forceinline__ void cl_head(Cl& c, int slot, int X, int out_dim) {
  const int grp = threadIdx.x >> 4, sub = threadIdx.x & 15;
  const int hd = sp_head(c, slot);
  const int np = R * out_dim;
  for (int p0 = 0; p0 < np; p0 += kThreads / 16) {                  // the trip count is the same for all threads (full-warp shuffles)
    const int p = min(p0 + grp, np - 1);
    const int r = p / out_dim, o = p - r * out_dim;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (int k = sub * 4; k < c.H; k += 64) {
      const float4 x = lds4(X + r * c.ld + k), w = lds4(hd + o * c.H + k);
      a0 = fmaf(x.x, w.x, a0); a1 = fmaf(x.y, w.y, a1); a2 = fmaf(x.z, w.z, a2); a3 = fmaf(x.w, w.w, a3);
    }
    float v = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int s = 8; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (sub == 0 && p0 + grp < np) smem_f[c.out + r * 2 + o] = v + smem_f[hd + 2 * c.H + o];
  }
  __syncthreads();
  stamp(c, 30);
}

template <int R, int CS>
__device__ __forceinline__ void cl_forward(Cl& c, int slot, const NetShape& s, bool keep, int keep0, const RowScratch* rs, int r0,
                                           const float* const* tiles, int& ti) {
  const int B = rs ? rs->B : 0;
  int x = keep ? c.act0 + keep0 * R * c.ld : next_dz(c);
  cl_first<R, CS>(c, slot, s.in, x, rs ? rs->h(0) : nullptr, r0, B);
  for (int l = 1; l < s.layers; ++l) {
    const int y = keep ? c.act0 + (keep0 + l) * R * c.ld : next_dz(c);
    const int T = tile_acquire(c, tiles[ti + c.ns - 1], ti);
    ++ti;
    cl_fwd_hidden<R, CS>(c, T, slot, l, x, y, rs ? rs->h(l) : nullptr, r0, B);
    x = y;
  }
  cl_head<R>(c, slot, x, s.out);
}

// ---- backward (pre-activation gradients; the parameter gradients are wgrad_kernel's) ---------------------------------------------
// dz_{L-1}[r][k] = relu'(h_{L-1}[r][k]) * sum_o dout[r][o] * Wout[o][k], own columns -> all CTAs (the next product reduces over it)
// dz_{l-1}[r][k] = relu'(h_{l-1}[r][k]) * sum_n dz_l[r][n] * W_l[n][k]
// want_din: din[r][j] = sum_c dz_0[r][c] * W0[c][j] (partials over own columns, summed over the cluster in rank order)
template <int R, int CS>
__device__ __forceinline__ void cl_backward(Cl& c, int slot, const NetShape& s, int keep0, const RowScratch* rs, int r0, bool want_din,
                                            const float* const* tiles, int& ti) {
  const int L = s.layers, B = rs ? rs->B : 0;
  const int t = threadIdx.x;
  if (rs && c.rank == 0 && t < R && r0 + t < B) {
    *reinterpret_cast<float4*>(rs->in0() + (int64_t)(r0 + t) * 4) = lds4(c.in0 + t * 4);
    *reinterpret_cast<float2*>(rs->dout() + (int64_t)(r0 + t) * 2) = *reinterpret_cast<const float2*>(smem_f + c.dout + t * 2);
  }
  const int q = c.ncols >> 2;
  int cur = next_dz(c);
  {
    const bool push = L > 1;                          // a product reduces over it
    const int hd = sp_head(c, slot), Hl = c.act0 + (keep0 + L - 1) * R * c.ld;
    float* g = rs ? rs->dz(L - 1) : nullptr;
    for (int idx = t; idx < R * q; idx += kThreads) {
      const int r = idx / q, cc = (idx - r * q) * 4, k = c.c_lo + cc;
      const float d0 = smem_f[c.dout + r * 2], d1 = smem_f[c.dout + r * 2 + 1];
      const float4 wa = lds4(hd + k), wb = lds4(hd + c.H + k), h = lds4(Hl + r * c.ld + k);
      float4 o;
      o.x = h.x > 0.f ? fmaf(d0, wa.x, d1 * wb.x) : 0.f;
      o.y = h.y > 0.f ? fmaf(d0, wa.y, d1 * wb.y) : 0.f;
      o.z = h.z > 0.f ? fmaf(d0, wa.z, d1 * wb.z) : 0.f;
      o.w = h.w > 0.f ? fmaf(d0, wa.w, d1 * wb.w) : 0.f;
      if (push) push4<CS>(c, cur + r * c.ld + k, o);
      else *reinterpret_cast<float4*>(smem_f + cur + r * c.ld + k) = o;
      if (g && r0 + r < B) *reinterpret_cast<float4*>(g + (int64_t)(r0 + r) * c.H + k) = o;
    }
    stamp(c, 40);
    if (push) stage_wait(c, (uint32_t)(R * c.H * 4)); else __syncthreads();
    stamp(c, 41);
  }
  for (int l = L - 1; l >= 1; --l) {
    const int T = tile_acquire(c, tiles[ti + c.ns - 1], ti);
    ++ti;
    stamp(c, 50);
    cl_product<R>(c, T, cur);
    stamp(c, 51);
    const int nxt = next_dz(c);
    const int Hp = c.act0 + (keep0 + l - 1) * R * c.ld;
    float* g = rs ? rs->dz(l - 1) : nullptr;
    const bool push = l - 1 >= 1;                    // another product follows
    for (int idx = t; idx < R * q; idx += kThreads) {
      const int r = idx / q, cc = (idx - r * q) * 4, k = c.c_lo + cc;
      const float4 p = cl_reduce4<R>(c, r, cc);
      const float4 h = lds4(Hp + r * c.ld + k);
      const float4 o = make_float4(h.x > 0.f ? p.x : 0.f, h.y > 0.f ? p.y : 0.f, h.z > 0.f ? p.z : 0.f, h.w > 0.f ? p.w : 0.f);
      if (push) push4<CS>(c, nxt + r * c.ld + k, o);
      else if (want_din) *reinterpret_cast<float4*>(smem_f + nxt + r * c.ld + k) = o;
      if (g && r0 + r < B) *reinterpret_cast<float4*>(g + (int64_t)(r0 + r) * c.H + k) = o;
    }
    stamp(c, 52);
    if (push) stage_wait(c, (uint32_t)(R * c.H * 4)); else __syncthreads();
    stamp(c, 53);
    cur = nxt;
  }
  if (want_din) {
    // partial over own columns: one warp per (r, j); the partials of the CS CTAs meet in dinp[rank][r][j] of every CTA
    const int warp = t >> 5, lane = t & 31;
    const int w0 = sp_w0(c, slot);
    for (int p = warp; p < R * 2; p += kThreads / 32) {          // only the action components j = 2, 3 are used (robot.py:386-391)
      const int r = p >> 1, j = 2 + (p & 1);
      float v = 0.f;
      for (int cc = lane; cc < c.ncols; cc += 32) v = fmaf(smem_f[cur + r * c.ld + c.c_lo + cc], smem_f[w0 + cc * 4 + j], v);
      v = warp_sum(v);
      if (lane == 0) smem_f[c.din + r * 4 + j] = v;
    }
    __syncthreads();
    if (t < R) {
      const uint32_t own = smem_u32(smem_f + c.dinp + (c.rank * R + t) * 2), bar = stage_bar(c);
      const float a = smem_f[c.din + t * 4 + 2], b = smem_f[c.din + t * 4 + 3];
#pragma unroll
      for (int k = 0; k < CS; ++k) st_async2(map_to_rank(own, (uint32_t)k), a, b, map_to_rank(bar, (uint32_t)k));
    }
    stage_wait(c, (uint32_t)(CS * R * 8));
    if (t < R) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int k = 0; k < CS; ++k) { a += smem_f[c.dinp + (k * R + t) * 2]; b += smem_f[c.dinp + (k * R + t) * 2 + 1]; }
      smem_f[c.din + t * 4 + 2] = a;
      smem_f[c.din + t * 4 + 3] = b;
    }
    __syncthreads();
    stamp(c, 60);
  }
}

constexpr int kMaxTiles = 7 * (kMaxLayers - 1) + 4;      // critic step: 7 (L-1) tiles, + ns slack of nulls

// common start: mbarriers, the small parameters of `nets` networks (warp 1 issues), the first weight tiles (warp 0)
struct SmallNet { const float* P; NetShape s; };
template <int NETS>
__device__ __forceinline__ void cl_begin(Cl& c, const SmallNet (&nets)[NETS], const float* const* tiles) {
  if (threadIdx.x == 0) {
    mbar_init(cl_bar(c, 0), 1);
    mbar_init(cl_bar(c, 1), 1);
    for (int i = 0; i < 3; ++i) mbar_init(cl_bar(c, 2 + i), 1);
    mbar_init(cl_bar(c, 5), 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 32) {
    uint32_t bytes = 0;
#pragma unroll
    for (int n = 0; n < NETS; ++n) bytes += small_bytes(c, nets[n].s);
    mbar_arrive_expect_tx(cl_bar(c, 5), bytes);
  }
  if (threadIdx.x >= 32 && threadIdx.x < 64) {
    __syncwarp();
    const int lane = threadIdx.x - 32;
#pragma unroll
    for (int n = 0; n < NETS; ++n)
      if (lane < kMaxLayers + 3) small_issue(c, n, nets[n].P, nets[n].s, lane);
  }
#pragma unroll
  for (int n = 0; n < NETS; ++n) small_zero(c, n, nets[n].s);
  for (int i = 0; i < c.ns - 1; ++i) tile_issue(c, tiles[i], i);
}
__device__ __forceinline__ void cl_started(Cl& c) {
  // the small parameters have landed; every CTA of the cluster runs and has its mbarriers set up
  mbar_wait(cl_bar(c, 5), 0);
  stamp(c, 1);
  cluster_sync();
  stamp(c, 2);
}

// ---- critic phase (robot.py:312-366 up to the optimiser steps); see td3_critic_kernel for the arithmetic ---------------------------
template <int R, int CS>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(kThreads, 1)
td3_critic_cluster_kernel(Arena ar, const float* __restrict__ params, const float* __restrict__ params_t, float* __restrict__ scratch, ReplayView rp,
                          const int32_t* __restrict__ idx, const float* __restrict__ noise, int B, Td3Hyper hp, float* __restrict__ loss,
                          float* __restrict__ q_out, float* __restrict__ y_out, int32_t* __restrict__ steps, double* __restrict__ beta_pows, int ns,
                          long long* prof) {
  Cl c;
  const int L = ar.critic.layers;
  cl_carve(c, R, CS, ar.critic.hid, L, ns, 5, L);
  if (blockIdx.x == 0) c.prof = prof;
  stamp(c, 0);
  const int r0 = (int)cluster_id_x() * R;
  const int t = threadIdx.x;
  if (blockIdx.x == 0 && t == 0) advance_adam_clock(steps, beta_pows, 1);

  // weight tiles in the order of use: passes 0-2 forward, passes 3-4 forward then backward
  __shared__ const float* tiles[kMaxTiles];
  if (t < kMaxTiles) {
    const int n = L - 1;
    const float* p = nullptr;
    if (n > 0 && t < 7 * n) {
      int pass, l;
      bool bwd = false;
      if (t < 3 * n) { pass = t / n; l = 1 + t % n; }
      else {
        const int j = t - 3 * n;
        pass = 3 + j / (2 * n);
        const int jj = j % (2 * n);
        bwd = jj >= n;
        l = bwd ? (L - 1 - (jj - n)) : 1 + jj;
      }
      const int net = pass == 0 ? 3 : (pass == 1 ? 4 : (pass == 2 ? 5 : pass - 2));
      p = (bwd ? params_t : params) + ar.off(net) + net_w_off(pass == 0 ? ar.actor : ar.critic, l);
    }
    tiles[t] = p;
  }
  const SmallNet nets[5] = {{params + ar.off(3), ar.actor}, {params + ar.off(4), ar.critic}, {params + ar.off(5), ar.critic},
                            {params + ar.off(1), ar.critic}, {params + ar.off(2), ar.critic}};
  cl_begin<5>(c, nets, tiles);
  float* S = smem_f + c.S;       // [R][8]: 0 s.x 1 s.y 2 a.x 3 a.y 4 reward 5 notdone 6 y 7 valid
  float* in0 = smem_f + c.in0;
  float* out = smem_f + c.out;
  float* dout = smem_f + c.dout;
  if (t < R) {
    const int row = r0 + t;
    const bool valid = row < B;
    const int j = valid ? idx[row] : 0;
    const float2 s = rp.s[j], a = rp.a[j], s2 = rp.s2[j];
    S[t * 8 + 0] = s.x; S[t * 8 + 1] = s.y; S[t * 8 + 2] = a.x; S[t * 8 + 3] = a.y;
    S[t * 8 + 4] = rp.r[j]; S[t * 8 + 5] = rp.notdone[j]; S[t * 8 + 7] = valid ? 1.f : 0.f;
    in0[t * 4 + 0] = s2.x; in0[t * 4 + 1] = s2.y; in0[t * 4 + 2] = 0.f; in0[t * 4 + 3] = 0.f;
  }
  cl_started(c);

  int ti = 0;
  for (int pass = 0; pass < 5; ++pass) {
    const NetShape shape = pass == 0 ? ar.actor : ar.critic;
    const bool train = pass >= 3;
    RowScratch rs{scratch + (train ? (pass - 3) : 0) * RowScratch::floats(B, ar.critic.hid, L), B, ar.critic.hid, L};
    cl_forward<R, CS>(c, pass, shape, train, 0, train ? &rs : nullptr, r0, tiles, ti);
    if (t < R) {
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
        if (y_out && c.rank == 0 && r0 + t < B) y_out[r0 + t] = y;
        in0[t * 4 + 0] = S[t * 8 + 0]; in0[t * 4 + 1] = S[t * 8 + 1];
        in0[t * 4 + 2] = S[t * 8 + 2]; in0[t * 4 + 3] = S[t * 8 + 3];
      } else {                                           // MSE loss and its gradient (robot.py:348-353)
        const int cr = pass - 3;
        const float valid = S[t * 8 + 7];
        const float q = out[t * 2];
        const float diff = (q - S[t * 8 + 6]) * valid;
        dout[t * 2] = 2.0f * diff / (float)B;
        dout[t * 2 + 1] = 0.f;
        if (q_out && c.rank == 0 && valid != 0.f) q_out[cr * B + r0 + t] = q;
        float l = diff * diff / (float)B;                // this row's share of the mean
#pragma unroll
        for (int o = R / 2; o > 0; o >>= 1) l += __shfl_xor_sync((R >= 32) ? 0xffffffffu : ((1u << R) - 1u), l, o);
        if (t == 0 && c.rank == 0) atomicAdd(loss + cr, l);
      }
    }
    __syncthreads();
    if (train) cl_backward<R, CS>(c, pass, shape, 0, &rs, r0, false, tiles, ti);
  }
  stamp(c, 98);
  cluster_sync();        // no CTA leaves while stores of its peers may still be on their way to it
  stamp(c, 99);
  if (c.prof && t == 0) c.prof[0] = c.pn;
}

// ---- actor phase (robot.py:369-398 up to the optimiser step) ----------------------------------------------------------------------
template <int R, int CS>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(kThreads, 1)
td3_actor_cluster_kernel(Arena ar, const float* __restrict__ params, const float* __restrict__ params_t, float* __restrict__ scratch, ReplayView rp,
                         const int32_t* __restrict__ idx, int B, float* __restrict__ loss, int32_t* __restrict__ steps, double* __restrict__ beta_pows,
                         int ns, long long* prof) {
  Cl c;
  const int L = ar.critic.layers;
  cl_carve(c, R, CS, ar.critic.hid, L, ns, 2, 2 * L);
  if (blockIdx.x == 0) c.prof = prof;
  stamp(c, 0);
  const int r0 = (int)cluster_id_x() * R;
  const int t = threadIdx.x;
  if (blockIdx.x == 0 && t == 0) advance_adam_clock(steps, beta_pows, 0);
  __shared__ const float* tiles[kMaxTiles];
  if (t < kMaxTiles) {
    const int n = L - 1;
    const float* p = nullptr;
    if (n > 0 && t < 4 * n) {
      const int seg = t / n, j = t % n;          // 0 actor forward, 1 critic forward, 2 critic backward, 3 actor backward
      const int net = (seg == 0 || seg == 3) ? 0 : 1;
      const int l = seg < 2 ? 1 + j : L - 1 - j;
      p = (seg < 2 ? params : params_t) + ar.off(net) + net_w_off(net == 0 ? ar.actor : ar.critic, l);
    }
    tiles[t] = p;
  }
  const SmallNet nets[2] = {{params + ar.off(0), ar.actor}, {params + ar.off(1), ar.critic}};
  cl_begin<2>(c, nets, tiles);
  float* S = smem_f + c.S;
  float* in0 = smem_f + c.in0;
  RowScratch rs{scratch, B, ar.actor.hid, ar.actor.layers};
  if (t < R) {
    const int row = r0 + t;
    const bool valid = row < B;
    const float2 s = rp.s[valid ? idx[row] : 0];
    in0[t * 4 + 0] = s.x; in0[t * 4 + 1] = s.y; in0[t * 4 + 2] = 0.f; in0[t * 4 + 3] = 0.f;
    S[t * 8 + 7] = valid ? 1.f : 0.f;
  }
  cl_started(c);

  int ti = 0;
  // forward: a = pi(s) (fed the raw replay state, robot.py:386), then Q1(s, a)
  cl_forward<R, CS>(c, 0, ar.actor, true, 0, &rs, r0, tiles, ti);
  if (t < R) {
    in0[t * 4 + 2] = smem_f[c.out + t * 2];
    in0[t * 4 + 3] = smem_f[c.out + t * 2 + 1];
  }
  __syncthreads();
  cl_forward<R, CS>(c, 1, ar.critic, true, L, nullptr, r0, tiles, ti);
  if (t < R) {
    const float valid = S[t * 8 + 7];
    smem_f[c.dout + t * 2] = -valid / (float)B;
    smem_f[c.dout + t * 2 + 1] = 0.f;
    float l = -smem_f[c.out + t * 2] * valid / (float)B;
#pragma unroll
    for (int o = R / 2; o > 0; o >>= 1) l += __shfl_xor_sync((R >= 32) ? 0xffffffffu : ((1u << R) - 1u), l, o);
    if (t == 0 && c.rank == 0) atomicAdd(loss, l);
  }
  __syncthreads();
  // backward through critic 1 (only dQ/d(action) is needed), then through the actor
  cl_backward<R, CS>(c, 1, ar.critic, L, nullptr, r0, true, tiles, ti);
  if (t < R) {
    smem_f[c.dout + t * 2] = smem_f[c.din + t * 4 + 2];
    smem_f[c.dout + t * 2 + 1] = smem_f[c.din + t * 4 + 3];
    // cl_backward stores in0 of the network it differentiates: the actor's input was (s, 0, 0)
    in0[t * 4 + 2] = 0.f; in0[t * 4 + 3] = 0.f;
  }
  __syncthreads();
  cl_backward<R, CS>(c, 0, ar.actor, 0, &rs, r0, false, tiles, ti);
  stamp(c, 98);
  cluster_sync();
  stamp(c, 99);
  if (c.prof && t == 0) c.prof[0] = c.pn;
}

}  // namespace rtd3

using namespace rtd3;

namespace {
// RTD3_CLUSTER environment variable: 0 = never; "R,CS" forces rows per cluster and cluster size (development); unset: see cluster_shape
int g_mode = -1, g_force_r = 0, g_force_cs = 0;
void read_mode() {
  if (g_mode >= 0) return;
  const char* e = getenv("RTD3_CLUSTER");
  g_mode = 2;
  if (e) {
    int r = 0, cs = 0;
    if (sscanf(e, "%d,%d", &r, &cs) == 2 && (r == 8 || r == 16) && (cs == 4 || cs == 8)) { g_force_r = r; g_force_cs = cs; g_mode = 1; }
    else g_mode = atoi(e) ? 2 : 0;
  }
}
struct ClShape { int R, CS; };
// clusters of 4 with 8 rows: 32 clusters for the benchmark batch of 256 (the device holds 33 of them at once, but only 15 clusters
// of 8); larger batches run in several waves (16 rows per cluster do not fit the shared memory next to two 64 KB weight tiles)
ClShape cluster_shape(int batch) {
  read_mode();
  if (g_mode == 1) return ClShape{g_force_r, g_force_cs};
  (void)batch;
  return ClShape{8, 4};
}
long long* g_prof[2] = {nullptr, nullptr};   // development: stamp buffers of the critic / actor kernels (rtd3_debug_cluster_prof)

template <typename F>
int32_t dispatch(const ClShape& s, F&& f) {
  if (s.R == 8 && s.CS == 4) return f(std::integral_constant<int, 8>{}, std::integral_constant<int, 4>{});
  if (s.R == 16 && s.CS == 4) return f(std::integral_constant<int, 16>{}, std::integral_constant<int, 4>{});
  if (s.R == 8 && s.CS == 8) return f(std::integral_constant<int, 8>{}, std::integral_constant<int, 8>{});
  return f(std::integral_constant<int, 16>{}, std::integral_constant<int, 8>{});
}
}  // namespace

bool rtd3::cluster_path_ok(const rtd3_td3* h, int batch) {
  read_mode();
  if (g_mode == 0) return false;
  if (g_mode == 2 && batch > 1024) return false;
  const ClShape s = cluster_shape(batch);
  const int H = h->ar.critic.hid, L = h->ar.critic.layers;
  if (H % 4 != 0 || H < 4) return false;
  return make_plan(s.R, s.CS, H, L, 5, L).bytes <= kClSmemLimit && make_plan(s.R, s.CS, H, L, 2, 2 * L).bytes <= kClSmemLimit;
}

int32_t rtd3::critic_cluster_launch(rtd3_td3* h, const float* params, const float* params_t, float* scratch, const ReplayView& rp, const int32_t* idx,
                                    const float* noise, int32_t batch, const Td3Hyper& hp, float* loss2, float* q_out, float* y_out, int32_t* steps,
                                    double* beta_pows, cudaStream_t st) {
  const ClShape s = cluster_shape(batch);
  const int H = h->ar.critic.hid, L = h->ar.critic.layers;
  const ClPlan p = make_plan(s.R, s.CS, H, L, 5, L);
  const int grid = (int)ceil_div(batch, s.R) * s.CS;
  return dispatch(s, [&](auto r, auto cs) -> int32_t {
    auto* fn = td3_critic_cluster_kernel<decltype(r)::value, decltype(cs)::value>;
    RTD3_CUDA(ensure_dyn_smem((const void*)fn, p.bytes));
    fn<<<grid, kThreads, p.bytes, st>>>(h->ar, params, params_t, scratch, rp, idx, noise, batch, hp, loss2, q_out, y_out, steps, beta_pows, p.ns,
                                        g_prof[0]);
    RTD3_LAUNCHED();
    return 0;
  });
}

int32_t rtd3::actor_cluster_launch(rtd3_td3* h, const float* params, const float* params_t, float* scratch, const ReplayView& rp, const int32_t* idx,
                                   int32_t batch, float* loss1, int32_t* steps, double* beta_pows, cudaStream_t st) {
  const ClShape s = cluster_shape(batch);
  const int H = h->ar.critic.hid, L = h->ar.critic.layers;
  const ClPlan p = make_plan(s.R, s.CS, H, L, 2, 2 * L);
  const int grid = (int)ceil_div(batch, s.R) * s.CS;
  return dispatch(s, [&](auto r, auto cs) -> int32_t {
    auto* fn = td3_actor_cluster_kernel<decltype(r)::value, decltype(cs)::value>;
    RTD3_CUDA(ensure_dyn_smem((const void*)fn, p.bytes));
    fn<<<grid, kThreads, p.bytes, st>>>(h->ar, params, params_t, scratch, rp, idx, batch, loss1, steps, beta_pows, p.ns, g_prof[1]);
    RTD3_LAUNCHED();
    return 0;
  });
}

extern "C" {

int32_t rtd3_td3_cluster_supported(const rtd3_td3* h, int32_t batch) { return (h && batch > 0 && rtd3::cluster_path_ok(h, batch)) ? 1 : 0; }

int32_t rtd3_td3_cluster_occupancy(const rtd3_td3* h, int32_t batch) {
  if (!h || batch <= 0) return -1;
  const ClShape s = cluster_shape(batch);
  const int H = h->ar.critic.hid, L = h->ar.critic.layers;
  const ClPlan p = make_plan(s.R, s.CS, H, L, 5, L);
  return dispatch(s, [&](auto r, auto cs) -> int32_t {
    auto* fn = td3_critic_cluster_kernel<decltype(r)::value, decltype(cs)::value>;
    if (ensure_dyn_smem((const void*)fn, p.bytes) != cudaSuccess) return -2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ceil_div(batch, s.R) * s.CS), 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = p.bytes;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)s.CS; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&n, fn, &cfg);
    if (e != cudaSuccess) { rtd3::set_error("cudaOccupancyMaxActiveClusters: %s", cudaGetErrorString(e)); return -3; }
    return n;
  });
}

int32_t rtd3_debug_cluster_prof(long long* critic_buf, long long* actor_buf) {
  g_prof[0] = critic_buf;
  g_prof[1] = actor_buf;
  return 0;
}

}  // extern "C"
