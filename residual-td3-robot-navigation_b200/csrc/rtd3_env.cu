// Environment hot path: dynamics / step / T-step rollout / seeded init + reset.
// Reference behaviour: /root/reference/environment.py:28-56, 98-137 (see include/rtd3.h per entry point).
#include <cuda.h>

#include "rtd3_common.cuh"
#include "rtd3_env_step.cuh"
#include "rtd3_mt.cuh"


namespace rtd3 {

constexpr uint32_t kTableBytes = kCells * sizeof(float2);   // 80 000 B, a multiple of 16
constexpr int kTableChunks = 4;                             // bulk copies of 20 000 B each
static_assert(kTableBytes % (16 * kTableChunks) == 0, "bulk copy sizes must be multiples of 16 B");

// rot = float32(angle*2*pi): numpy >= 2 keeps float32*int*pyfloat in float32 (environment.py:107).
// cos/sin are taken in float64 like the reference does for (action_angle + rotation).
__global__ void build_table_kernel(const float* __restrict__ speed, const float* __restrict__ angle,
                                   float2* __restrict__ table) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kCells) return;
  float rot = __fmul_rn(__fmul_rn(angle[i], 2.0f), 3.14159274101257324f);
  double s, c;
  sincos((double)rot, &s, &c);
  double sp = (double)speed[i];
  float tc = (float)(sp * c), ts = (float)(sp * s);
  // A non-finite map cell would make the reference raise (int(nan)); here it becomes a dead cell so that the
  // state recurrence below can never leave [0, 100) and index out of the table.
  if (!isfinite(tc) || !isfinite(ts)) { tc = 0.0f; ts = 0.0f; }
  table[i] = make_float2(tc, ts);
}


// Stage the 80 KB table into shared memory with bulk-async copies signalled on one mbarrier.
// All threads of the CTA call this; returns once the table is readable.
__device__ __forceinline__ void stage_table(float2* s_table, uint64_t* bar, const float2* __restrict__ g_table) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, kTableBytes);
    constexpr uint32_t chunk = kTableBytes / kTableChunks;
#pragma unroll
    for (int c = 0; c < kTableChunks; ++c)
      bulk_g2s(reinterpret_cast<char*>(s_table) + c * chunk, reinterpret_cast<const char*>(g_table) + c * chunk, chunk,
               bar);
  }
}

// ---- single step over n envs ------------------------------------------------------------------
// Each thread owns 4 consecutive envs per iteration: four 16 B loads (x,y,ax,ay) and two 16 B stores,
// all coalesced (a warp covers 512 B contiguous per plane).  kKeepOnNan=false gives pure dynamics().
template <bool kSmem, bool kKeepOnNan>
__global__ void __launch_bounds__(kSmem ? 512 : 256, kSmem ? 2 : 4)
env_step_kernel(const float2* __restrict__ g_table, const float* __restrict__ x, const float* __restrict__ y,
                const float* __restrict__ ax, const float* __restrict__ ay, float* __restrict__ ox,
                float* __restrict__ oy, int64_t n, int vec_ok) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* s_table = reinterpret_cast<float2*>(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + kTableBytes);

  const int64_t n4 = vec_ok ? (n >> 2) : 0;   // planes not 16 B aligned (odd n): everything goes the scalar way
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const float4* y4 = reinterpret_cast<const float4*>(y);
  const float4* ax4 = reinterpret_cast<const float4*>(ax);
  const float4* ay4 = reinterpret_cast<const float4*>(ay);

  if constexpr (kSmem) stage_table(s_table, bar, g_table);

  // issue the first iteration's streaming loads before waiting for the table
  int64_t i = tid;
  float4 vx, vy, vax, vay;
  bool have = i < n4;
  if (have) { vx = __ldcs(x4 + i); vy = __ldcs(y4 + i); vax = __ldcs(ax4 + i); vay = __ldcs(ay4 + i); }
  if constexpr (kSmem) mbar_wait(bar, 0);

  auto table = [&]() {
    if constexpr (kSmem) return SmemTable{s_table};
    else return LdgTable{g_table};
  }();

  while (have) {
    const int64_t inext = i + nthreads;
    const bool have_next = inext < n4;
    float4 px, py, pax, pay;
    if (have_next) { px = __ldcs(x4 + inext); py = __ldcs(y4 + inext); pax = __ldcs(ax4 + inext); pay = __ldcs(ay4 + inext); }
    float4 rx, ry;
    step_one<kKeepOnNan>(table, vx.x, vy.x, vax.x, vay.x, rx.x, ry.x);
    step_one<kKeepOnNan>(table, vx.y, vy.y, vax.y, vay.y, rx.y, ry.y);
    step_one<kKeepOnNan>(table, vx.z, vy.z, vax.z, vay.z, rx.z, ry.z);
    step_one<kKeepOnNan>(table, vx.w, vy.w, vax.w, vay.w, rx.w, ry.w);
    __stcs(reinterpret_cast<float4*>(ox) + i, rx);
    __stcs(reinterpret_cast<float4*>(oy) + i, ry);
    i = inext; have = have_next;
    if (have) { vx = px; vy = py; vax = pax; vay = pay; }
  }

  // ragged tail (n % 4 envs), handled by the first threads of the grid
  for (int64_t j = (n4 << 2) + tid; j < n; j += nthreads) {
    float rx, ry;
    step_one<kKeepOnNan>(table, x[j], y[j], ax[j], ay[j], rx, ry);
    ox[j] = rx; oy[j] = ry;
  }
}

// ---- T-step rollout: state stays in registers, table in smem, actions prefetched kU steps ahead ----
constexpr int kU = 16;
constexpr int kDeep = 8;      // chunks in flight per thread, latency-bound launches (128 steps of lead)
constexpr int kShallow = 2;   // chunks in flight per thread, occupancy-bound launches

// The dependent chain of one rollout step.  trunc(x) comes from a round-toward-zero add of 2^23 (the low mantissa
// bits of x + 2^23 are floor(x) for 0 <= x < 2^23), so the shared-memory byte address is two IMADs away from the state:
//   addr = bits(x+2^23)*800 + bits(y+2^23)*8 + (table_base - 0x4B000000*808)
__device__ __forceinline__ void rollout_step(uint32_t addr_bias, const StepIn& in, float& x, float& y) {
  const uint32_t bx = __float_as_uint(__fadd_rz(x, 8388608.0f));
  const uint32_t by = __float_as_uint(__fadd_rz(y, 8388608.0f));
  const uint32_t addr = bx * 800u + (by * 8u + addr_bias);
  float2 cs;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(cs.x), "=f"(cs.y) : "r"(addr));
  advance(cs, in, x, y);
}

// Actions reach the SM through a per-thread cp.async ring in shared memory: kStages chunks of kU steps are in
// flight, so the HBM latency (~1 us) is hidden even when a CTA is a single warp (4096 envs on 148 SMs); every
// thread reads back only the words it copied itself, so cp.async.wait_group is the only synchronisation.
__device__ __forceinline__ void cp_async4(uint32_t smem_addr, const float* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kN>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kN) : "memory"); }

template <bool kTraj, int kStages>
__global__ void __launch_bounds__(256)
env_rollout_kernel(const float2* __restrict__ g_table, float* __restrict__ x, float* __restrict__ y,
                   const float* __restrict__ actions, float* __restrict__ traj, int64_t n, int64_t T) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* s_table = reinterpret_cast<float2*>(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + kTableBytes);
  float* ring = reinterpret_cast<float*>(smem_raw + kTableBytes + 16);   // [kStages][kU][2][blockDim]
  stage_table(s_table, bar, g_table);   // contains the only __syncthreads of the kernel

  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;                   // thread 0 of every launched CTA is live, so the bulk copy always has an owner
  // Sanitise once: inside [0, 100) the recurrence can never leave the table (finite table, clamped or kept state).
  // States outside the world are invalid input for the reference as well (IndexError at environment.py:107).
  float sx = fminf(fmaxf(x[i], 0.0f), 99.99999f);
  float sy = fminf(fmaxf(y[i], 0.0f), 99.99999f);
  const int64_t n2 = 2 * n;
  const float* pa = actions + i;        // ax(t) = pa[0], ay(t) = pa[n]; advances by 2n per step
  float* pt = traj + i;
  const uint32_t addr_bias = smem_u32(s_table) - 0x4B000000u * 808u;
  const uint32_t bd = blockDim.x;
  const uint32_t my_ring = smem_u32(ring) + threadIdx.x * 4u;
  const uint32_t stage_bytes = kU * 2u * bd * 4u;

  const int64_t full = T / kU;          // chunks of kU steps without per-step bounds checks
  auto issue_chunk = [&](int64_t c) {   // chunk c -> ring slot c % kStages
    if (c < full) {
      const float* src = pa + c * (kU * n2);
      const uint32_t dst = my_ring + (uint32_t)(c % kStages) * stage_bytes;
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        cp_async4(dst + (2 * u) * bd * 4u, src + u * n2);
        cp_async4(dst + (2 * u + 1) * bd * 4u, src + u * n2 + n);
      }
    }
    cp_async_commit();                  // always commit, so that group counting stays uniform
  };
#pragma unroll
  for (int sgi = 0; sgi < kStages; ++sgi) issue_chunk(sgi);
  mbar_wait(bar, 0);

  for (int64_t c = 0; c < full; ++c) {
    cp_async_wait<kStages - 1>();       // chunk c has landed
    StepIn in[kU];
    const uint32_t src = my_ring + (uint32_t)(c % kStages) * stage_bytes;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      float ax, ay;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ax) : "r"(src + (2 * u) * bd * 4u));
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ay) : "r"(src + (2 * u + 1) * bd * 4u));
      in[u] = prep_action(ax, ay);
    }
    issue_chunk(c + kStages);           // refill the slot just drained (its values are in registers now)
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      rollout_step(addr_bias, in[u], sx, sy);
      if (kTraj) { __stcs(pt, sx); __stcs(pt + n, sy); pt += n2; }
    }
  }
  pa += full * (kU * n2);
  for (int64_t t = full * kU; t < T; ++t) {   // ragged tail, T % kU steps
    const StepIn in = prep_action(__ldcs(pa), __ldcs(pa + n));
    pa += n2;
    rollout_step(addr_bias, in, sx, sy);
    if (kTraj) { __stcs(pt, sx); __stcs(pt + n, sy); pt += n2; }
  }
  cp_async_wait<0>();
  x[i] = sx;
  y[i] = sy;
}

// ---- T-step rollout, TMA variant (n % 4 == 0): one warp = 32 envs = one independent pipeline -------------------------
// Actions [T][2][n] and trajectories are 2-D tensors (inner dim n, outer dim 2T) described by CUtensorMaps; a warp moves
// its [kU steps x 2][32 envs] tile (32 rows of 128 B) with ONE cp.async.bulk.tensor.2d (SASS UTMALDG / UTMASTG) per
// chunk and direction: loads complete on a per-stage mbarrier, stores are tracked by the bulk async-group.  This removes
// the per-step LDGSTS/STG and their 64-bit address arithmetic from the single warp that has to issue everything
// (ncu r1: 47 instr/step at IPC 0.38 bounded the cp.async version; 32 row-wise 128 B bulk copies per chunk were tried
// first and were 2.6x SLOWER - the TMA unit is op-bound, not byte-bound, at that size).  Out-of-range rows / envs of
// edge tiles are zero-filled on load and clipped on store by the TMA unit itself.
constexpr int kTileFloats = kU * 2 * 32;          // 1024 floats = 4 KB per warp per stage

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(tm)),
               "r"(c0), "r"(c1), "r"(smem_u32(smem_src))
               : "memory");
}

template <bool kTraj, int kStages>
__global__ void __launch_bounds__(256)
env_rollout_tma_kernel(const float2* __restrict__ g_table, float* __restrict__ x, float* __restrict__ y,
                       const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_traj, int64_t n, int64_t T) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* s_table = reinterpret_cast<float2*>(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + kTableBytes);            // table barrier
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kTableBytes + 16) + warp * kStages;     // [nwarps][kStages]
  // tiles start on a 128 B boundary behind the barriers
  float* tiles = reinterpret_cast<float*>(smem_raw + ((kTableBytes + 16 + nwarps * kStages * 8 + 127) & ~127u));
  float* my_in = tiles + (size_t)warp * kStages * kTileFloats;                                     // action ring of this warp
  float* my_out = tiles + (size_t)nwarps * kStages * kTileFloats + (size_t)warp * 2 * kTileFloats; // 2 trajectory staging tiles
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(full + s, 1);
    fence_mbar_init();
  }
  stage_table(s_table, bar, g_table);   // init + fence + __syncthreads, then the table's bulk copies

  const int64_t i0 = ((int64_t)blockIdx.x * nwarps + warp) * 32;                  // first env of this warp
  if (i0 >= n) return;
  const int64_t i = i0 + lane;
  const bool live = i < n;
  float sx = live ? fminf(fmaxf(x[i], 0.0f), 99.99999f) : 0.0f;
  float sy = live ? fminf(fmaxf(y[i], 0.0f), 99.99999f) : 0.0f;
  const uint32_t addr_bias = smem_u32(s_table) - 0x4B000000u * 808u;
  const int64_t chunks = (T + kU - 1) / kU;

  auto issue_chunk = [&](int64_t c) {
    if (c < chunks && lane == 0) {
      const int s = (int)(c % kStages);
      mbar_arrive_expect_tx(full + s, kTileFloats * 4);
      tma_load_2d(my_in + s * kTileFloats, &tm_act, (int)i0, (int)(c * (2 * kU)), full + s);
    }
  };
#pragma unroll
  for (int s = 0; s < kStages; ++s) issue_chunk(s);
  mbar_wait(bar, 0);

  for (int64_t c = 0; c < chunks; ++c) {
    const int s = (int)(c % kStages);
    mbar_wait(full + s, (uint32_t)((c / kStages) & 1));
    float cax[kU], cay[kU];
    float nanacc = 0.0f;
    const float* tin = my_in + s * kTileFloats + lane;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      cax[u] = tin[(2 * u) * 32];
      cay[u] = tin[(2 * u + 1) * 32];
      nanacc = fmaf(cax[u], 0.0f, nanacc);      // stays 0 unless an action is NaN or inf
      nanacc = fmaf(cay[u], 0.0f, nanacc);
    }
    __syncwarp();                               // every lane holds its values: the slot can be refilled
    issue_chunk(c + kStages);
    float* tout = my_out + (c & 1) * kTileFloats + lane;
    if (kTraj && lane == 0) bulk_wait_read<1>();   // the staging tile used two chunks ago has been read by its store
    __syncwarp();
    const int steps = (int)min((int64_t)kU, T - c * kU);
    const bool careful = __any_sync(0xffffffffu, nanacc != 0.0f) || steps < kU;
    if (!careful) {
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        StepIn in;
        in.ax = fminf(fmaxf(cax[u], -kMaxAction), kMaxAction);
        in.ay = fminf(fmaxf(cay[u], -kMaxAction), kMaxAction);
        in.lo = 0.0f; in.hi = kClipHi; in.bad = false;
        rollout_step(addr_bias, in, sx, sy);
        if (kTraj) { tout[(2 * u) * 32] = sx; tout[(2 * u + 1) * 32] = sy; }
      }
    } else {
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (u < steps) {
          const StepIn in = prep_action(cax[u], cay[u]);
          rollout_step(addr_bias, in, sx, sy);
        }
        if (kTraj) { tout[(2 * u) * 32] = sx; tout[(2 * u + 1) * 32] = sy; }
      }
    }
    if (kTraj) {
      fence_proxy_async();                      // generic-proxy writes of the tile -> visible to the TMA (async proxy) read
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tm_traj, (int)i0, (int)(c * (2 * kU)), my_out + (c & 1) * kTileFloats);
        bulk_commit();
      }
    }
  }
  if (live) { x[i] = sx; y[i] = sy; }
  if (kTraj && lane == 0) bulk_wait<0>();       // the stores must have left shared memory before the CTA exits
}

// ---- T-step rollout, warp-pair variant: chain warp + helper warp per 32 envs -------------------------------------------
// At 4096 envs every SM gets one warp, and that warp's issue slots - not memory - bound the TMA kernel above (31
// instr/step, of which only 13 are the state recurrence; ncu r1: 42 % fixed-latency waits).  Here the recurrence runs
// alone on the CHAIN warp (per step: one LDS.64 of the pre-clipped action, the 13-instruction chain, two STS of the new
// state); a HELPER warp on another scheduler of the same SM waits for the TMA action tiles, clips them (or flags the
// chunk for the careful NaN path), re-arms the loads and issues the trajectory tile stores.  The two meet on mbarriers
// over double-buffered shared-memory tiles.
constexpr int kPairU = 32;        // steps per tile of the pair kernel
constexpr int kPairStages = 4;    // action tiles in flight (128 steps of lead, as kDeep x kU)
struct PairSmem {
  static constexpr int kBars = 16;   // full[8] | prepped[2] | consumed[2] | outfull[2] | outfree[2]
};

template <bool kTraj, int kStages, int kUp>
__global__ void __launch_bounds__(256)
env_rollout_pair_kernel(const float2* __restrict__ g_table, float* __restrict__ x, float* __restrict__ y,
                        const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_traj, int64_t n, int64_t T) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int kU = kUp;                            // steps per action / trajectory tile of this kernel (shadows the file-wide kU)
  constexpr int kTileFloats = kU * 2 * 32;
  float2* s_table = reinterpret_cast<float2*>(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + kTableBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, npairs = blockDim.x >> 6;
  const int pair = warp >> 1;
  const bool is_chain = (warp & 1) == 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kTableBytes + 16) + pair * PairSmem::kBars;
  uint64_t* full = bars, *prepped = bars + 8, *consumed = bars + 10, *outfull = bars + 12, *outfree = bars + 14;
  // per pair: kStages raw tiles | 2 prepped tiles (float2 per step and env = 2 tiles' worth each... [kU][32] float2 = 4 KB) | 2 out tiles
  float* tiles = reinterpret_cast<float*>(smem_raw + ((kTableBytes + 16 + npairs * PairSmem::kBars * 8 + 127) & ~127u));
  float* my = tiles + (size_t)pair * (kStages + 4) * kTileFloats;
  float* raw = my;                                   // [kStages][kU*2][32]
  float2* prep = reinterpret_cast<float2*>(my + kStages * kTileFloats);                 // [2][kU][32]
  float* outt = my + (kStages + 2) * kTileFloats;    // [2][kU*2][32]
  __shared__ int s_flag[4][2];                       // careful flag per pair and prepped buffer
  if (lane == 0 && is_chain) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(full + s, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(prepped + b, 32); mbar_init(consumed + b, 32); mbar_init(outfull + b, 32); mbar_init(outfree + b, 1); }
    fence_mbar_init();
  }
  stage_table(s_table, bar, g_table);   // contains the __syncthreads that publishes the barrier inits

  const int64_t i0 = ((int64_t)blockIdx.x * npairs + pair) * 32;
  if (i0 >= n) return;
  const int64_t chunks = (T + kU - 1) / kU;

  if (!is_chain) {
    // ================================= helper warp =================================
    auto issue_load = [&](int64_t c) {
      if (c < chunks && lane == 0) {
        const int s = (int)(c % kStages);
        mbar_arrive_expect_tx(full + s, kTileFloats * 4);
        tma_load_2d(raw + s * kTileFloats, &tm_act, (int)i0, (int)(c * (2 * kU)), full + s);
      }
    };
#pragma unroll
    for (int s = 0; s < kStages; ++s) issue_load(s);
    for (int64_t c = 0; c < chunks; ++c) {
      const int s = (int)(c % kStages), b = (int)(c & 1);
      mbar_wait(full + s, (uint32_t)((c / kStages) & 1));
      mbar_wait(consumed + b, (uint32_t)(((c >> 1) & 1) ^ 1));       // the chain warp is done with prepped tile b (first two pass)
      const float* tin = raw + s * kTileFloats + lane;
      float ax[kU], ay[kU], nanacc = 0.0f;
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        ax[u] = tin[(2 * u) * 32];
        ay[u] = tin[(2 * u + 1) * 32];
        nanacc = fmaf(ax[u], 0.0f, nanacc);
        nanacc = fmaf(ay[u], 0.0f, nanacc);
      }
      const bool careful = __any_sync(0xffffffffu, nanacc != 0.0f) || (T - c * kU) < kU;
      float2* tp = prep + b * (kU * 32) + lane;
#pragma unroll
      for (int u = 0; u < kU; ++u)
        tp[u * 32] = careful ? make_float2(ax[u], ay[u])
                             : make_float2(fminf(fmaxf(ax[u], -kMaxAction), kMaxAction), fminf(fmaxf(ay[u], -kMaxAction), kMaxAction));
      if (lane == 0) s_flag[pair][b] = careful ? 1 : 0;
      __syncwarp();
      issue_load(c + kStages);                                       // raw stage s is in registers / prepped now
      mbar_arrive(prepped + b);
      if (kTraj && c >= 1) {                                         // trajectory tile of the previous chunk
        const int pb = (int)((c - 1) & 1);
        mbar_wait(outfull + pb, (uint32_t)(((c - 1) >> 1) & 1));
        fence_proxy_async();
        if (lane == 0) {
          tma_store_2d(&tm_traj, (int)i0, (int)((c - 1) * (2 * kU)), outt + pb * kTileFloats);
          bulk_commit();
          bulk_wait_read<1>();                                       // the store of chunk c-2 has read its tile ...
          if (c >= 2) mbar_arrive(outfree + (int)(c & 1));           // ... which is the tile chunk c will write
        }
        __syncwarp();
      }
    }
    if (kTraj) {
      const int pb = (int)((chunks - 1) & 1);
      mbar_wait(outfull + pb, (uint32_t)(((chunks - 1) >> 1) & 1));
      fence_proxy_async();
      if (lane == 0) {
        tma_store_2d(&tm_traj, (int)i0, (int)((chunks - 1) * (2 * kU)), outt + pb * kTileFloats);
        bulk_commit();
        bulk_wait<0>();
      }
      __syncwarp();
    }
    return;
  }

  // ================================= chain warp =================================
  const int64_t i = i0 + lane;
  const bool live = i < n;
  float sx = live ? fminf(fmaxf(x[i], 0.0f), 99.99999f) : 0.0f;
  float sy = live ? fminf(fmaxf(y[i], 0.0f), 99.99999f) : 0.0f;
  const uint32_t addr_bias = smem_u32(s_table) - 0x4B000000u * 808u;
  mbar_wait(bar, 0);
  for (int64_t c = 0; c < chunks; ++c) {
    const int b = (int)(c & 1);
    mbar_wait(prepped + b, (uint32_t)((c >> 1) & 1));
    if (kTraj && c >= 2) mbar_wait(outfree + b, (uint32_t)(((c >> 1) - 1) & 1));   // out tile b has been stored (chunk c-2)
    const bool careful = s_flag[pair][b] != 0;
    const float2* tp = prep + b * (kU * 32) + lane;
    float* tout = outt + b * kTileFloats + lane;
    if (!careful) {
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const float2 a = tp[u * 32];
        StepIn in;
        in.ax = a.x; in.ay = a.y; in.lo = 0.0f; in.hi = kClipHi; in.bad = false;
        rollout_step(addr_bias, in, sx, sy);
        if (kTraj) { tout[(2 * u) * 32] = sx; tout[(2 * u + 1) * 32] = sy; }
      }
    } else {
      const int steps = (int)min((int64_t)kU, T - c * kU);
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (u < steps) {
          const float2 a = tp[u * 32];
          const StepIn in = prep_action(a.x, a.y);
          rollout_step(addr_bias, in, sx, sy);
        }
        if (kTraj) { tout[(2 * u) * 32] = sx; tout[(2 * u + 1) * 32] = sy; }
      }
    }
    mbar_arrive(consumed + b);
    if (kTraj) mbar_arrive(outfull + b);
  }
  if (live) { x[i] = sx; y[i] = sy; }
}

// ---- seeded init / reset on per-env legacy MT19937 streams ------------------------------------
__global__ void mt_seed_kernel(rtd3_mt_bank b, const uint32_t* __restrict__ seeds) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n) return;
  MtStream s{b.mt + i, b.n, 0};
  s.seed(seeds[i]);
  b.pos[i] = s.pos;
  b.has_gauss[i] = 0;
  b.gauss[i] = 0.0;
}

__global__ void mt_draw_u32_kernel(rtd3_mt_bank b, uint32_t* __restrict__ out, int64_t k) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n) return;
  MtStream s{b.mt + i, b.n, b.pos[i]};
  for (int64_t j = 0; j < k; ++j) out[j * b.n + i] = s.next_u32();
  b.pos[i] = s.pos;
}

__global__ void mt_draw_gauss_kernel(rtd3_mt_bank b, double* __restrict__ out, int64_t k, const int8_t* __restrict__ where, int equals) {
  // legacy_gauss (rtd3_mt.cuh: mt_gauss) for one stream per lane, warp-synchronous so that wrapping streams are twisted by the whole
  // warp: this is the exploration noise of every tick in the exact-noise mode, and ~1 % of the streams wrap per call
  // `where` (nullable): only the streams with where[i] == equals draw - the others keep their position and get zeros (the
  // reference draws its exploration noise on 'step' ticks only, robot.py:560 via robot-learning.py:97)
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool inside = i < b.n;
  const bool active = inside && (!where || (int)where[i] == equals);
  if (inside && !active)
    for (int64_t j = 0; j < k; ++j) out[j * b.n + i] = 0.0;
  if (!__any_sync(0xffffffffu, active)) return;
  const int64_t ii = active ? i : 0;
  MtStream s{b.mt + ii, b.n, active ? b.pos[ii] : 0};
  int hg = active ? b.has_gauss[ii] : 0;
  double sp = active ? b.gauss[ii] : 0.0;
  for (int64_t j = 0; j < k; ++j) {
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
    if (active) out[j * b.n + i] = v;
  }
  if (!active) return;
  b.pos[i] = s.pos;
  b.has_gauss[i] = hg;
  b.gauss[i] = sp;
}

// environment.py:28-56
__global__ void init_goal_region_kernel(rtd3_mt_bank b, double* __restrict__ goal, double* __restrict__ region) {
  // Warp-synchronous: the goal rejection loop (acceptance ~3 %) consumes ~130 words per env on average, so a fifth of the streams
  // wrap at least once in here; with the per-lane serial twist this launch took 19 ms for 65 536 envs (ncu launch list r1).
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = b.n;
  const bool active = i < n;
  if (!__any_sync(0xffffffffu, active)) return;
  const int64_t ii = active ? i : 0;
  MtStream s{b.mt + ii, n, active ? b.pos[ii] : 0};
  auto uni = [](double lo, double hi, double u) { return __dadd_rn(lo, __dmul_rn(__dsub_rn(hi, lo), u)); };   // numpy's two roundings
  const double W = (double)kWorld, R = 25.0;   // constants.py:6, 25
  const uint32_t side = mt_next_u32_warp(s, active) & 3u;   // np.random.choice([0,1,2,3]): one masked draw, never rejected
  const double free_edge = uni(0.0, W - R, mt_next_double_warp(s, active));
  double l, r, bt, tp;
  if (side == 0)      { l = 0.0;       r = R;            bt = free_edge; tp = __dadd_rn(free_edge, R); }
  else if (side == 1) { l = free_edge; r = __dadd_rn(free_edge, R); bt = W - R; tp = W; }
  else if (side == 2) { l = W - R;     r = W;            bt = free_edge; tp = __dadd_rn(free_edge, R); }
  else                { l = free_edge; r = __dadd_rn(free_edge, R); bt = 0.0;   tp = R; }
  const double mx = __dmul_rn(0.5, __dadd_rn(l, r)), my = __dmul_rn(0.5, __dadd_rn(bt, tp));
  double gx = 0.0, gy = 0.0;
  bool searching = active;
  while (__any_sync(0xffffffffu, searching)) {
    double ux, uy;
    mt_next_double2_warp(s, searching, ux, uy);
    if (searching) {
      gx = uni(5.0, W - 5.0, ux);
      gy = uni(5.0, W - 5.0, uy);
      if (!(norm2_np(__dsub_rn(gx, mx), __dsub_rn(gy, my)) < 90.0)) searching = false;
    }
  }
  if (!active) return;
  goal[i] = gx; goal[n + i] = gy;
  region[i] = l; region[n + i] = r; region[2 * n + i] = bt; region[3 * n + i] = tp;
  b.pos[i] = s.pos;
}

// environment.py:130-137
__global__ void env_reset_kernel(rtd3_mt_bank b, const double* __restrict__ region, const uint8_t* __restrict__ mask, int mask_equals,
                                 float* __restrict__ x, float* __restrict__ y, double* __restrict__ state64) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < b.n && (!mask || (mask_equals >= 0 ? mask[i] == (uint8_t)mask_equals : mask[i] != 0));
  reset_env_warp(b, region, active, i, x, y, state64);
}

static int32_t check_bank(const rtd3_mt_bank* b) {
  RTD3_CHECK_ARG(b && b->mt && b->pos && b->has_gauss && b->gauss, "null MT bank pointer");
  RTD3_CHECK_ARG(b->n >= 0, "negative stream count");
  return 0;
}

}  // namespace rtd3

using namespace rtd3;

static bool g_force_plain_rollout = false;
static int g_rollout_variant = 0;            // test hook: 1 = single-warp TMA kernel instead of the warp-pair kernel

// cuTensorMapEncodeTiled comes from the driver (libcuda); it is looked up at run time so that librtd3.so links against
// the runtime only.  If the lookup fails the rollout falls back to the cp.async kernel.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;

static void load_encode_tiled() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
    g_encode_tiled = (EncodeTiledFn)fn;
}

// [T][2][n] float32 planes as a 2-D tensor: inner dim n, outer dim 2T, box = 32 envs x 2*kU rows
static int32_t make_plane_map(CUtensorMap* tm, const float* base, int64_t n, int64_t T, int steps_per_tile = kU) {
  const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)(2 * T)};
  const cuuint64_t strides[1] = {(cuuint64_t)n * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)(2 * steps_per_tile)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    rtd3::set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return RTD3_ERR_STATE;
  }
  return 0;
}   // test hook: exercise the cp.async variant on aligned sizes too

extern "C" {

void rtd3_env_force_plain_rollout(int32_t on) { g_force_plain_rollout = on == 1; g_rollout_variant = on == 2 ? 1 : 0; }

int32_t rtd3_env_create(rtd3_env** out, int32_t device) {
  RTD3_CHECK_ARG(out, "out is null");
  int prev = 0;
  RTD3_CUDA(cudaGetDevice(&prev));
  RTD3_CUDA(cudaSetDevice(device));
  rtd3_env* h = new rtd3_env();
  h->device = device;
  h->has_map = false;
  h->table = nullptr;
  h->host_pipe = nullptr;
  RTD3_CUDA(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
  RTD3_CUDA(cudaMalloc(&h->table, kTableBytes));
  if (!g_encode_tiled) load_encode_tiled();
  const int smem = kTableBytes + 16;
  RTD3_CUDA(cudaFuncSetAttribute(env_step_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  RTD3_CUDA(cudaFuncSetAttribute(env_step_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int smem_roll = 227 * 1024;
  RTD3_CUDA(cudaFuncSetAttribute(env_rollout_kernel<true, kDeep>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_roll));
  RTD3_CUDA(cudaFuncSetAttribute(env_rollout_kernel<false, kDeep>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_roll));
  RTD3_CUDA(cudaFuncSetAttribute(env_rollout_kernel<true, kShallow>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_roll));
  RTD3_CUDA(cudaFuncSetAttribute(env_rollout_kernel<false, kShallow>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_roll));
  RTD3_CUDA(cudaFuncSetAttribute(env_rollout_pair_kernel<true, kPairStages, kPairU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_roll - 1024));
  RTD3_CUDA(cudaFuncSetAttribute(env_rollout_pair_kernel<false, kPairStages, kPairU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_roll - 1024));
  RTD3_CUDA(cudaFuncSetAttribute(env_rollout_tma_kernel<true, kDeep>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_roll));
  RTD3_CUDA(cudaFuncSetAttribute(env_rollout_tma_kernel<false, kDeep>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_roll));
  RTD3_CUDA(cudaFuncSetAttribute(env_rollout_tma_kernel<true, kShallow>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_roll));
  RTD3_CUDA(cudaFuncSetAttribute(env_rollout_tma_kernel<false, kShallow>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_roll));
  RTD3_CUDA(cudaSetDevice(prev));
  *out = h;
  return 0;
}

int32_t rtd3_env_destroy(rtd3_env* h) {
  if (!h) return 0;
  rtd3::host_pipe_destroy(h);
  if (h->table) cudaFree(h->table);
  delete h;
  return 0;
}

int32_t rtd3_env_set_map(rtd3_env* h, const float* speed, const float* angle, void* stream) {
  RTD3_CHECK_ARG(h && speed && angle, "null argument");
  build_table_kernel<<<(int)ceil_div(kCells, 256), 256, 0, (cudaStream_t)stream>>>(speed, angle, h->table);
  RTD3_LAUNCHED();
  h->has_map = true;
  return 0;
}

static int32_t launch_step(rtd3_env* h, const float* x, const float* y, const float* ax, const float* ay, float* ox,
                           float* oy, int64_t n, int32_t variant, bool keep_on_nan, cudaStream_t st) {
  RTD3_CHECK_ARG(h && h->has_map, "environment has no dynamics map (call rtd3_env_set_map)");
  RTD3_CHECK_ARG(n >= 0, "negative n");
  if (n == 0) return 0;
  RTD3_CHECK_ARG(x && y && ax && ay && ox && oy, "null state/action pointer");
  const int vec_ok =
      (((uintptr_t)x | (uintptr_t)y | (uintptr_t)ax | (uintptr_t)ay | (uintptr_t)ox | (uintptr_t)oy) % 16 == 0) ? 1 : 0;
  RTD3_CHECK_ARG(variant >= RTD3_STEP_AUTO && variant <= RTD3_STEP_LDG, "unknown step variant");
  const int64_t n4 = vec_ok ? (n + 3) / 4 : n;
  if (variant == RTD3_STEP_AUTO) variant = (n >= 262144) ? RTD3_STEP_SMEM : RTD3_STEP_LDG;
  if (variant == RTD3_STEP_SMEM) {
    const int block = 512;
    const int grid = (int)std::min<int64_t>(ceil_div(n4, block), (int64_t)h->num_sms * 2);
    const int smem = kTableBytes + 16;
    if (keep_on_nan) env_step_kernel<true, true><<<grid, block, smem, st>>>(h->table, x, y, ax, ay, ox, oy, n, vec_ok);
    else env_step_kernel<true, false><<<grid, block, smem, st>>>(h->table, x, y, ax, ay, ox, oy, n, vec_ok);
  } else {
    const int block = 256;
    const int grid = (int)std::min<int64_t>(ceil_div(n4, block), (int64_t)h->num_sms * 4);
    if (keep_on_nan) env_step_kernel<false, true><<<grid, block, 0, st>>>(h->table, x, y, ax, ay, ox, oy, n, vec_ok);
    else env_step_kernel<false, false><<<grid, block, 0, st>>>(h->table, x, y, ax, ay, ox, oy, n, vec_ok);
  }
  RTD3_LAUNCHED();
  return 0;
}

int32_t rtd3_env_step(rtd3_env* h, float* x, float* y, const float* ax, const float* ay, int64_t n, int32_t variant,
                      void* stream) {
  return launch_step(h, x, y, ax, ay, x, y, n, variant, true, (cudaStream_t)stream);
}

int32_t rtd3_env_dynamics(rtd3_env* h, const float* x, const float* y, const float* ax, const float* ay, float* out_x,
                          float* out_y, int64_t n, void* stream) {
  return launch_step(h, x, y, ax, ay, out_x, out_y, n, RTD3_STEP_AUTO, false, (cudaStream_t)stream);
}

int32_t rtd3_env_rollout(rtd3_env* h, float* x, float* y, const float* actions, float* traj, int64_t n, int64_t T,
                         void* stream) {
  RTD3_CHECK_ARG(h && h->has_map, "environment has no dynamics map (call rtd3_env_set_map)");
  RTD3_CHECK_ARG(n >= 0 && T >= 0, "negative n or T");
  if (n == 0 || T == 0) return 0;
  RTD3_CHECK_ARG(x && y && actions, "null state/action pointer");
  // Small batches are latency-bound (one dependent chain per env): spread them over all SMs in CTAs as small as a
  // warp and keep kDeep chunks of actions in flight per thread.  Large batches hide latency with occupancy instead:
  // 128-thread CTAs, kShallow chunks in flight, up to 2 CTAs per SM next to the 80 KB table.
  const int64_t per_sm = ceil_div(n, (int64_t)h->num_sms);
  const bool deep = per_sm <= 64;
  const int block = deep ? (int)std::max<int64_t>(32, ceil_div(per_sm, 32) * 32) : 128;
  const int grid = (int)ceil_div(n, block);
  const int stages = deep ? kDeep : kShallow;
  const int smem = kTableBytes + 16 + stages * kU * 2 * block * 4;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tma_ok = (n % 4 == 0) && (n < (1ll << 31)) && (2 * T < (1ll << 31)) && ((uintptr_t)actions % 16 == 0) &&
                      (!traj || (uintptr_t)traj % 16 == 0) && !g_force_plain_rollout && g_encode_tiled;
  CUtensorMap tm_act, tm_traj;
  // an encode failure (unexpected shape limits) is not an error: the cp.async kernel below handles every shape
  if (tma_ok && make_plane_map(&tm_act, actions, n, T) == 0 && make_plane_map(&tm_traj, traj ? traj : actions, n, T) == 0) {
    // one warp per 32 envs; few warps per CTA while the batch cannot fill the SMs, 8 once it can
    const int64_t warps_total = ceil_div(n, 32);
    const int wpc = deep ? (int)std::max<int64_t>(1, std::min<int64_t>(8, ceil_div(warps_total, (int64_t)h->num_sms))) : 8;
    const int bgrid = (int)ceil_div(warps_total, wpc);
    const int bsmem = ((kTableBytes + 16 + wpc * stages * 8 + 127) & ~127) + wpc * (stages + 2) * kTileFloats * 4;
    if (deep && g_rollout_variant != 1) {
      // latency-bound batches: one chain warp + one helper warp per 32 envs, up to 2 pairs per CTA; tiles of kPairU steps
      // (the per-tile hand-over - two barrier waits, the flag read, two arrives - was ~20 % of the chain warp's samples at 16)
      const int ppc = (int)std::max<int64_t>(1, std::min<int64_t>(2, ceil_div(warps_total, (int64_t)h->num_sms)));
      const int pgrid = (int)ceil_div(warps_total, ppc);
      const int psmem = ((kTableBytes + 16 + ppc * PairSmem::kBars * 8 + 127) & ~127) + ppc * (kPairStages + 4) * (kPairU * 2 * 32) * 4;
      if (make_plane_map(&tm_act, actions, n, T, kPairU) != 0 || make_plane_map(&tm_traj, traj ? traj : actions, n, T, kPairU) != 0) return RTD3_ERR_STATE;
      if (traj) env_rollout_pair_kernel<true, kPairStages, kPairU><<<pgrid, ppc * 64, psmem, st>>>(h->table, x, y, tm_act, tm_traj, n, T);
      else env_rollout_pair_kernel<false, kPairStages, kPairU><<<pgrid, ppc * 64, psmem, st>>>(h->table, x, y, tm_act, tm_traj, n, T);
    } else if (deep) {
      if (traj) env_rollout_tma_kernel<true, kDeep><<<bgrid, wpc * 32, bsmem, st>>>(h->table, x, y, tm_act, tm_traj, n, T);
      else env_rollout_tma_kernel<false, kDeep><<<bgrid, wpc * 32, bsmem, st>>>(h->table, x, y, tm_act, tm_traj, n, T);
    } else {
      if (traj) env_rollout_tma_kernel<true, kShallow><<<bgrid, wpc * 32, bsmem, st>>>(h->table, x, y, tm_act, tm_traj, n, T);
      else env_rollout_tma_kernel<false, kShallow><<<bgrid, wpc * 32, bsmem, st>>>(h->table, x, y, tm_act, tm_traj, n, T);
    }
    RTD3_LAUNCHED();
    return 0;
  }
  if (deep) {
    if (traj) env_rollout_kernel<true, kDeep><<<grid, block, smem, st>>>(h->table, x, y, actions, traj, n, T);
    else env_rollout_kernel<false, kDeep><<<grid, block, smem, st>>>(h->table, x, y, actions, nullptr, n, T);
  } else {
    if (traj) env_rollout_kernel<true, kShallow><<<grid, block, smem, st>>>(h->table, x, y, actions, traj, n, T);
    else env_rollout_kernel<false, kShallow><<<grid, block, smem, st>>>(h->table, x, y, actions, nullptr, n, T);
  }
  RTD3_LAUNCHED();
  return 0;
}

int32_t rtd3_mt_seed(const rtd3_mt_bank* bank, const uint32_t* seeds, void* stream) {
  if (int32_t e = check_bank(bank)) return e;
  RTD3_CHECK_ARG(seeds, "null seeds");
  if (bank->n == 0) return 0;
  mt_seed_kernel<<<(int)ceil_div(bank->n, 128), 128, 0, (cudaStream_t)stream>>>(*bank, seeds);
  RTD3_LAUNCHED();
  return 0;
}

int32_t rtd3_mt_draw_u32(const rtd3_mt_bank* bank, uint32_t* out, int64_t k, void* stream) {
  if (int32_t e = check_bank(bank)) return e;
  RTD3_CHECK_ARG(out && k >= 0, "bad out/k");
  if (bank->n == 0 || k == 0) return 0;
  mt_draw_u32_kernel<<<(int)ceil_div(bank->n, 128), 128, 0, (cudaStream_t)stream>>>(*bank, out, k);
  RTD3_LAUNCHED();
  return 0;
}

int32_t rtd3_mt_draw_gauss(const rtd3_mt_bank* bank, double* out, int64_t k, void* stream) {
  return rtd3_mt_draw_gauss_where(bank, out, k, nullptr, 0, stream);
}

int32_t rtd3_mt_draw_gauss_where(const rtd3_mt_bank* bank, double* out, int64_t k, const int8_t* where, int32_t equals, void* stream) {
  if (int32_t e = check_bank(bank)) return e;
  RTD3_CHECK_ARG(out && k >= 0, "bad out/k");
  if (bank->n == 0 || k == 0) return 0;
  mt_draw_gauss_kernel<<<(int)ceil_div(bank->n, 128), 128, 0, (cudaStream_t)stream>>>(*bank, out, k, where, equals);
  RTD3_LAUNCHED();
  return 0;
}

int32_t rtd3_env_init_goal_region(const rtd3_mt_bank* bank, double* goal, double* region, void* stream) {
  if (int32_t e = check_bank(bank)) return e;
  RTD3_CHECK_ARG(goal && region, "null output");
  if (bank->n == 0) return 0;
  init_goal_region_kernel<<<(int)ceil_div(bank->n, 128), 128, 0, (cudaStream_t)stream>>>(*bank, goal, region);
  RTD3_LAUNCHED();
  return 0;
}

int32_t rtd3_env_reset(const rtd3_mt_bank* bank, const double* region, const uint8_t* mask, int32_t mask_equals, float* x, float* y,
                       double* state64, void* stream) {
  if (int32_t e = check_bank(bank)) return e;
  RTD3_CHECK_ARG(region && x && y, "null argument");
  if (bank->n == 0) return 0;
  env_reset_kernel<<<(int)ceil_div(bank->n, 128), 128, 0, (cudaStream_t)stream>>>(*bank, region, mask, mask_equals, x, y, state64);
  RTD3_LAUNCHED();
  return 0;
}

}  // extern "C"
