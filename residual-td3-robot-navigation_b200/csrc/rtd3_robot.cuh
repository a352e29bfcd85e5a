// process_transition for one env per thread (robot.py:645-675, 727-762, 509-538) as a device function, shared by
// robot_transition_kernel (rtd3_robot.cu) and the fused tick kernel (rtd3_tick.cu).
#pragma once
#include "rtd3_common.cuh"
#include "rtd3_mt.cuh"

namespace rtd3 {

constexpr int kStuckSteps = 5;          // robot.py:40
constexpr double kStuckThreshold = 2.0; // robot.py:39
constexpr double kStuckPenalty = 50.0;  // robot.py:41
constexpr double kGoalReward = 50.0;    // robot.py:42
constexpr double kGoalRadius = 5.0;     // constants.py:50
constexpr int kNumDemo = 3;             // robot.py:22

struct RobotState {
  const double* goal;       // [2][n]
  float* hist;              // [5][2][n] ring of pre-step states (robot.py:425, 509-538)
  int32_t* hist_count;      // [n]
  int32_t* hist_head;       // [n] slot of the oldest entry
  uint8_t* goal_reached;    // [n]
  uint8_t* stuck_flag;      // [n]
  const uint8_t* demo_flag; // [n]
  const int32_t* plan_index;
  const int32_t* path_length;
};

struct ReplayRing {
  float2* s; float2* a; float* r; float2* s2; float* notdone;
  int64_t capacity, position;
  unsigned long long* total;   // device counter of rows ever pushed (nullable); authoritative for masked pushes
};

constexpr int kDemoGrid = RTD3_DEMO_GRID;             // cells per side of the demonstration-state grid
constexpr double kDemoCell = RTD3_DEMO_CELL;           // cell side (a power of two: cell edges are exact in float64)
constexpr int kDemoCells = kDemoGrid * kDemoGrid;
constexpr int64_t kDemoStageMax = 13000;               // points that fit the shared-memory copy (208 KB) next to the static arrays

// baseline_action = state - goal (robot.py:556 / 586), cast to float32 as torch.FloatTensor does (robot.py:612)
__device__ __forceinline__ float2 baseline_env(float x, float y, double gx, double gy) {
  return make_float2((float)((double)x - gx), (float)((double)y - gy));
}

// action = clip(baseline + residual + noise, +-5)   (robot.py:560-567 / 590-593); noise = unit_normal * noise_scale * 5 (robot.py:640)
__device__ __forceinline__ void compose_env(float x, float y, double gx, double gy, float2 res, bool has_noise, double zx, double zy,
                                            double noise_scale, double& cx, double& cy) {
  cx = __dadd_rn(__dsub_rn((double)x, gx), (double)res.x);
  cy = __dadd_rn(__dsub_rn((double)y, gy), (double)res.y);
  if (has_noise) {
    const double sc = __dmul_rn(noise_scale, 5.0);
    cx = __dadd_rn(cx, __dmul_rn(sc, zx));
    cy = __dadd_rn(cy, __dmul_rn(sc, zy));
  }
  cx = cx < -5.0 ? -5.0 : (cx > 5.0 ? 5.0 : cx);
  cy = cy < -5.0 ? -5.0 : (cy > 5.0 ? 5.0 : cy);
}

// get_next_action_type + reset for env i (robot.py:443-506).  Returns 0 'step', 1 'demo', 2 'reset'; `upd` = the reference
// would call td3_update here (robot.py:480-483).
__device__ __forceinline__ int action_type_env(int32_t* __restrict__ num_episodes, uint8_t* __restrict__ demo_flag, int32_t* __restrict__ plan_index,
                                               int32_t* __restrict__ path_length, uint8_t* __restrict__ goal_reached,
                                               uint8_t* __restrict__ stuck_flag, double* __restrict__ noise_scale, int64_t i, bool& upd) {
  int ne = num_episodes[i];
  bool df = demo_flag[i] != 0;
  int type = 0;
  upd = false;
  if (ne <= kNumDemo && !df) { ne += 1; type = 1; }
  if (ne > kNumDemo && !df) { df = true; ne += 1; type = 2; }
  if (plan_index[i] == path_length[i] - 1 || goal_reached[i] || stuck_flag[i]) {
    ne += 1;                                        // Robot.reset  robot.py:492-506
    plan_index[i] = 0;
    goal_reached[i] = 0;
    stuck_flag[i] = 0;
    noise_scale[i] = __dmul_rn(noise_scale[i], 0.75);
    path_length[i] += 20;
    type = 2;
    upd = true;
  } else {
    plan_index[i] += 1;
  }
  num_episodes[i] = ne;
  demo_flag[i] = df ? 1 : 0;
  return type;
}

// All threads of the CTA must call this (block-wide barriers for the demo staging, warp ballots for the compacted push).
// `live`: this thread holds an env that steps in this tick; (sx,sy) pre-step state, (ax,ay) action, (nx,ny) next state.
// Demo points ([m][2] float64, shared by all envs) are swept from shared memory.
template <bool kStagePts>
__device__ __forceinline__ void transition_env(const RobotState& st, const float sxi, const float syi, const float axi, const float ayi,
                                               const float nxi, const float nyi, const bool live, const int64_t i, const int64_t n,
                                               const double* __restrict__ demo, const int32_t* __restrict__ cell_start /*nullable*/, const int64_t m,
                                               float* __restrict__ reward_out, double* __restrict__ reward64, uint8_t* __restrict__ done_out,
                                               const ReplayRing& ring, const bool masked_push) {
  __shared__ double2 tile[512];
  __shared__ int32_t s_cell[kDemoCells + 1];
  const int64_t ii = live ? i : 0;
  const double px = (double)nxi, py = (double)nyi;
  // compute_reward([next_state])  robot.py:741-762
  const double gd = norm2_np(__dsub_rn(px, st.goal[ii]), __dsub_rn(py, st.goal[n + ii]));
  const bool reached = (-gd >= -kGoalRadius);
  // nearest demonstration state (only needed when the goal was not reached and there are demos): min over m points
  double best = INFINITY;
  if (m > 0 && cell_start) {
    // Exact search on a two-level uniform grid, warp-cooperative.  The points are sorted by fine cell
    // (kDemoGrid x kDemoGrid cells of side kDemoCell; points outside the grid sit in the nearest border cell, whose box is
    // therefore open on its outer sides).  A warp serves the queries of its 32 envs one after the other, all lanes on the
    // same query (a per-thread traversal diverges into 32 serial walks: measured 2x SLOWER than the full sweep for queries
    // far from the demonstrations): a strided subsample of the points gives an upper bound; the 64 blocks of 4 x 4 cells are
    // tested two per lane, the 16 cells of a surviving block one per lane, and the points of a surviving cell are evaluated
    // 32 at a time (coalesced 16 B loads).  A box is skipped when the distance from the query to it already exceeds the best
    // distance found (1e-9 relative margin, far above the rounding of the three operations).  Every evaluated candidate goes
    // through the same three float64 operations as the full sweep, and the minimum over any subset that contains the true
    // nearest point is the same number, so the result is bit-identical to the sweep.
    // With kStagePts the sorted points (16 B each, up to kDemoStageMax of them) are first copied to shared memory: the
    // walk is a chain of short dependent loads, and from L2 their latency (not the arithmetic) was the whole cost - 230 us
    // for 65 536 queries against 11 355 points, the same as the per-thread walk.
    extern __shared__ __align__(16) unsigned char s_dyn[];
    double2* s_pts = reinterpret_cast<double2*>(s_dyn);
    for (int k = threadIdx.x; k <= kDemoCells; k += blockDim.x) s_cell[k] = cell_start[k];
    if (kStagePts)
      for (int64_t k = threadIdx.x; k < m; k += blockDim.x) s_pts[k] = __ldg(reinterpret_cast<const double2*>(demo) + k);
    __syncthreads();
    {
      const double2* gpts = reinterpret_cast<const double2*>(demo);
      const int lane = threadIdx.x & 31;
      constexpr double kKeep = 1.0 - 1e-9;
      constexpr int kB = kDemoGrid / 4;                              // blocks per side
      auto gap = [](double p, int c, int cells) {                    // distance from p to the slab of cells [c, c + cells), open at the grid border
        const double lo = c == 0 ? -INFINITY : (double)c * kDemoCell;
        const double hi = c + cells >= kDemoGrid ? INFINITY : (double)(c + cells) * kDemoCell;
        return fmax(fmax(lo - p, p - hi), 0.0);
      };
      auto warp_min = [](double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
      };
      uint32_t todo = __ballot_sync(0xffffffffu, live && !reached);
      const int64_t stride = max((int64_t)1, m / 128);
      while (todo) {
        const int q = __ffs(todo) - 1;
        todo &= todo - 1;
        const double qx = __shfl_sync(0xffffffffu, px, q), qy = __shfl_sync(0xffffffffu, py, q);
        double pm = INFINITY;                                        // this lane's partial minimum for query q
        auto eval = [&](int64_t k) {
          const double2 t = kStagePts ? s_pts[k] : __ldg(gpts + k);
          const double dx = qx - t.x, dy = qy - t.y;
          pm = fmin(pm, fma(dy, dy, dx * dx));
        };
        for (int64_t k = (int64_t)lane * stride; k < m; k += 32 * stride) eval(k);
        double wb = warp_min(pm);                                    // upper bound, uniform over the warp
        // block level: lane tests blocks `lane` and `lane + 32`
        uint32_t bmask[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int b = lane + 32 * h, X = (b / kB) * 4, Y = (b % kB) * 4;
          bool any = false;
#pragma unroll
          for (int gx = 0; gx < 4; ++gx) any |= s_cell[(X + gx) * kDemoGrid + Y + 4] != s_cell[(X + gx) * kDemoGrid + Y];
          const double bx = gap(qx, X, 4), by = gap(qy, Y, 4);
          bmask[h] = __ballot_sync(0xffffffffu, any && (bx * bx + by * by) * kKeep <= wb);
        }
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          uint32_t bm = bmask[h];
          while (bm) {
            const int b = __ffs(bm) - 1 + 32 * h;
            bm &= bm - 1;
            const int X = (b / kB) * 4, Y = (b % kB) * 4;
            const double bx = gap(qx, X, 4), by = gap(qy, Y, 4);
            if ((bx * bx + by * by) * kKeep > wb) continue;          // the bound has tightened since the block test (uniform)
            // cell level: lanes 0-15 own the 16 cells of the block
            const int gx = X + ((lane & 15) >> 2), gy = Y + (lane & 3);
            const int k0 = s_cell[gx * kDemoGrid + gy], k1 = s_cell[gx * kDemoGrid + gy + 1];
            const double fx = gap(qx, gx, 1), fy = gap(qy, gy, 1);
            uint32_t cm = __ballot_sync(0xffffffffu, lane < 16 && k1 > k0 && (fx * fx + fy * fy) * kKeep <= wb);
            while (cm) {
              const int c = __ffs(cm) - 1;
              cm &= cm - 1;
              const int c0 = __shfl_sync(0xffffffffu, k0, c), c1 = __shfl_sync(0xffffffffu, k1, c);
              for (int k = c0 + lane; k < c1; k += 32) eval(k);
            }
            wb = warp_min(pm);
          }
        }
        if (lane == q) best = wb;
      }
    }
  } else if (m > 0) {
    for (int64_t base = 0; base < m; base += 512) {
      const int cnt = (int)min((int64_t)512, m - base);
      __syncthreads();
      for (int k = threadIdx.x; k < cnt; k += blockDim.x) tile[k] = reinterpret_cast<const double2*>(demo)[base + k];
      __syncthreads();
#pragma unroll 4
      for (int k = 0; k < cnt; ++k) {
        const double dx = px - tile[k].x, dy = py - tile[k].y;
        best = fmin(best, fma(dy, dy, dx * dx));     // squared distance; sqrt once at the end (monotone)
      }
    }
  }
  double reward = 0.0;
  bool done = false;
  if (live) {
    if (reached) {
      st.goal_reached[i] = 1;
      reward = kGoalReward;
    } else if (m == 0) {
      reward = -gd;
    } else {
      const double prox = st.demo_flag[i] ? -sqrt(best) : 0.0;
      reward = __dadd_rn(-gd, __dmul_rn(10.0, prox));   // DEMO_PROXIMITY_FACTOR, robot.py:43, 760
    }
    // check_if_stuck(state)  robot.py:509-538, on the pre-step state
    const double cxs = (double)sxi, cys = (double)syi;
    int cnt = st.hist_count[i], head = st.hist_head[i];
    bool stuck = false;
    if (cnt >= kStuckSteps) {
      stuck = true;
      for (int k = 0; k < kStuckSteps; ++k) {
        const double hx = (double)st.hist[(int64_t)(2 * k) * n + i], hy = (double)st.hist[(int64_t)(2 * k + 1) * n + i];
        if (!(norm2_np(__dsub_rn(cxs, hx), __dsub_rn(cys, hy)) < kStuckThreshold)) stuck = false;
      }
      if (stuck) { cnt = 0; head = 0; }                 // previous_states.clear()
      else { head = (head + 1) % kStuckSteps; cnt -= 1; }   // pop(0)
    }
    const int slot = (head + cnt) % kStuckSteps;        // append(state)
    st.hist[(int64_t)(2 * slot) * n + i] = sxi;
    st.hist[(int64_t)(2 * slot + 1) * n + i] = syi;
    st.hist_count[i] = cnt + 1;
    st.hist_head[i] = head;
    if (stuck) {
      st.stuck_flag[i] = 1;
      reward = __dsub_rn(reward, kStuckPenalty);
    }
    done = st.plan_index[i] == st.path_length[i] - 1;   // robot.py:672: time-out only
    reward_out[i] = (float)reward;
    if (reward64) reward64[i] = reward;
    done_out[i] = done ? 1 : 0;
  }
  if (ring.s) {                                         // memory.push  robot.py:675
    int64_t p;
    if (masked_push) {
      // only some envs push: compact them with a warp ballot, one atomic per warp on the ring's row counter
      const uint32_t act = __ballot_sync(0xffffffffu, live);
      const int lane = threadIdx.x & 31;
      unsigned long long base = 0;
      if (lane == 0 && act) base = atomicAdd(ring.total, (unsigned long long)__popc(act));
      base = __shfl_sync(0xffffffffu, base, 0);
      p = (int64_t)((base + __popc(act & ((1u << lane) - 1u))) % (unsigned long long)ring.capacity);
    } else {
      p = (ring.position + i) % ring.capacity;
      if (i == 0 && ring.total) atomicAdd(ring.total, (unsigned long long)n);
    }
    if (live) {
      ring.s[p] = make_float2(sxi, syi);
      ring.a[p] = make_float2(axi, ayi);
      ring.r[p] = (float)reward;
      ring.s2[p] = make_float2(nxi, nyi);
      ring.notdone[p] = done ? 0.f : 1.f;
    }
  }
}

}  // namespace rtd3
