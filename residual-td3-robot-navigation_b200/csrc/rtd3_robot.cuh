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
  // per-env demonstration sets (batched process_demonstration; all NULL / 0: the shared set passed next to this struct is used)
  const double* env_pts;        // [n][env_cap][2] float64, every env's states sorted by grid cell (demo_grid_build_kernel)
  const int32_t* env_cells;     // [n][kEnvCells + 1] start offsets of the cells in env_pts
  const int32_t* env_count;     // [n] states held
  int64_t env_cap;
};

// Per-env demonstration sets (each env of a batch has bought its own demonstrations, as each reference run does): a uniform
// 25 x 25 grid of 4 x 4 cells per env, states outside the world filed under the nearest border cell.  The query scans the rings of
// cells around its own cell until the best squared distance found is no larger than the squared distance to the border of the
// scanned square (no state outside it can be nearer) - the same three float64 operations per candidate as the full sweep, and the
// minimum over a set that contains the true nearest state: bit-identical to robot.py:753's cdist(...).min().
constexpr int kEnvGrid = 25;
constexpr double kEnvCell = 4.0;
constexpr int kEnvCells = kEnvGrid * kEnvGrid;
__device__ __forceinline__ int env_cell_coord(double v) {
  const int c = (int)floor(v / kEnvCell);
  return c < 0 ? 0 : (c > kEnvGrid - 1 ? kEnvGrid - 1 : c);
}
__device__ __forceinline__ double env_scan_cell(const double2* __restrict__ pts, const int32_t* __restrict__ cells, int gx, int gy, double px,
                                                double py, double best) {
  const int c = gx * kEnvGrid + gy;
  const int k1 = cells[c + 1];
  for (int k = cells[c]; k < k1; ++k) {
    const double2 t = pts[k];
    const double dx = px - t.x, dy = py - t.y;
    best = fmin(best, fma(dy, dy, dx * dx));
  }
  return best;
}
__device__ __forceinline__ double nearest_demo_env_sq(double px, double py, const double2* __restrict__ pts, const int32_t* __restrict__ cells) {
  const int cx = env_cell_coord(px), cy = env_cell_coord(py);
  double best = INFINITY;
  for (int r = 0; r < kEnvGrid; ++r) {
    const int x0 = cx - r, x1 = cx + r, y0 = cy - r, y1 = cy + r;
    for (int gx = max(x0, 0); gx <= min(x1, kEnvGrid - 1); ++gx) {
      if (gx == x0 || gx == x1) {
        for (int gy = max(y0, 0); gy <= min(y1, kEnvGrid - 1); ++gy) best = env_scan_cell(pts, cells, gx, gy, px, py, best);
      } else {
        if (y0 >= 0) best = env_scan_cell(pts, cells, gx, y0, px, py, best);
        if (y1 <= kEnvGrid - 1 && y1 != y0) best = env_scan_cell(pts, cells, gx, y1, px, py, best);
      }
    }
    // states not scanned yet lie beyond the border of the scanned square on a side where cells remain
    double bound = INFINITY;
    if (x0 > 0) bound = fmin(bound, px - (double)x0 * kEnvCell);
    if (x1 < kEnvGrid - 1) bound = fmin(bound, (double)(x1 + 1) * kEnvCell - px);
    if (y0 > 0) bound = fmin(bound, py - (double)y0 * kEnvCell);
    if (y1 < kEnvGrid - 1) bound = fmin(bound, (double)(y1 + 1) * kEnvCell - py);
    if (bound == INFINITY || best <= bound * bound * (1.0 - 1e-12)) break;
  }
  return best;
}

struct ReplayRing {
  float2* s; float2* a; float* r; float2* s2; float* notdone;
  int64_t capacity, position;
  unsigned long long* total;   // device counter of rows ever pushed (nullable); authoritative for masked pushes
};

constexpr int kDemoGrid = RTD3_DEMO_GRID;             // cells per side of the candidate-list grid (the world's own 100 x 100 cells)
constexpr double kDemoCell = RTD3_DEMO_CELL;           // cell side 1: cell edges are exact in float64
constexpr int kDemoCells = kDemoGrid * kDemoGrid;
constexpr double kDemoKeep = 1.0 + 1e-9;               // slack of the candidate test, far above the rounding of a squared distance

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

// Nearest demonstration state (squared distance, float64) for a query inside the world.
// Candidate lists (built by demo_lists_kernel, rtd3_robot.cu): for every 1 x 1 cell of the world the
// demonstration states that can be the nearest one for SOME point of the cell.  A state p cannot be nearest anywhere in the
// cell if another state p* is closer at all four corners: |q-p|^2 - |q-p*|^2 is linear in q, so if it is positive at the corners
// it is positive on the whole (convex) cell.  The lists keep what survives that test against five p* (the states nearest to the
// four corners and to the centre); for the reference's 11 355 states after three demonstrations the mean list holds 5 states
// (median 1, maximum 79).  The query evaluates every candidate with the same three float64 operations as the full sweep, and the
// minimum over any subset that contains the true nearest state is the same number: bit-identical to robot.py:753's cdist min.
__device__ __forceinline__ double nearest_demo_sq(const double px, const double py, const double* __restrict__ demo, const int64_t m,
                                                  const int32_t* __restrict__ list_start, const double* __restrict__ list_pts) {
  double best = INFINITY;
  const double2* pts = reinterpret_cast<const double2*>(demo);
  int64_t k0 = 0, k1 = m;
  if (list_start && px >= 0.0 && px < (double)kDemoGrid * kDemoCell && py >= 0.0 && py < (double)kDemoGrid * kDemoCell) {
    const int c = (int)(px / kDemoCell) * kDemoGrid + (int)(py / kDemoCell);
    k0 = list_start[c];
    k1 = list_start[c + 1];
    pts = reinterpret_cast<const double2*>(list_pts);
  }                                                                  // a query outside the grid sweeps the whole set
  for (int64_t k = k0; k < k1; ++k) {
    const double2 t = __ldg(pts + k);
    const double dx = px - t.x, dy = py - t.y;
    best = fmin(best, fma(dy, dy, dx * dx));
  }
  return best;
}

// Whole warps must call this (warp ballots for the compacted push); with kAllowSweep all threads of the CTA (block-wide barriers
// of the demo sweep).
// `live`: this thread holds an env that steps in this tick; (sx,sy) pre-step state, (ax,ay) action, (nx,ny) next state.
// Demo points ([m][2] float64, shared by all envs): candidate lists when `list_start` is given, else swept from shared memory.
template <bool kAllowSweep = true>
__device__ __forceinline__ void transition_env(const RobotState& st, const float sxi, const float syi, const float axi, const float ayi,
                                               const float nxi, const float nyi, const bool live, const int64_t i, const int64_t n,
                                               const double* __restrict__ demo, const int32_t* __restrict__ list_start /*nullable*/,
                                               const double* __restrict__ list_pts, const int64_t m,
                                               float* __restrict__ reward_out, double* __restrict__ reward64, uint8_t* __restrict__ done_out,
                                               const ReplayRing& ring, const bool masked_push) {
  const int64_t ii = live ? i : 0;
  const double px = (double)nxi, py = (double)nyi;
  // Every load this env needs is issued here, before the first dependent use and before any store (the compiler keeps loads
  // behind earlier stores through pointers it cannot prove distinct): the kernel is a chain of L2 round trips (ncu: long-scoreboard
  // stalls on the stuck ring, the time-out test and the flags, one after the other), and this makes it one.
  const double goal_x = st.goal[ii], goal_y = st.goal[n + ii];
  const bool demo_phase_over = st.demo_flag[ii] != 0;
  int cnt = st.hist_count[ii], head = st.hist_head[ii];
  float hx[kStuckSteps], hy[kStuckSteps];
#pragma unroll
  for (int k = 0; k < kStuckSteps; ++k) { hx[k] = st.hist[(int64_t)(2 * k) * n + ii]; hy[k] = st.hist[(int64_t)(2 * k + 1) * n + ii]; }
  const bool timeout = st.plan_index[ii] == st.path_length[ii] - 1;
  // memory.push (robot.py:675): the row's slot.  With a masked push the stepping envs are compacted by a warp ballot and one
  // atomic per warp on the ring's row counter - issued now, consumed at the end.
  int64_t p = 0;
  if (ring.s) {
    if (masked_push) {
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
  }
  // compute_reward([next_state])  robot.py:741-762
  const double gd = norm2_np(__dsub_rn(px, goal_x), __dsub_rn(py, goal_y));
  const bool reached = (-gd >= -kGoalRadius);
  // nearest demonstration state (only needed when the goal was not reached, the demo phase is over and there are demos)
  double best = INFINITY;
  int64_t m_env = m;                                    // demonstration states this env's reward looks at
  if (st.env_pts) {
    m_env = live ? (int64_t)st.env_count[i] : 0;
    if (live && !reached && demo_phase_over && m_env > 0)
      best = nearest_demo_env_sq(px, py, reinterpret_cast<const double2*>(st.env_pts) + (int64_t)i * st.env_cap,
                                 st.env_cells + (int64_t)i * (kEnvCells + 1));
  } else if (m > 0 && list_start) {
    if (live && !reached && demo_phase_over) best = nearest_demo_sq(px, py, demo, m, list_start, list_pts);
  } else if (kAllowSweep && m > 0) {
    // (callers that cannot take a block-wide barrier here - a subset of the CTA's warps - instantiate kAllowSweep = false and
    // guarantee lists or m == 0)
    __shared__ double2 tile[512];
    double b0 = INFINITY, b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;      // four independent minimum chains
    for (int64_t base = 0; base < m; base += 512) {
      const int cnt_t = (int)min((int64_t)512, m - base);
      __syncthreads();
      for (int k = threadIdx.x; k < cnt_t; k += blockDim.x) tile[k] = reinterpret_cast<const double2*>(demo)[base + k];
      __syncthreads();
      int k = 0;
      for (; k + 4 <= cnt_t; k += 4) {
        const double dx0 = px - tile[k].x, dy0 = py - tile[k].y, dx1 = px - tile[k + 1].x, dy1 = py - tile[k + 1].y;
        const double dx2 = px - tile[k + 2].x, dy2 = py - tile[k + 2].y, dx3 = px - tile[k + 3].x, dy3 = py - tile[k + 3].y;
        b0 = fmin(b0, fma(dy0, dy0, dx0 * dx0));     // squared distance; sqrt once at the end (monotone)
        b1 = fmin(b1, fma(dy1, dy1, dx1 * dx1));
        b2 = fmin(b2, fma(dy2, dy2, dx2 * dx2));
        b3 = fmin(b3, fma(dy3, dy3, dx3 * dx3));
      }
      for (; k < cnt_t; ++k) {
        const double dx = px - tile[k].x, dy = py - tile[k].y;
        b0 = fmin(b0, fma(dy, dy, dx * dx));
      }
    }
    best = fmin(fmin(b0, b1), fmin(b2, b3));
  }
  double reward = 0.0;
  bool done = false;
  if (live) {
    if (reached) {
      st.goal_reached[i] = 1;
      reward = kGoalReward;
    } else if (m_env == 0) {
      reward = -gd;
    } else {
      const double prox = demo_phase_over ? -sqrt(best) : 0.0;
      reward = __dadd_rn(-gd, __dmul_rn(10.0, prox));   // DEMO_PROXIMITY_FACTOR, robot.py:43, 760
    }
    // check_if_stuck(state)  robot.py:509-538, on the pre-step state
    const double cxs = (double)sxi, cys = (double)syi;
    bool stuck = false;
    if (cnt >= kStuckSteps) {
      stuck = true;
#pragma unroll
      for (int k = 0; k < kStuckSteps; ++k)
        if (!(norm2_np(__dsub_rn(cxs, (double)hx[k]), __dsub_rn(cys, (double)hy[k])) < kStuckThreshold)) stuck = false;
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
    done = timeout;                                     // robot.py:672: time-out only
    reward_out[i] = (float)reward;
    if (reward64) reward64[i] = reward;
    done_out[i] = done ? 1 : 0;
    if (ring.s) {
      ring.s[p] = make_float2(sxi, syi);
      ring.a[p] = make_float2(axi, ayi);
      ring.r[p] = (float)reward;
      ring.s2[p] = make_float2(nxi, nyi);
      ring.notdone[p] = done ? 0.f : 1.f;
    }
  }
}

}  // namespace rtd3
