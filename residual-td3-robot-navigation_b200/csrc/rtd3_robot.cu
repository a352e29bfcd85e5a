// Robot per-step hooks batched over n envs.  Reference behaviour: /root/reference/robot.py
//   get_next_action_training / testing, residual_action, generate_noise      robot.py:541-642
//   process_transition, compute_reward, check_if_stuck                      robot.py:645-675, 727-762, 509-538
//   get_next_action_type, reset                                             robot.py:443-506
// The actor forward inside the act hook is rtd3_mlp_forward (rtd3_td3.cu); the kernels here are the glue around it.
// Threshold compares (goal radius 5, stuck radius 2) are evaluated in float64 from the float32 states and the
// float64 goal with numpy's norm rounding (rtd3_mt.cuh: norm2_np), so the flags match the reference bit for bit.
#include "rtd3_common.cuh"
#include "rtd3_mt.cuh"
#include "rtd3_robot.cuh"

namespace rtd3 {

// baseline_action = state - goal (robot.py:556 / 586), cast to float32 as torch.FloatTensor does (robot.py:612)
__global__ void robot_baseline_kernel(const float* __restrict__ x, const float* __restrict__ y, const double* __restrict__ goal,
                                      float* __restrict__ base /*[n][2]*/, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  reinterpret_cast<float2*>(base)[i] = baseline_env(x[i], y[i], goal[i], goal[n + i]);
}

// action = clip(baseline + residual + noise, +-5)   (robot.py:560-567 / 590-593); noise = unit_normal * noise_scale * 5 (robot.py:640)
__global__ void robot_compose_kernel(const float* __restrict__ x, const float* __restrict__ y, const double* __restrict__ goal,
                                     const float* __restrict__ residual /*[n][2]*/, const double* __restrict__ unit_noise /*nullable [2][n]*/,
                                     const double* __restrict__ noise_scale /*[n]*/, const int8_t* __restrict__ type /*nullable [n]*/,
                                     float* __restrict__ ax, float* __restrict__ ay, double* __restrict__ action64 /*nullable [2][n]*/,
                                     int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (type && type[i] != 0) {            // this env does not step in this tick (robot-learning.py:82-95): a null action
    ax[i] = 0.f; ay[i] = 0.f;
    if (action64) { action64[i] = 0.0; action64[n + i] = 0.0; }
    return;
  }
  double cx, cy;
  compose_env(x[i], y[i], goal[i], goal[n + i], reinterpret_cast<const float2*>(residual)[i], unit_noise != nullptr,
              unit_noise ? unit_noise[i] : 0.0, unit_noise ? unit_noise[n + i] : 0.0, unit_noise ? noise_scale[i] : 0.0, cx, cy);
  ax[i] = (float)cx;
  ay[i] = (float)cy;
  if (action64) { action64[i] = cx; action64[n + i] = cy; }
}

// process_transition for n envs; demo points ([m][2] float64, shared by all envs): candidate lists or a sweep from shared memory.
__global__ void __launch_bounds__(256)
robot_transition_kernel(RobotState st, const float* __restrict__ sx, const float* __restrict__ sy, const float* __restrict__ ax,
                        const float* __restrict__ ay, const float* __restrict__ nx, const float* __restrict__ ny,
                        const double* __restrict__ demo, const int32_t* __restrict__ list_start /*nullable*/,
                        const double* __restrict__ list_pts, int64_t m, float* __restrict__ reward_out, double* __restrict__ reward64,
                        uint8_t* __restrict__ done_out, ReplayRing ring, const int8_t* __restrict__ type /*nullable: only type 0 steps*/,
                        int64_t n) {
  static_assert(kEnvCells == RTD3_ENV_DEMO_CELLS, "grid size of the per-env demonstration sets");
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n && (!type || type[i] == 0);
  const int64_t ii = live ? i : 0;
  transition_env(st, sx[ii], sy[ii], ax[ii], ay[ii], nx[ii], ny[ii], live, i, n, demo, list_start, list_pts, m, reward_out, reward64, done_out,
                 ring, type != nullptr);
}

// Candidate lists of the nearest-demonstration search (see nearest_demo_sq, rtd3_robot.cuh): one CTA per 1 x 1 cell.
//   phase A: the states nearest to the cell's four corners and to its centre (p*_0..4);
//   phase B: state p is kept iff, for EVERY p*_a, p is at least as close as p*_a at SOME corner (1e-9 relative slack).
// kFill = false counts the survivors per cell, kFill = true writes them to out[start[cell] ...] (any order: the query takes a min).
template <bool kFill>
__global__ void __launch_bounds__(256) demo_lists_kernel(const double2* __restrict__ pts, int m, int32_t* __restrict__ counts,
                                                         const int32_t* __restrict__ start, double2* __restrict__ out) {
  __shared__ double s_bd[8][5];
  __shared__ int s_bi[8][5];
  __shared__ double s_thr[5][4];
  __shared__ int s_count;
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double x0 = (double)(c / kDemoGrid) * kDemoCell, y0 = (double)(c % kDemoGrid) * kDemoCell;
  const double ax[5] = {x0, x0, x0 + kDemoCell, x0 + kDemoCell, x0 + 0.5 * kDemoCell};
  const double ay[5] = {y0, y0 + kDemoCell, y0, y0 + kDemoCell, y0 + 0.5 * kDemoCell};
  auto dist2 = [](double qx, double qy, double2 p) { const double dx = qx - p.x, dy = qy - p.y; return fma(dy, dy, dx * dx); };
  double bd[5];
  int bi[5];
#pragma unroll
  for (int a = 0; a < 5; ++a) { bd[a] = INFINITY; bi[a] = 0; }
  for (int k = tid; k < m; k += 256) {
    const double2 p = __ldg(pts + k);
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      const double d = dist2(ax[a], ay[a], p);
      if (d < bd[a]) { bd[a] = d; bi[a] = k; }
    }
  }
#pragma unroll
  for (int a = 0; a < 5; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, bd[a], o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi[a], o);
      if (od < bd[a]) { bd[a] = od; bi[a] = oi; }
    }
    if (lane == 0) { s_bd[warp][a] = bd[a]; s_bi[warp][a] = bi[a]; }
  }
  if (tid == 0) s_count = 0;
  __syncthreads();
  if (tid < 20) {                                                    // thread (a, corner): |corner - p*_a|^2 with the slack
    const int a = tid >> 2, j = tid & 3;
    double d = s_bd[0][a];
    int k = s_bi[0][a];
    for (int w = 1; w < 8; ++w)
      if (s_bd[w][a] < d) { d = s_bd[w][a]; k = s_bi[w][a]; }
    s_thr[a][j] = dist2(ax[j], ay[j], __ldg(pts + k)) * kDemoKeep;
  }
  __syncthreads();
  int mine = 0;
  for (int base = 0; base < m; base += 256) {
    const int k = base + tid;
    bool keep = false;
    double2 p = make_double2(0.0, 0.0);
    if (k < m) {
      p = __ldg(pts + k);
      const double d0 = dist2(ax[0], ay[0], p), d1 = dist2(ax[1], ay[1], p), d2 = dist2(ax[2], ay[2], p), d3 = dist2(ax[3], ay[3], p);
      keep = true;
#pragma unroll
      for (int a = 0; a < 5; ++a) keep = keep && (d0 <= s_thr[a][0] || d1 <= s_thr[a][1] || d2 <= s_thr[a][2] || d3 <= s_thr[a][3]);
    }
    const uint32_t kept = __ballot_sync(0xffffffffu, keep);
    if (kFill) {
      int wbase = 0;
      if (lane == 0 && kept) wbase = atomicAdd(&s_count, __popc(kept));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (keep) out[(int64_t)start[c] + wbase + __popc(kept & ((1u << lane) - 1u))] = p;
    } else if (lane == 0) {
      mine += __popc(kept);
    }
  }
  if (!kFill) {
    if (lane == 0 && mine) atomicAdd(&s_count, mine);
    __syncthreads();
    if (tid == 0) counts[c] = s_count;
  }
}

// get_next_action_type + reset  (robot.py:443-506).  type: 0 'step', 1 'demo', 2 'reset'; update[i] = 1 where the
// reference would call td3_update (robot.py:480-483); any_update accumulates how many envs did.
__global__ void robot_action_type_kernel(int32_t* __restrict__ num_episodes, uint8_t* __restrict__ demo_flag, int32_t* __restrict__ plan_index,
                                         int32_t* __restrict__ path_length, uint8_t* __restrict__ goal_reached, uint8_t* __restrict__ stuck_flag,
                                         double* __restrict__ noise_scale, int8_t* __restrict__ type_out, uint8_t* __restrict__ update_out,
                                         int32_t* __restrict__ any_update, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool upd = false;
  if (i < n) {
    type_out[i] = (int8_t)action_type_env(num_episodes, demo_flag, plan_index, path_length, goal_reached, stuck_flag, noise_scale, i, upd);
    update_out[i] = upd ? 1 : 0;
  }
  const uint32_t ended = __ballot_sync(0xffffffffu, upd);                                   // warp-ballot of the episode-end flags
  if (ended && (threadIdx.x & 31) == 0) atomicAdd(any_update, __popc(ended));               // one atomic per warp that needs it
}

}  // namespace rtd3

using namespace rtd3;

__global__ void trainer_tally_kernel(const int8_t* __restrict__ type, int64_t* __restrict__ steps, int64_t* __restrict__ resets, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int8_t t = type[i];
  if (t == 0) steps[i] += 1;
  else if (t == 2) resets[i] += 1;
}

extern "C" {

int32_t rtd3_trainer_tally(const int8_t* type, int64_t* steps, int64_t* resets, int64_t n, void* stream) {
  RTD3_CHECK_ARG(type && steps && resets && n >= 0, "bad argument");
  if (n == 0) return 0;
  trainer_tally_kernel<<<(int)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(type, steps, resets, n);
  RTD3_LAUNCHED();
  return 0;
}

int32_t rtd3_robot_baseline(const float* x, const float* y, const double* goal, float* base, int64_t n, void* stream) {
  RTD3_CHECK_ARG(x && y && goal && base && n >= 0, "bad argument");
  if (n == 0) return 0;
  robot_baseline_kernel<<<(int)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, goal, base, n);
  RTD3_LAUNCHED();
  return 0;
}

int32_t rtd3_robot_compose_action(const float* x, const float* y, const double* goal, const float* residual, const double* unit_noise,
                                  const double* noise_scale, const int8_t* type, float* ax, float* ay, double* action64, int64_t n,
                                  void* stream) {
  RTD3_CHECK_ARG(x && y && goal && residual && ax && ay && n >= 0, "bad argument");
  RTD3_CHECK_ARG(!unit_noise || noise_scale, "noise needs noise_scale");
  if (n == 0) return 0;
  robot_compose_kernel<<<(int)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, goal, residual, unit_noise, noise_scale, type, ax, ay, action64, n);
  RTD3_LAUNCHED();
  return 0;
}

int32_t rtd3_demo_lists(const double* demo, int64_t num_demo, int32_t* counts, const int32_t* start, double* list, void* stream) {
  RTD3_CHECK_ARG(demo && num_demo >= 1 && num_demo < (1ll << 31), "bad demonstration set");
  RTD3_CHECK_ARG((list == nullptr) ? (counts != nullptr) : (start != nullptr), "count pass needs counts, fill pass needs start and list");
  if (!list) demo_lists_kernel<false><<<kDemoCells, 256, 0, (cudaStream_t)stream>>>((const double2*)demo, (int)num_demo, counts, nullptr, nullptr);
  else demo_lists_kernel<true><<<kDemoCells, 256, 0, (cudaStream_t)stream>>>((const double2*)demo, (int)num_demo, nullptr, start, (double2*)list);
  RTD3_LAUNCHED();
  return 0;
}

int32_t rtd3_robot_transition(const double* goal, float* hist, int32_t* hist_count, int32_t* hist_head, uint8_t* goal_reached,
                              uint8_t* stuck_flag, const uint8_t* demo_flag, const int32_t* plan_index, const int32_t* path_length,
                              const float* sx, const float* sy, const float* ax, const float* ay, const float* nx, const float* ny,
                              const double* demo, const int32_t* demo_list_start, const double* demo_list, int64_t num_demo, float* reward,
                              double* reward64, uint8_t* done, float* rp_s, float* rp_a, float* rp_r, float* rp_s2, float* rp_notdone,
                              int64_t capacity, int64_t position, uint64_t* rp_total, const int8_t* type, const double* env_demo_pts,
                              const int32_t* env_demo_cells, const int32_t* env_demo_count, int64_t env_demo_cap, int64_t n, void* stream) {
  RTD3_CHECK_ARG(goal && hist && hist_count && hist_head && goal_reached && stuck_flag && demo_flag && plan_index && path_length,
                 "null robot state");
  RTD3_CHECK_ARG(!env_demo_pts || (env_demo_cells && env_demo_count && env_demo_cap > 0), "incomplete per-env demonstration sets");
  RTD3_CHECK_ARG(sx && sy && ax && ay && nx && ny && reward && done, "null transition array");
  RTD3_CHECK_ARG(num_demo == 0 || demo, "demo set missing");
  RTD3_CHECK_ARG((demo_list_start == nullptr) == (demo_list == nullptr), "demo_list_start and demo_list go together");
  RTD3_CHECK_ARG(n >= 0, "negative n");
  RTD3_CHECK_ARG(!rp_s || (rp_a && rp_r && rp_s2 && rp_notdone && capacity > 0 && position >= 0 && position < capacity && n <= capacity),
                 "bad replay ring");
  if (n == 0) return 0;
  RobotState st{goal, hist, hist_count, hist_head, goal_reached, stuck_flag, demo_flag, plan_index, path_length,
                env_demo_pts, env_demo_cells, env_demo_count, env_demo_cap};
  RTD3_CHECK_ARG(!(type && rp_s) || rp_total, "a masked push needs the ring's device row counter");
  ReplayRing ring{(float2*)rp_s, (float2*)rp_a, rp_r, (float2*)rp_s2, rp_notdone, capacity, position, (unsigned long long*)rp_total};
  const int block = n <= 148 * 256 ? 128 : 256;      // a small batch spread over more SMs
  robot_transition_kernel<<<(int)ceil_div(n, block), block, 0, (cudaStream_t)stream>>>(st, sx, sy, ax, ay, nx, ny, demo, demo_list_start, demo_list,
                                                                                      num_demo, reward, reward64, done, ring, type, n);
  RTD3_LAUNCHED();
  return 0;
}

int32_t rtd3_robot_next_action_type(int32_t* num_episodes, uint8_t* demo_flag, int32_t* plan_index, int32_t* path_length,
                                    uint8_t* goal_reached, uint8_t* stuck_flag, double* noise_scale, int8_t* type_out, uint8_t* update_out,
                                    int32_t* any_update, int64_t n, void* stream) {
  RTD3_CHECK_ARG(num_episodes && demo_flag && plan_index && path_length && goal_reached && stuck_flag && noise_scale && type_out &&
                     update_out && any_update && n >= 0,
                 "bad argument");
  if (n == 0) return 0;
  robot_action_type_kernel<<<(int)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(num_episodes, demo_flag, plan_index, path_length,
                                                                                   goal_reached, stuck_flag, noise_scale, type_out, update_out,
                                                                                   any_update, n);
  RTD3_LAUNCHED();
  return 0;
}

}  // extern "C"
