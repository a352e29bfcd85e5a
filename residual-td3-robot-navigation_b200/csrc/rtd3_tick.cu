// One tick of the batched driver loop (robot-learning.py:66-101, training branch) in TWO launches around the actor forward:
//   rtd3_tick_pre  : get_next_action_type + reset state machine (robot.py:443-506), actor input = state - goal (robot.py:556)
//   [actor forward : rtd3_mlp_forward / rtd3_mlp_forward_tf32]
//   rtd3_tick_post : action = clip(baseline + residual + noise) (robot.py:560-567), Environment.step (environment.py:98-127),
//                    process_transition incl. reward / stuck / done / replay push (robot.py:645-675), the money counters
//                    (robot-learning.py:78, 86, 99) and Environment.reset for the envs whose tick is a 'reset' (environment.py:130-137)
// The same device functions as the one-hook-per-launch entry points (rtd3_robot.cuh, rtd3_env_step.cuh) are used, in the same
// order per env, so a fused tick leaves every array bit-identical to the sequence
//   next_action_type, baseline, forward, compose_action, env_step, robot_transition, env_reset(mask = type 2), trainer_tally
// (tests/test_tick_gpu.py).  At 8 192 envs per GPU that sequence is ten dependent launches of 3-6 us each around a 15 us forward:
// launch latency, not arithmetic, was the tick.
#include "rtd3_common.cuh"
#include "rtd3_env_step.cuh"
#include "rtd3_mt.cuh"
#include "rtd3_robot.cuh"

namespace rtd3 {

// Two unit normals for (env, tick) from Philox4x32-10 + Box-Muller in float64 (throughput-mode exploration noise: counter-based,
// so a replayed CUDA graph draws fresh noise every tick from the device tick counter).
__device__ __forceinline__ void philox_normal2(uint64_t seed, uint64_t tick, uint64_t env, double& z0, double& z1) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)env, (uint32_t)(env >> 32), (uint32_t)tick, (uint32_t)(tick >> 32)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const double u1 = ((double)(r.x >> 5) * 67108864.0 + (double)(r.y >> 6)) / 9007199254740992.0;   // [0,1), 53 bits
  const double u2 = ((double)(r.z >> 5) * 67108864.0 + (double)(r.w >> 6)) / 9007199254740992.0;
  const double rad = sqrt(-2.0 * log(1.0 - u1));                                                  // 1 - u1 in (0,1]
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  z0 = rad * c;
  z1 = rad * s;
}

__global__ void __launch_bounds__(256) tick_pre_kernel(rtd3_tick_state t) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = t.n;
  bool upd = false;
  if (i < n) {
    t.type[i] = (int8_t)action_type_env(t.num_episodes, t.demo_flag, t.plan_index, t.path_length, t.goal_reached, t.stuck_flag,
                                        t.noise_scale, i, upd);
    t.update[i] = upd ? 1 : 0;
    reinterpret_cast<float2*>(t.base)[i] = baseline_env(t.x[i], t.y[i], t.goal[i], t.goal[n + i]);
  }
  const uint32_t ended = __ballot_sync(0xffffffffu, upd);
  if (ended && (threadIdx.x & 31) == 0) atomicAdd(t.any_update, __popc(ended));
  if (i == 0 && t.tick_counter) t.tick_counter[0] += 1ull;
}

__global__ void __launch_bounds__(256, 2) tick_post_kernel(rtd3_tick_state t, const float2* __restrict__ table,
                                                        const float* __restrict__ residual /*[n][2]*/,
                                                        const double* __restrict__ unit_noise /*[2][n], mode 1*/, int noise_mode) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = t.n;
  const bool in = i < n;
  const int64_t ii = in ? i : 0;
  const int type = in ? (int)t.type[ii] : 1;
  const bool live = in && type == 0;
  const float x = t.x[ii], y = t.y[ii];
  // get_next_action_training: envs that do not step in this tick get a null action (robot-learning.py:82-95)
  float ax = 0.f, ay = 0.f;
  if (live) {
    double zx = 0.0, zy = 0.0;
    if (noise_mode == RTD3_TICK_NOISE_GIVEN) { zx = unit_noise[i]; zy = unit_noise[n + i]; }
    else if (noise_mode == RTD3_TICK_NOISE_PHILOX) philox_normal2(t.philox_seed, t.tick_counter[0], (uint64_t)i, zx, zy);
    double cx, cy;
    compose_env(x, y, t.goal[i], t.goal[n + i], reinterpret_cast<const float2*>(residual)[i], noise_mode != RTD3_TICK_NOISE_NONE, zx, zy,
                noise_mode != RTD3_TICK_NOISE_NONE ? t.noise_scale[i] : 0.0, cx, cy);
    ax = (float)cx;
    ay = (float)cy;
  }
  // Environment.step
  float nx, ny;
  step_one<true>(LdgTable{table}, x, y, ax, ay, nx, ny);
  // process_transition (reward, stuck ring, done, compacted replay push)
  const RobotState st{t.goal, t.hist, t.hist_count, t.hist_head, t.goal_reached, t.stuck_flag, t.demo_flag, t.plan_index, t.path_length};
  const ReplayRing ring{(float2*)t.rp_s, (float2*)t.rp_a, t.rp_r, (float2*)t.rp_s2, t.rp_notdone, t.capacity, 0,
                        (unsigned long long*)t.rp_total};
  transition_env(st, x, y, ax, ay, nx, ny, live, i, n, t.demo, t.demo_list_start, t.demo_list, t.num_demo, t.reward, t.reward64, t.done, ring,
                 true);
  if (in) {
    t.ax[i] = ax;
    t.ay[i] = ay;
    if (t.prev_x) { t.prev_x[i] = x; t.prev_y[i] = y; }
    t.x[i] = nx;
    t.y[i] = ny;
    if (type == 0) t.steps_bought[i] += 1;
    else if (type == 2) t.resets_bought[i] += 1;
  }
  // Environment.reset where the tick is a 'reset' (warp-synchronous: wrapping MT19937 streams are twisted by the whole warp)
  reset_env_warp(t.env_bank, t.region, in && type == 2, i, t.x, t.y, t.state64);
}

static int32_t check_state(const rtd3_tick_state* t) {
  RTD3_CHECK_ARG(t, "null tick state");
  RTD3_CHECK_ARG(t->n >= 0, "negative n");
  RTD3_CHECK_ARG(t->x && t->y && t->goal && t->region, "null env array");
  RTD3_CHECK_ARG(t->env_bank.mt && t->env_bank.pos && t->env_bank.n == t->n, "env MT19937 bank missing or of another size");
  RTD3_CHECK_ARG(t->num_episodes && t->demo_flag && t->plan_index && t->path_length && t->goal_reached && t->stuck_flag && t->noise_scale,
                 "null robot episode state");
  RTD3_CHECK_ARG(t->hist && t->hist_count && t->hist_head && t->type && t->update && t->any_update, "null robot state");
  RTD3_CHECK_ARG(t->base && t->ax && t->ay && t->reward && t->done, "null tick output");
  RTD3_CHECK_ARG((t->prev_x == nullptr) == (t->prev_y == nullptr), "prev_x / prev_y go together");
  RTD3_CHECK_ARG(t->num_demo >= 0 && (t->num_demo == 0 || t->demo), "demo set missing");
  RTD3_CHECK_ARG((t->demo_list_start == nullptr) == (t->demo_list == nullptr), "demo_list_start and demo_list go together");
  RTD3_CHECK_ARG(t->rp_s && t->rp_a && t->rp_r && t->rp_s2 && t->rp_notdone && t->rp_total && t->capacity > 0 && t->n <= t->capacity,
                 "bad replay ring");
  RTD3_CHECK_ARG(t->steps_bought && t->resets_bought, "null money counters");
  return 0;
}

}  // namespace rtd3

using namespace rtd3;

extern "C" {

int32_t rtd3_tick_pre(const rtd3_tick_state* t, void* stream) {
  if (int32_t e = check_state(t)) return e;
  if (t->n == 0) return 0;
  tick_pre_kernel<<<(int)ceil_div(t->n, 256), 256, 0, (cudaStream_t)stream>>>(*t);
  RTD3_LAUNCHED();
  return 0;
}

int32_t rtd3_tick_post(rtd3_env* h, const rtd3_tick_state* t, const float* residual, const double* unit_noise, int32_t noise_mode,
                       void* stream) {
  if (int32_t e = check_state(t)) return e;
  RTD3_CHECK_ARG(h && h->has_map, "environment has no dynamics map (call rtd3_env_set_map)");
  RTD3_CHECK_ARG(residual, "null residual");
  RTD3_CHECK_ARG(noise_mode >= RTD3_TICK_NOISE_NONE && noise_mode <= RTD3_TICK_NOISE_PHILOX, "unknown noise mode");
  RTD3_CHECK_ARG(noise_mode != RTD3_TICK_NOISE_GIVEN || unit_noise, "noise mode 'given' needs unit_noise");
  RTD3_CHECK_ARG(noise_mode != RTD3_TICK_NOISE_PHILOX || t->tick_counter, "noise mode 'philox' needs the tick counter");
  if (t->n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  // elementwise work plus a few candidate demo states per env: small CTAs spread a small batch over more SMs
  const int block = t->n <= (int64_t)h->num_sms * 256 ? 128 : 256;
  tick_post_kernel<<<(int)ceil_div(t->n, block), block, 0, st>>>(*t, h->table, residual, unit_noise, noise_mode);
  RTD3_LAUNCHED();
  return 0;
}

}  // extern "C"
