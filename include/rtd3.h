/*
 * rtd3.h - C ABI of librtd3.so: the B200 (sm_100a) hot path of
 * benmcclusky/Residual-TD3-Robot-Navigation.
 *
 * The reference has no FFI of its own (it is pure Python); the functions below are what a
 * ctypes binding placed under the reference's `Environment` / `Robot` / `TD3` / `ReplayBuffer`
 * classes calls.  Each entry point cites the reference interface it replaces (file:line under
 * the reference repo root).  See INTEGRATION.md for the binding stubs.
 *
 * Conventions
 *   - every function returns int32_t: 0 = ok, >0 = cudaError_t, <0 = argument error;
 *     rtd3_last_error() returns a thread-local, NUL-terminated message for the last failure;
 *   - unless a parameter is marked HOST, pointers are DEVICE pointers owned by the caller
 *     (torch tensors); the library never frees user buffers;
 *   - launching entry points take an explicit cudaStream_t (as void*), never synchronise it
 *     and never allocate on it;
 *   - environment-indexed arrays are struct-of-arrays: x[n], y[n] ... ; a "[k][n]" array is k
 *     planes of n elements (plane-major) so that consecutive envs are consecutive in memory;
 *   - sizes are int64_t; handles are per-device and not re-entrant (one host thread per handle).
 */
#ifndef RTD3_H_
#define RTD3_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTD3_VERSION 100
#define RTD3_WORLD_SIZE 100          /* constants.py:6  */
#define RTD3_MAP_CELLS 10000         /* 100 x 100 cells, indexed [x][y] (environment.py:105-111) */
#define RTD3_MT_N 624                /* MT19937 state words */

#define RTD3_ERR_ARG (-1)
#define RTD3_ERR_STATE (-2)

int32_t rtd3_version(void);
const char* rtd3_last_error(void);
/* Number of kernels this library has launched since load / since the last reset (bench.py's gpu_launches). */
int64_t rtd3_launch_count(void);
void rtd3_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------------
 * Environment: dynamics / step / rollout        (environment.py:98-127)
 * ---------------------------------------------------------------------------------------------- */
typedef struct rtd3_env rtd3_env;

/* Per-device handle; owns the packed dynamics table (speed*cos(rot), speed*sin(rot)) in HBM. */
int32_t rtd3_env_create(rtd3_env** out, int32_t device);
int32_t rtd3_env_destroy(rtd3_env* h);

/* Replaces assigning Environment.dynamics_speed / dynamics_angle (environment.py:20-21, 84, 95).
 * speed, angle: DEVICE float32 [100*100], row-major [x][y].  Builds the packed table on `stream`:
 * rot = float32(angle*2*pi) (float32 as numpy>=2 evaluates environment.py:107), then
 * (speed*cos(rot), speed*sin(rot)) evaluated in float64 and rounded once to float32. */
int32_t rtd3_env_set_map(rtd3_env* h, const float* speed, const float* angle, void* stream);

/* Variant selector for rtd3_env_step (evidence for each lives in profiles/). */
#define RTD3_STEP_AUTO 0
#define RTD3_STEP_SMEM 1   /* table staged in shared memory by one bulk-async copy per CTA */
#define RTD3_STEP_LDG 2    /* table read through the read-only L1 path */

/* Environment.step over n envs (environment.py:122-127 -> dynamics :98-119).
 * x,y: state in/out; ax,ay: actions (clipped to +-5 inside).  A NaN result keeps the old state. */
int32_t rtd3_env_step(rtd3_env* h, float* x, float* y, const float* ax, const float* ay, int64_t n,
                      int32_t variant, void* stream);

/* Environment.dynamics (pure, environment.py:98-119): out_x,out_y = f(x,y,ax,ay); NaN propagates. */
int32_t rtd3_env_dynamics(rtd3_env* h, const float* x, const float* y, const float* ax, const float* ay,
                          float* out_x, float* out_y, int64_t n, void* stream);

/* T consecutive Environment.step calls in one launch (the loop at robot-learning.py:97-100 with
 * actions supplied): actions [T][2][n] float32, traj (nullable) [T][2][n] receives the state after
 * every step; x,y are updated to the final state.  State lives in registers across the T steps. */
int32_t rtd3_env_rollout(rtd3_env* h, float* x, float* y, const float* actions, float* traj, int64_t n,
                         int64_t T, void* stream);

/* ------------------------------------------------------------------------------------------------
 * numpy-legacy MT19937 streams, one per env      (robot-learning.py:19; numpy RandomState)
 * ---------------------------------------------------------------------------------------------- */
typedef struct rtd3_mt_bank {
  uint32_t* mt;        /* [624][n] state words, word-major                      */
  int32_t* pos;        /* [n] next word index, 624 = regenerate on next draw    */
  int32_t* has_gauss;  /* [n] legacy_gauss spare flag                           */
  double* gauss;       /* [n] legacy_gauss spare value                          */
  int64_t n;
} rtd3_mt_bank;

/* np.random.seed(seeds[i]) for stream i (mt19937_seed / init_genrand). seeds: DEVICE uint32 [n]. */
int32_t rtd3_mt_seed(const rtd3_mt_bank* bank, const uint32_t* seeds, void* stream);
/* Raw draws for tests: out[k][n] = k-th next uint32 of every stream. */
int32_t rtd3_mt_draw_u32(const rtd3_mt_bank* bank, uint32_t* out, int64_t k, void* stream);
/* np.random.normal(0,1) draws (legacy polar method with spare): out[k][n] float64. */
int32_t rtd3_mt_draw_gauss(const rtd3_mt_bank* bank, double* out, int64_t k, void* stream);

/* Environment.set_init_and_goal (environment.py:28-56) for n envs, each on its own stream.
 * goal [2][n] float64, region [4][n] float64 = left,right,bottom,top.  Bit-exact vs numpy. */
int32_t rtd3_env_init_goal_region(const rtd3_mt_bank* bank, double* goal, double* region, void* stream);

/* Environment.reset / get_random_robot_init_state (environment.py:130-137) for the envs whose
 * mask byte is non-zero (mask NULL = all).  Writes float32 state x,y; if state64 ([2][n] float64)
 * is non-NULL also the reference's float64 draw, which is bit-exact vs numpy. */
int32_t rtd3_env_reset(const rtd3_mt_bank* bank, const double* region, const uint8_t* mask, float* x, float* y,
                       double* state64, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RTD3_H_ */
