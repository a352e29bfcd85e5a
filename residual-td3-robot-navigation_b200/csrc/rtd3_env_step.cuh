// Single env-step arithmetic shared by the step / rollout kernels (rtd3_env.cu) and the fused tick (rtd3_tick.cu).
// Reference behaviour: /root/reference/environment.py:98-127.
#pragma once
#include "rtd3_common.cuh"
#include "rtd3_mt.cuh"

struct rtd3_env {
  int device;
  int num_sms;
  float2* table;   // [100*100] (speed*cos(rot), speed*sin(rot)), indexed x*100+y
  bool has_map;
  void* host_pipe;   // rtd3::HostPipe of rtd3_env_rollout_host (rtd3_env_host.cu), created on first use
};

namespace rtd3 {

void host_pipe_destroy(rtd3_env* h);   // rtd3_env_host.cu: streams, staging buffers and cached graphs of the host-buffer rollout

// One env-step, the rotation form of environment.py:100-117 (no atan2):
//   s' = clip(s + speed*(ax*cos(rot) - ay*sin(rot), ax*sin(rot) + ay*cos(rot)), 0, 98.9999)
// split into the part that does not depend on the state (StepIn, off the dependent chain of a rollout) and the
// state recurrence itself, which is FADD.RZ -> IMAD -> IMAD -> LDS.64 -> FFMA -> FFMA -> FMNMX -> FMNMX.
struct StepIn {
  float ax, ay;   // clipped action (np.clip keeps NaN), zeroed when NaN
  float lo, hi;   // clip bounds: [0, 98.9999], or [-inf, +inf] when the action is NaN so that the state is kept
  bool bad;       // NaN action: environment.py:125 rejects the (NaN) next state and keeps robot_state
};

__device__ __forceinline__ StepIn prep_action(float ax, float ay) {
  StepIn in;
  ax = clip_keep_nan(ax, -kMaxAction, kMaxAction);
  ay = clip_keep_nan(ay, -kMaxAction, kMaxAction);
  in.bad = (ax != ax) || (ay != ay);
  in.ax = in.bad ? 0.0f : ax;
  in.ay = in.bad ? 0.0f : ay;
  in.lo = in.bad ? -INFINITY : 0.0f;
  in.hi = in.bad ? INFINITY : kClipHi;
  return in;
}

// s' from (s, table entry). Two dependent FFMAs per axis; the only roundings are at the magnitude of the state.
__device__ __forceinline__ void advance(float2 cs, const StepIn& in, float& x, float& y) {
  const float nx = fmaf(in.ax, cs.x, fmaf(-in.ay, cs.y, x));
  const float ny = fmaf(in.ax, cs.y, fmaf(in.ay, cs.x, y));
  x = fminf(fmaxf(nx, in.lo), in.hi);
  y = fminf(fmaxf(ny, in.lo), in.hi);
}

// int(state) -> cell, clamped so that any input stays inside the table (used by the single-step kernels)
__device__ __forceinline__ int cell_index(float x, float y) {
  const int cx = min(max(__float2int_rz(x), 0), kWorld - 1);
  const int cy = min(max(__float2int_rz(y), 0), kWorld - 1);
  return cx * kWorld + cy;
}

template <bool kKeepOnNan, typename TableT>
__device__ __forceinline__ void step_one(const TableT& table, float x, float y, float ax, float ay, float& ox, float& oy) {
  const StepIn in = prep_action(ax, ay);
  const float2 cs = table[cell_index(x, y)];
  advance(cs, in, x, y);
  if (!kKeepOnNan && in.bad) { x = NAN; y = NAN; }   // pure dynamics(): the NaN propagates
  ox = x;
  oy = y;
}

struct LdgTable {
  const float2* __restrict__ p;
  __device__ __forceinline__ float2 operator[](int i) const { return __ldg(p + i); }
};
struct SmemTable {
  const float2* p;
  __device__ __forceinline__ float2 operator[](int i) const { return p[i]; }
};

// Environment.reset for the lanes of one warp (environment.py:130-137): `active` lanes draw a new start state from their env's
// MT19937 stream.  All 32 lanes must call both halves (streams that wrap are twisted by the whole warp, rtd3_mt.cuh).
// Split in two so that a caller can put other work between the loads and their use: reset_load_warp issues the loads (stream
// position, init region and - unless a lane's stream wraps within the next four words, ~0.6 % of the draws - the four state
// words themselves, independent of each other), reset_finish_warp turns them into the state.
struct ResetLoads {
  int pos;
  double left, right, bottom, top;
  uint32_t w[4];
  bool have_words;   // warp-uniform
};

__device__ __forceinline__ ResetLoads reset_load_warp(const rtd3_mt_bank& b, const double* __restrict__ region, bool active, int64_t i) {
  ResetLoads r;
  r.pos = 0; r.left = r.right = r.bottom = r.top = 0.0; r.have_words = false;
#pragma unroll
  for (int q = 0; q < 4; ++q) r.w[q] = 0u;
  if (!__any_sync(0xffffffffu, active)) return r;
  const int64_t n = b.n;
  if (active) {
    r.pos = b.pos[i];
    r.left = region[i]; r.right = region[n + i]; r.bottom = region[2 * n + i]; r.top = region[3 * n + i];
  }
  r.have_words = !__any_sync(0xffffffffu, active && r.pos + 4 > RTD3_MT_N);
  if (r.have_words && active) {
#pragma unroll
    for (int q = 0; q < 4; ++q) r.w[q] = mt_temper(b.mt[(int64_t)(r.pos + q) * n + i]);
  }
  return r;
}

__device__ __forceinline__ void reset_finish_warp(const rtd3_mt_bank& b, bool active, int64_t i, ResetLoads& r, float* __restrict__ x,
                                                  float* __restrict__ y, double* __restrict__ state64) {
  if (!__any_sync(0xffffffffu, active)) return;
  const int64_t n = b.n;
  int pos_after = r.pos + 4;
  if (!r.have_words) {                               // a stream of this warp wraps: draw word by word through the warp twist
    int pos = r.pos;
    mt_next_words4_warp_slow(b.mt + (active ? i : 0), n, &pos, active, r.w);
    pos_after = pos;
  }
  if (!active) return;
  // lo + (hi - lo) * u with numpy's two roundings
  const double ux = mt_double_from_words(r.w[0], r.w[1]), uy = mt_double_from_words(r.w[2], r.w[3]);
  const double sx = __dadd_rn(r.left, __dmul_rn(__dsub_rn(r.right, r.left), ux));        // x in [left, right)
  const double sy = __dadd_rn(r.bottom, __dmul_rn(__dsub_rn(r.top, r.bottom), uy));      // y in [bottom, top)
  // float32 rounding must not reach 100.0 (the cell index would leave the map): cap at the largest float below it
  x[i] = fminf((float)sx, 99.99999f); y[i] = fminf((float)sy, 99.99999f);
  if (state64) { state64[i] = sx; state64[n + i] = sy; }
  b.pos[i] = pos_after;
}

__device__ __forceinline__ void reset_env_warp(const rtd3_mt_bank& b, const double* __restrict__ region, bool active, int64_t i,
                                               float* __restrict__ x, float* __restrict__ y, double* __restrict__ state64) {
  ResetLoads r = reset_load_warp(b, region, active, i);
  reset_finish_warp(b, active, i, r, x, y, state64);
}

}  // namespace rtd3
