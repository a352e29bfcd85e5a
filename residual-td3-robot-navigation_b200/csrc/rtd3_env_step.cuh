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
};

namespace rtd3 {

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
// MT19937 stream.  All 32 lanes must call it (streams that wrap are twisted by the whole warp, rtd3_mt.cuh).
__device__ __forceinline__ void reset_env_warp(const rtd3_mt_bank& b, const double* __restrict__ region, bool active, int64_t i,
                                               float* __restrict__ x, float* __restrict__ y, double* __restrict__ state64) {
  if (!__any_sync(0xffffffffu, active)) return;
  const int64_t n = b.n;
  const int64_t ii = active ? i : 0;
  MtStream s{b.mt + ii, n, active ? b.pos[ii] : 0};
  // lo + (hi - lo) * u with numpy's two roundings; the draws go through the warp-cooperative wrap (see rtd3_mt.cuh)
  double ux, uy;
  mt_next_double2_warp(s, active, ux, uy);
  if (!active) return;
  const double sx = __dadd_rn(region[i], __dmul_rn(__dsub_rn(region[n + i], region[i]), ux));                   // x in [left, right)
  const double sy = __dadd_rn(region[2 * n + i], __dmul_rn(__dsub_rn(region[3 * n + i], region[2 * n + i]), uy));   // y in [bottom, top)
  // float32 rounding must not reach 100.0 (the cell index would leave the map): cap at the largest float below it
  x[i] = fminf((float)sx, 99.99999f); y[i] = fminf((float)sy, 99.99999f);
  if (state64) { state64[i] = sx; state64[n + i] = sy; }
  b.pos[i] = s.pos;
}

}  // namespace rtd3
