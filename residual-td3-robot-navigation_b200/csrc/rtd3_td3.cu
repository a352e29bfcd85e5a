// Residual-TD3 learner hot path: replay ring, minibatch gather, twin-critic / actor steps, Adam + Polyak.
// Reference behaviour: /root/reference/robot.py:58-124 (ReplayBuffer), :128-206 (networks), :258-398 (TD3).
#include "rtd3_common.cuh"
#include "rtd3_mlp.cuh"

namespace rtd3 {

struct ReplayView {
  const float2* s;
  const float2* a;
  const float* r;
  const float2* s2;
  const float* notdone;
};

struct Td3Hyper {
  float gamma, policy_noise, noise_clip, max_action;
};

// Parameter arena: [actor | critic1 | critic2 | target actor | target critic1 | target critic2], each slot padded to 4 floats.
struct Arena {
  NetShape actor, critic;
  __host__ __device__ int64_t sa() const { return net_stride(actor); }
  __host__ __device__ int64_t sc() const { return net_stride(critic); }
  __host__ __device__ int64_t off(int net) const {   // 0 actor, 1 critic1, 2 critic2, 3..5 targets
    const int64_t a = sa(), c = sc();
    switch (net) {
      case 0: return 0;
      case 1: return a;
      case 2: return a + c;
      case 3: return a + 2 * c;
      case 4: return 2 * a + 2 * c;
      default: return 2 * a + 3 * c;
    }
  }
  __host__ __device__ int64_t online_total() const { return sa() + 2 * sc(); }
  __host__ __device__ int64_t total() const { return 2 * online_total(); }
};

// ---- critic phase: robot.py:312-366 up to (not including) the optimiser steps ------------------------------------
//   y = r + gamma * min(Q1', Q2')(s2, clip(pi'(s2) + clip(noise*sigma, +-c), +-5)) * notdone
//   L_i = mean((Q_i(s,a) - y)^2);  gradients of L_1, L_2 accumulated (RED.ADD) into grads[critic1], grads[critic2]
// One CTA = R batch rows through the whole chain; steps[1] (the critics' Adam step counter) is advanced by block 0.
template <int R>
__global__ void __launch_bounds__(kThreads, 1)
td3_critic_kernel(Arena ar, const float* __restrict__ params, float* __restrict__ grads, ReplayView rp, const int32_t* __restrict__ idx,
                  const float* __restrict__ noise /*[B][2] unit normal*/, int B, Td3Hyper hp, float* __restrict__ loss /*[2]*/,
                  float* __restrict__ q_out /*nullable [2][B]*/, float* __restrict__ y_out /*nullable [B]*/, int32_t* __restrict__ steps) {
  extern __shared__ __align__(16) float smem_f[];
  MlpSmem<R> sm;
  sm.carve(smem_f, ar.critic.hid, ar.critic.layers);
  const int r0 = blockIdx.x * R;
  const int t = threadIdx.x;
  if (blockIdx.x == 0 && t == 0) steps[1] += 1;

  float* S = sm.scratch;   // [R][8]: 0 s.x 1 s.y 2 a.x 3 a.y 4 reward 5 notdone 6 y 7 valid
  if (t < R) {
    const int row = r0 + t;
    const bool valid = row < B;
    const int j = valid ? idx[row] : 0;
    const float2 s = rp.s[j], a = rp.a[j], s2 = rp.s2[j];
    S[t * 8 + 0] = s.x; S[t * 8 + 1] = s.y; S[t * 8 + 2] = a.x; S[t * 8 + 3] = a.y;
    S[t * 8 + 4] = rp.r[j]; S[t * 8 + 5] = rp.notdone[j]; S[t * 8 + 7] = valid ? 1.f : 0.f;
    sm.in0[t * 4 + 0] = s2.x; sm.in0[t * 4 + 1] = s2.y; sm.in0[t * 4 + 2] = 0.f; sm.in0[t * 4 + 3] = 0.f;
  }
  __syncthreads();

  // target actor on s2, then smoothing noise and clip (robot.py:338-339)
  mlp_forward<R>(params + ar.off(3), ar.actor, sm, false);
  if (t < R) {
    const int row = min(r0 + t, B - 1);
    float a2[2];
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      float e = noise[row * 2 + o] * hp.policy_noise;
      e = fminf(fmaxf(e, -hp.noise_clip), hp.noise_clip);
      a2[o] = fminf(fmaxf(sm.out[t * 2 + o] + e, -hp.max_action), hp.max_action);
    }
    sm.in0[t * 4 + 2] = a2[0];
    sm.in0[t * 4 + 3] = a2[1];
  }
  __syncthreads();
  // target critics on (s2, a') and the clipped double-Q target (robot.py:342-345)
  mlp_forward<R>(params + ar.off(4), ar.critic, sm, false);
  if (t < R) S[t * 8 + 6] = sm.out[t * 2];
  __syncthreads();
  mlp_forward<R>(params + ar.off(5), ar.critic, sm, false);
  if (t < R) {
    const float qmin = fminf(S[t * 8 + 6], sm.out[t * 2]);
    const float y = S[t * 8 + 4] + hp.gamma * qmin * S[t * 8 + 5];
    S[t * 8 + 6] = y;
    if (y_out && r0 + t < B) y_out[r0 + t] = y;
    sm.in0[t * 4 + 0] = S[t * 8 + 0]; sm.in0[t * 4 + 1] = S[t * 8 + 1];
    sm.in0[t * 4 + 2] = S[t * 8 + 2]; sm.in0[t * 4 + 3] = S[t * 8 + 3];
  }
  __syncthreads();

  // both critics: forward (activations kept), MSE loss, backward (robot.py:348-363)
  for (int c = 0; c < 2; ++c) {
    const float* P = params + ar.off(1 + c);
    float* G = grads + ar.off(1 + c);
    mlp_forward<R>(P, ar.critic, sm, true);
    if (t < R) {
      const float valid = S[t * 8 + 7];
      const float q = sm.out[t * 2];
      const float diff = (q - S[t * 8 + 6]) * valid;
      sm.dout[t * 2] = 2.0f * diff / (float)B;
      sm.dout[t * 2 + 1] = 0.f;
      if (q_out && valid != 0.f) q_out[c * B + r0 + t] = q;
      float l = diff * diff / (float)B;                 // this row's share of the mean
#pragma unroll
      for (int o = R / 2; o > 0; o >>= 1) l += __shfl_xor_sync((R >= 32) ? 0xffffffffu : ((1u << R) - 1u), l, o);
      if (t == 0) atomicAdd(loss + c, l);
    }
    __syncthreads();
    mlp_backward<R>(P, G, ar.critic, sm, false);
    __syncthreads();
  }
}

// ---- actor phase: robot.py:369-398 up to the optimiser step ---------------------------------------------------------
//   L = -mean(Q1(s, pi(s)));  gradient w.r.t. the actor only (critic-1 parameter gradients are discarded by the reference)
template <int R>
__global__ void __launch_bounds__(kThreads, 1)
td3_actor_kernel(Arena ar, const float* __restrict__ params, float* __restrict__ grads, ReplayView rp, const int32_t* __restrict__ idx,
                 int B, float* __restrict__ loss /*[1]*/, int32_t* __restrict__ steps) {
  extern __shared__ __align__(16) float smem_f[];
  MlpSmem<R> sm;                      // critic pass (activations kept, backward for dQ/da)
  sm.carve(smem_f, ar.critic.hid, ar.critic.layers);
  MlpSmem<R> sa;                      // actor pass: own activation buffers behind the critic's working set
  float* actor_base = smem_f + MlpSmem<R>::bytes(ar.critic.hid, ar.critic.layers) / sizeof(float);
  const int r0 = blockIdx.x * R;
  const int t = threadIdx.x;
  if (blockIdx.x == 0 && t == 0) steps[0] += 1;
  // the actor working set shares the weight-tile ring and scratch with `sm`; only its kept activations are separate
  sa = sm;
  for (int l = 0; l < ar.actor.layers; ++l) sa.act[l] = actor_base + l * R * sm.ld;
  float* actor_in = actor_base + ar.actor.layers * R * sm.ld;   // [R][4]
  float* S = sm.scratch;

  if (t < R) {
    const int row = r0 + t;
    const bool valid = row < B;
    const float2 s = rp.s[valid ? idx[row] : 0];
    actor_in[t * 4 + 0] = s.x; actor_in[t * 4 + 1] = s.y; actor_in[t * 4 + 2] = 0.f; actor_in[t * 4 + 3] = 0.f;
    S[t * 8 + 7] = valid ? 1.f : 0.f;
  }
  __syncthreads();
  sa.in0 = actor_in;
  mlp_forward<R>(params + ar.off(0), ar.actor, sa, true);        // a = pi(s), fed the raw replay state (robot.py:386)
  if (t < R) {
    sm.in0[t * 4 + 0] = actor_in[t * 4 + 0]; sm.in0[t * 4 + 1] = actor_in[t * 4 + 1];
    sm.in0[t * 4 + 2] = sa.out[t * 2]; sm.in0[t * 4 + 3] = sa.out[t * 2 + 1];
  }
  __syncthreads();
  mlp_forward<R>(params + ar.off(1), ar.critic, sm, true);       // Q1(s, a)
  if (t < R) {
    const float valid = S[t * 8 + 7];
    sm.dout[t * 2] = -valid / (float)B;
    sm.dout[t * 2 + 1] = 0.f;
    float l = -sm.out[t * 2] * valid / (float)B;
#pragma unroll
    for (int o = R / 2; o > 0; o >>= 1) l += __shfl_xor_sync((R >= 32) ? 0xffffffffu : ((1u << R) - 1u), l, o);
    if (t == 0) atomicAdd(loss, l);
  }
  __syncthreads();
  mlp_backward<R>(params + ar.off(1), nullptr, ar.critic, sm, true);   // only dQ/d(input) is needed
  if (t < R) {
    sa.dout[t * 2] = sm.din[t * 4 + 2];
    sa.dout[t * 2 + 1] = sm.din[t * 4 + 3];
  }
  __syncthreads();
  mlp_backward<R>(params + ar.off(0), grads + ar.off(0), ar.actor, sa, false);
}

// ---- Adam (torch.optim.Adam defaults, robot.py:237-239) + optional Polyak (robot.py:293-310), one pass --------------
// nets: bit 0 actor, bit 1 critic1, bit 2 critic2 get an Adam step from `grads` (then the gradients are zeroed);
// polyak: same bit layout; afterwards the selected target slots are blended with their (updated) online net:
// t = t*(1-tau) + p*tau.
__global__ void td3_adam_polyak_kernel(Arena ar, float* __restrict__ params, float* __restrict__ grads, float* __restrict__ m,
                                       float* __restrict__ v, const int32_t* __restrict__ steps, int nets, float lr_actor, float lr_critic,
                                       float grad_scale, int polyak, float tau) {
  __shared__ float s_step[2], s_bc2[2];
  if (threadIdx.x < 2) {
    const double tt = (double)steps[threadIdx.x];                     // 0: actor optimiser, 1: both critic optimisers
    const double bc1 = 1.0 - pow(0.9, tt), bc2 = 1.0 - pow(0.999, tt);
    const double lr = threadIdx.x == 0 ? (double)lr_actor : (double)lr_critic;
    s_step[threadIdx.x] = (float)(lr / bc1);
    s_bc2[threadIdx.x] = (float)sqrt(bc2);
  }
  __syncthreads();
  const int64_t n_online = ar.online_total();
  const int64_t c1 = ar.off(1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_online; i += (int64_t)gridDim.x * blockDim.x) {
    const int net = i < c1 ? 0 : (i < ar.off(2) ? 1 : 2);
    float p = params[i];
    if ((nets >> net) & 1) {
      const int o = net == 0 ? 0 : 1;
      const float g = grads[i] * grad_scale;
      grads[i] = 0.f;
      const float mi = m[i] + (g - m[i]) * 0.1f;                      // exp_avg.lerp_(grad, 1 - beta1)
      const float vi = v[i] * 0.999f + (g * g) * 0.001f;              // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
      m[i] = mi;
      v[i] = vi;
      const float denom = sqrtf(vi) / s_bc2[o] + 1e-8f;
      p = p - s_step[o] * (mi / denom);
      params[i] = p;
    }
    if ((polyak >> net) & 1) {
      const int64_t ti = n_online + i;                                // target slots mirror the online layout
      // torch evaluates target*(1-tau) + source*tau as three separately rounded float32 ops (robot.py:309): no fma here
      params[ti] = __fadd_rn(__fmul_rn(params[ti], 1.0f - tau), __fmul_rn(p, tau));
    }
  }
}

// ---- plain forward of one network over B rows (actor inference for get_next_action, parity checks of Q-values) ------
template <int R>
__global__ void __launch_bounds__(kThreads, 1)
mlp_forward_kernel(NetShape s, const float* __restrict__ P, const float* __restrict__ x /*[B][in]*/, float* __restrict__ y /*[B][out]*/,
                   int B) {
  extern __shared__ __align__(16) float smem_f[];
  MlpSmem<R> sm;
  sm.carve(smem_f, s.hid, 0);
  const int r0 = blockIdx.x * R, t = threadIdx.x;
  if (t < R * 4) {
    const int r = t >> 2, j = t & 3, row = r0 + r;
    sm.in0[t] = (row < B && j < s.in) ? x[(int64_t)row * s.in + j] : 0.f;
  }
  __syncthreads();
  mlp_forward<R>(P, s, sm, false);
  if (t < R * s.out) {
    const int r = t / s.out, o = t - r * s.out;
    if (r0 + r < B) y[(int64_t)(r0 + r) * s.out + o] = sm.out[r * 2 + o];
  }
}

// ---- replay ring (robot.py:79-96): push n rows starting at `position`, wrapping at capacity -------------------------
__global__ void replay_push_kernel(float2* s, float2* a, float* r, float2* s2, float* notdone, int64_t capacity, int64_t position,
                                   const float* __restrict__ sx, const float* __restrict__ sy, const float* __restrict__ ax,
                                   const float* __restrict__ ay, const float* __restrict__ rew, const float* __restrict__ nx,
                                   const float* __restrict__ ny, const uint8_t* __restrict__ done, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = (position + i) % capacity;
  s[p] = make_float2(sx[i], sy[i]);
  a[p] = make_float2(ax[i], ay[i]);
  r[p] = rew[i];
  s2[p] = make_float2(nx[i], ny[i]);
  notdone[p] = done[i] ? 0.f : 1.f;
}

// ReplayBuffer.sample's gather (robot.py:113-115): rows idx[b] -> dense minibatch arrays
__global__ void replay_gather_kernel(ReplayView rp, const int32_t* __restrict__ idx, int B, float2* os, float2* oa, float* orw,
                                     float2* os2, float* ond) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int j = idx[b];
  os[b] = rp.s[j]; oa[b] = rp.a[j]; orw[b] = rp.r[j]; os2[b] = rp.s2[j]; ond[b] = rp.notdone[j];
}

static bool shape_ok(const NetShape& s) {
  return s.in >= 1 && s.in <= 4 && s.out >= 1 && s.out <= 2 && s.layers >= 1 && s.layers <= 4 && s.hid >= 4 && s.hid <= kMaxHidden &&
         s.hid % 4 == 0;
}

}  // namespace rtd3

using namespace rtd3;

struct rtd3_td3 {
  Arena ar;
  int device;
  int num_sms;
  size_t smem_critic[2], smem_actor[2], smem_fwd[2];   // per row-tile size: [0] R=8, [1] R=16
};

static const int kRowTiles[2] = {8, 16};

static inline int pick_tile(int B, int num_sms) { return (B > 8 * num_sms * 2) ? 1 : 0; }

static size_t actor_phase_smem(const Arena& ar, int R) {
  return mlp_smem_bytes(R, ar.critic.hid, ar.critic.layers) + ((size_t)ar.actor.layers * R * (ar.critic.hid + 4) + R * 4) * sizeof(float);
}

extern "C" {

int32_t rtd3_td3_create(rtd3_td3** out, int32_t device, int32_t hidden, int32_t layers) {
  RTD3_CHECK_ARG(out, "out is null");
  rtd3_td3* h = new rtd3_td3();
  h->ar.actor = NetShape{2, hidden, layers, 2};
  h->ar.critic = NetShape{4, hidden, layers, 1};
  if (!shape_ok(h->ar.actor)) {
    delete h;
    rtd3::set_error("rtd3_td3_create: hidden must be a multiple of 4 in [4,%d], layers in [1,4]", kMaxHidden);
    return RTD3_ERR_ARG;
  }
  h->device = device;
  int prev = 0;
  RTD3_CUDA(cudaGetDevice(&prev));
  RTD3_CUDA(cudaSetDevice(device));
  RTD3_CUDA(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
  for (int i = 0; i < 2; ++i) {
    const int R = kRowTiles[i];
    h->smem_critic[i] = mlp_smem_bytes(R, hidden, layers);
    h->smem_actor[i] = actor_phase_smem(h->ar, R);
    h->smem_fwd[i] = mlp_smem_bytes(R, hidden, 0);
  }
  RTD3_CUDA(cudaFuncSetAttribute(td3_critic_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_critic[0]));
  RTD3_CUDA(cudaFuncSetAttribute(td3_critic_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_critic[1]));
  RTD3_CUDA(cudaFuncSetAttribute(td3_actor_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_actor[0]));
  RTD3_CUDA(cudaFuncSetAttribute(td3_actor_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_actor[1]));
  RTD3_CUDA(cudaFuncSetAttribute(mlp_forward_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_fwd[0]));
  RTD3_CUDA(cudaFuncSetAttribute(mlp_forward_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_fwd[1]));
  RTD3_CUDA(cudaSetDevice(prev));
  *out = h;
  return 0;
}

int32_t rtd3_td3_destroy(rtd3_td3* h) {
  delete h;
  return 0;
}

int64_t rtd3_td3_param_count(const rtd3_td3* h, int32_t net) {
  if (!h) return -1;
  return net_param_count(net == 0 || net == 3 ? h->ar.actor : h->ar.critic);
}
int64_t rtd3_td3_param_offset(const rtd3_td3* h, int32_t net) { return h ? h->ar.off(net) : -1; }
int64_t rtd3_td3_arena_floats(const rtd3_td3* h) { return h ? h->ar.total() : -1; }

int32_t rtd3_td3_critic_step(rtd3_td3* h, const float* params, float* grads, const float* rp_s, const float* rp_a, const float* rp_r,
                             const float* rp_s2, const float* rp_notdone, const int32_t* idx, const float* noise, int32_t batch,
                             float gamma, float policy_noise, float noise_clip, float max_action, float* loss2, float* q_out, float* y_out,
                             int32_t* steps, void* stream) {
  RTD3_CHECK_ARG(h && params && grads && rp_s && rp_a && rp_r && rp_s2 && rp_notdone && idx && noise && loss2 && steps, "null argument");
  RTD3_CHECK_ARG(batch > 0, "batch must be positive");
  ReplayView rp{(const float2*)rp_s, (const float2*)rp_a, rp_r, (const float2*)rp_s2, rp_notdone};
  Td3Hyper hp{gamma, policy_noise, noise_clip, max_action};
  const int ti = pick_tile(batch, h->num_sms);
  const int R = kRowTiles[ti];
  const int grid = (batch + R - 1) / R;
  cudaStream_t st = (cudaStream_t)stream;
  if (ti == 0)
    td3_critic_kernel<8><<<grid, kThreads, h->smem_critic[0], st>>>(h->ar, params, grads, rp, idx, noise, batch, hp, loss2, q_out, y_out, steps);
  else
    td3_critic_kernel<16><<<grid, kThreads, h->smem_critic[1], st>>>(h->ar, params, grads, rp, idx, noise, batch, hp, loss2, q_out, y_out, steps);
  RTD3_LAUNCHED();
  return 0;
}

int32_t rtd3_td3_actor_step(rtd3_td3* h, const float* params, float* grads, const float* rp_s, const int32_t* idx, int32_t batch,
                            float* loss1, int32_t* steps, void* stream) {
  RTD3_CHECK_ARG(h && params && grads && rp_s && idx && loss1 && steps, "null argument");
  RTD3_CHECK_ARG(batch > 0, "batch must be positive");
  ReplayView rp{(const float2*)rp_s, nullptr, nullptr, nullptr, nullptr};
  const int ti = pick_tile(batch, h->num_sms);
  const int R = kRowTiles[ti];
  const int grid = (batch + R - 1) / R;
  cudaStream_t st = (cudaStream_t)stream;
  if (ti == 0) td3_actor_kernel<8><<<grid, kThreads, h->smem_actor[0], st>>>(h->ar, params, grads, rp, idx, batch, loss1, steps);
  else td3_actor_kernel<16><<<grid, kThreads, h->smem_actor[1], st>>>(h->ar, params, grads, rp, idx, batch, loss1, steps);
  RTD3_LAUNCHED();
  return 0;
}

int32_t rtd3_td3_adam_polyak(rtd3_td3* h, float* params, float* grads, float* adam_m, float* adam_v, const int32_t* steps, int32_t nets,
                             float lr_actor, float lr_critic, float grad_scale, int32_t polyak, float tau, void* stream) {
  RTD3_CHECK_ARG(h && params && grads && adam_m && adam_v && steps, "null argument");
  const int64_t n = h->ar.online_total();
  const int block = 256;
  const int grid = (int)std::min<int64_t>(ceil_div(n, block), (int64_t)h->num_sms * 4);
  td3_adam_polyak_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(h->ar, params, grads, adam_m, adam_v, steps, nets, lr_actor, lr_critic,
                                                                    grad_scale, polyak, tau);
  RTD3_LAUNCHED();
  return 0;
}

int32_t rtd3_mlp_forward(rtd3_td3* h, int32_t net, const float* params, const float* x, float* y, int64_t batch, void* stream) {
  RTD3_CHECK_ARG(h && params && x && y, "null argument");
  RTD3_CHECK_ARG(net >= 0 && net < 6, "net index out of range");
  RTD3_CHECK_ARG(batch >= 0 && batch < (1ll << 31), "bad batch");
  if (batch == 0) return 0;
  const NetShape s = (net == 0 || net == 3) ? h->ar.actor : h->ar.critic;
  const int ti = pick_tile((int)batch, h->num_sms);
  const int R = kRowTiles[ti];
  const int grid = (int)((batch + R - 1) / R);
  cudaStream_t st = (cudaStream_t)stream;
  if (ti == 0) mlp_forward_kernel<8><<<grid, kThreads, h->smem_fwd[0], st>>>(s, params + h->ar.off(net), x, y, (int)batch);
  else mlp_forward_kernel<16><<<grid, kThreads, h->smem_fwd[1], st>>>(s, params + h->ar.off(net), x, y, (int)batch);
  RTD3_LAUNCHED();
  return 0;
}

int32_t rtd3_replay_push(float* s, float* a, float* r, float* s2, float* notdone, int64_t capacity, int64_t position, const float* sx,
                         const float* sy, const float* ax, const float* ay, const float* reward, const float* nx, const float* ny,
                         const uint8_t* done, int64_t n, void* stream) {
  RTD3_CHECK_ARG(s && a && r && s2 && notdone && sx && sy && ax && ay && reward && nx && ny && done, "null argument");
  RTD3_CHECK_ARG(capacity > 0 && position >= 0 && position < capacity && n >= 0, "bad capacity/position/n");
  if (n == 0) return 0;
  replay_push_kernel<<<(int)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>((float2*)s, (float2*)a, r, (float2*)s2, notdone, capacity,
                                                                              position, sx, sy, ax, ay, reward, nx, ny, done, n);
  RTD3_LAUNCHED();
  return 0;
}

int32_t rtd3_replay_gather(const float* s, const float* a, const float* r, const float* s2, const float* notdone, const int32_t* idx,
                           int32_t batch, float* out_s, float* out_a, float* out_r, float* out_s2, float* out_notdone, void* stream) {
  RTD3_CHECK_ARG(s && a && r && s2 && notdone && idx && out_s && out_a && out_r && out_s2 && out_notdone, "null argument");
  RTD3_CHECK_ARG(batch >= 0, "negative batch");
  if (batch == 0) return 0;
  ReplayView rp{(const float2*)s, (const float2*)a, r, (const float2*)s2, notdone};
  replay_gather_kernel<<<(batch + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rp, idx, batch, (float2*)out_s, (float2*)out_a, out_r,
                                                                              (float2*)out_s2, out_notdone);
  RTD3_LAUNCHED();
  return 0;
}

}  // extern "C"
