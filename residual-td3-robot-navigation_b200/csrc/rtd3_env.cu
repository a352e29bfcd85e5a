// Environment hot path: dynamics / step / T-step rollout / seeded init + reset.
// Reference behaviour: /root/reference/environment.py:28-56, 98-137 (see include/rtd3.h per entry point).
#include "rtd3_common.cuh"
#include "rtd3_mt.cuh"

struct rtd3_env {
  int device;
  int num_sms;
  float2* table;   // [100*100] (speed*cos(rot), speed*sin(rot)), indexed x*100+y
  bool has_map;
};

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
  table[i] = make_float2((float)(sp * c), (float)(sp * s));
}

// One env: the rotation form of environment.py:100-117 (no atan2):
//   s' = clip(s + speed*(ax*cos(rot) - ay*sin(rot), ax*sin(rot) + ay*cos(rot)), 0, 98.9999)
template <typename TableT>
__device__ __forceinline__ void dynamics_one(const TableT& table, float x, float y, float ax, float ay, float& nx,
                                             float& ny) {
  ax = clip_keep_nan(ax, -kMaxAction, kMaxAction);
  ay = clip_keep_nan(ay, -kMaxAction, kMaxAction);
  int cx = min(max(__float2int_rz(x), 0), kWorld - 1);   // int(state[0]); clamped only for memory safety
  int cy = min(max(__float2int_rz(y), 0), kWorld - 1);
  float2 cs = table[cx * kWorld + cy];
  nx = clip_keep_nan(x + fmaf(ax, cs.x, -ay * cs.y), 0.0f, kClipHi);
  ny = clip_keep_nan(y + fmaf(ax, cs.y, ay * cs.x), 0.0f, kClipHi);
}

// environment.py:125 - accept unless NaN (the clip already bounds everything else)
__device__ __forceinline__ bool in_world(float nx, float ny) {
  return nx >= 0.0f && nx < (float)kWorld && ny >= 0.0f && ny < (float)kWorld;
}

struct LdgTable {
  const float2* __restrict__ p;
  __device__ __forceinline__ float2 operator[](int i) const { return __ldg(p + i); }
};
struct SmemTable {
  const float2* p;
  __device__ __forceinline__ float2 operator[](int i) const { return p[i]; }
};

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
template <bool kKeepOnNan, typename TableT>
__device__ __forceinline__ void step_one(const TableT& table, float x, float y, float ax, float ay, float& ox, float& oy) {
  float nx, ny;
  dynamics_one(table, x, y, ax, ay, nx, ny);
  const bool keep = kKeepOnNan && !in_world(nx, ny);
  ox = keep ? x : nx;
  oy = keep ? y : ny;
}

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

template <bool kTraj>
__global__ void __launch_bounds__(256)
env_rollout_kernel(const float2* __restrict__ g_table, float* __restrict__ x, float* __restrict__ y,
                   const float* __restrict__ actions, float* __restrict__ traj, int64_t n, int64_t T) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* s_table = reinterpret_cast<float2*>(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + kTableBytes);
  stage_table(s_table, bar, g_table);

  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  const int64_t ii = live ? i : 0;
  float sx = x[ii], sy = y[ii];
  const float* a = actions + ii;          // ax(t) = a[(2t)*n], ay(t) = a[(2t+1)*n]
  float* tr = traj + ii;
  const int64_t n2 = 2 * n;

  float bax[kU], bay[kU];
  auto load_chunk = [&](int64_t t0) {
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t t = t0 + u;
      if (t < T) {
        bax[u] = __ldcs(a + t * n2);
        bay[u] = __ldcs(a + t * n2 + n);
      }
    }
  };
  load_chunk(0);
  mbar_wait(bar, 0);
  SmemTable table{s_table};

  for (int64_t t0 = 0; t0 < T; t0 += kU) {
    float cax[kU], cay[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) { cax[u] = bax[u]; cay[u] = bay[u]; }
    if (t0 + kU < T) load_chunk(t0 + kU);
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t t = t0 + u;
      if (t < T) {
        float nx, ny;
        dynamics_one(table, sx, sy, cax[u], cay[u], nx, ny);
        if (in_world(nx, ny)) { sx = nx; sy = ny; }
        if (kTraj && live) {
          __stcs(tr + t * n2, sx);
          __stcs(tr + t * n2 + n, sy);
        }
      }
    }
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

__global__ void mt_draw_gauss_kernel(rtd3_mt_bank b, double* __restrict__ out, int64_t k) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n) return;
  MtStream s{b.mt + i, b.n, b.pos[i]};
  int hg = b.has_gauss[i];
  double sp = b.gauss[i];
  for (int64_t j = 0; j < k; ++j) out[j * b.n + i] = mt_gauss(s, hg, sp);
  b.pos[i] = s.pos;
  b.has_gauss[i] = hg;
  b.gauss[i] = sp;
}

// environment.py:28-56
__global__ void init_goal_region_kernel(rtd3_mt_bank b, double* __restrict__ goal, double* __restrict__ region) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n) return;
  const int64_t n = b.n;
  MtStream s{b.mt + i, n, b.pos[i]};
  const double W = (double)kWorld, R = 25.0;   // constants.py:6, 25
  const uint32_t side = s.interval(3);         // np.random.choice([0,1,2,3])
  const double free_edge = s.uniform(0.0, W - R);
  double l, r, bt, tp;
  if (side == 0)      { l = 0.0;       r = R;            bt = free_edge; tp = __dadd_rn(free_edge, R); }
  else if (side == 1) { l = free_edge; r = __dadd_rn(free_edge, R); bt = W - R; tp = W; }
  else if (side == 2) { l = W - R;     r = W;            bt = free_edge; tp = __dadd_rn(free_edge, R); }
  else                { l = free_edge; r = __dadd_rn(free_edge, R); bt = 0.0;   tp = R; }
  const double mx = __dmul_rn(0.5, __dadd_rn(l, r)), my = __dmul_rn(0.5, __dadd_rn(bt, tp));
  double gx, gy, dist;
  do {
    gx = s.uniform(5.0, W - 5.0);
    gy = s.uniform(5.0, W - 5.0);
    dist = norm2_np(__dsub_rn(gx, mx), __dsub_rn(gy, my));
  } while (dist < 90.0);
  goal[i] = gx; goal[n + i] = gy;
  region[i] = l; region[n + i] = r; region[2 * n + i] = bt; region[3 * n + i] = tp;
  b.pos[i] = s.pos;
}

// environment.py:130-137
__global__ void env_reset_kernel(rtd3_mt_bank b, const double* __restrict__ region, const uint8_t* __restrict__ mask,
                                 float* __restrict__ x, float* __restrict__ y, double* __restrict__ state64) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n) return;
  if (mask && !mask[i]) return;
  const int64_t n = b.n;
  MtStream s{b.mt + i, n, b.pos[i]};
  const double sx = s.uniform(region[i], region[n + i]);           // x in [left, right)
  const double sy = s.uniform(region[2 * n + i], region[3 * n + i]);   // y in [bottom, top)
  x[i] = (float)sx; y[i] = (float)sy;
  if (state64) { state64[i] = sx; state64[n + i] = sy; }
  b.pos[i] = s.pos;
}

static int32_t check_bank(const rtd3_mt_bank* b) {
  RTD3_CHECK_ARG(b && b->mt && b->pos && b->has_gauss && b->gauss, "null MT bank pointer");
  RTD3_CHECK_ARG(b->n >= 0, "negative stream count");
  return 0;
}

}  // namespace rtd3

using namespace rtd3;

extern "C" {

int32_t rtd3_env_create(rtd3_env** out, int32_t device) {
  RTD3_CHECK_ARG(out, "out is null");
  int prev = 0;
  RTD3_CUDA(cudaGetDevice(&prev));
  RTD3_CUDA(cudaSetDevice(device));
  rtd3_env* h = new rtd3_env();
  h->device = device;
  h->has_map = false;
  h->table = nullptr;
  RTD3_CUDA(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
  RTD3_CUDA(cudaMalloc(&h->table, kTableBytes));
  const int smem = kTableBytes + 16;
  RTD3_CUDA(cudaFuncSetAttribute(env_step_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  RTD3_CUDA(cudaFuncSetAttribute(env_step_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  RTD3_CUDA(cudaFuncSetAttribute(env_rollout_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  RTD3_CUDA(cudaFuncSetAttribute(env_rollout_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  RTD3_CUDA(cudaSetDevice(prev));
  *out = h;
  return 0;
}

int32_t rtd3_env_destroy(rtd3_env* h) {
  if (!h) return 0;
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
  // spread small batches over all SMs (latency-bound: one dependent chain per env), cap CTA size at 256
  int64_t per_sm = ceil_div(n, (int64_t)h->num_sms);
  int block = (int)std::min<int64_t>(256, std::max<int64_t>(32, ceil_div(per_sm, 32) * 32));
  const int grid = (int)ceil_div(n, block);
  const int smem = kTableBytes + 16;
  if (traj) env_rollout_kernel<true><<<grid, block, smem, (cudaStream_t)stream>>>(h->table, x, y, actions, traj, n, T);
  else env_rollout_kernel<false><<<grid, block, smem, (cudaStream_t)stream>>>(h->table, x, y, actions, nullptr, n, T);
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
  if (int32_t e = check_bank(bank)) return e;
  RTD3_CHECK_ARG(out && k >= 0, "bad out/k");
  if (bank->n == 0 || k == 0) return 0;
  mt_draw_gauss_kernel<<<(int)ceil_div(bank->n, 128), 128, 0, (cudaStream_t)stream>>>(*bank, out, k);
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

int32_t rtd3_env_reset(const rtd3_mt_bank* bank, const double* region, const uint8_t* mask, float* x, float* y,
                       double* state64, void* stream) {
  if (int32_t e = check_bank(bank)) return e;
  RTD3_CHECK_ARG(region && x && y, "null argument");
  if (bank->n == 0) return 0;
  env_reset_kernel<<<(int)ceil_div(bank->n, 128), 128, 0, (cudaStream_t)stream>>>(*bank, region, mask, x, y, state64);
  RTD3_LAUNCHED();
  return 0;
}

}  // extern "C"
