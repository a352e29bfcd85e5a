// Standalone probe: tcgen05.mma kind::tf32 with MN-major operands (dW = dZ^T H over 64 batch rows) + 2-D TMA reduce-add.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mn_test mn_test.cu -lcuda ; run on a B200.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../residual-td3-robot-navigation_b200/csrc/rtd3_tc.cuh"
namespace rtd3 { void set_error(const char*, ...) {} void count_launch(int) {} }
using namespace rtd3;

constexpr int H = 128, ROWS = 64;

__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, int c0, int c1, const void* smem_src) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(tm)),
               "r"(c0), "r"(c1), "r"(smem_u32(smem_src)) : "memory");
}

__global__ void __launch_bounds__(160, 1)
probe(const float* dz /*[64][H]*/, const float* h /*[64][H]*/, float* d_direct /*[H][H]*/, uint32_t lbo, uint32_t sbo, uint32_t majors, int layout,
      const __grid_constant__ CUtensorMap tm) {
  extern __shared__ unsigned char raw[];
  float* sm = reinterpret_cast<float*>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
  float* A = sm;                   // [H/4][64][4]
  float* Bm = sm + H * ROWS;
  float* stg = sm + 2 * H * ROWS;  // 4 warps x 1024 floats
  uint64_t* bar = reinterpret_cast<uint64_t*>(stg + 4096);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int i = t; i < ROWS * H; i += blockDim.x) {
    const int r = i / H, c = i % H;
    if (layout == 0) {
      A[((c >> 2) * ROWS + r) * 4 + (c & 3)] = dz[i];
      Bm[((c >> 2) * ROWS + r) * 4 + (c & 3)] = h[i];
    } else {
      const int g = c >> 5, kg = r >> 2, row = r & 3, u = ((c & 31) >> 3) ^ row;
      const int o = ((g * (ROWS / 4) + kg) * 4 + row) * 32 + u * 8 + (c & 7);
      A[o] = dz[i];
      Bm[o] = h[i];
    }
  }
  if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 4 && lane == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | majors | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int kb = 0; kb < ROWS / 8; ++kb) {
      const uint64_t lt = layout == 0 ? 0ull : (1ull << 61);
      const uint32_t step = layout == 0 ? 32u : 256u;         // floats per K step of 8 rows
      const uint64_t ad = umma_desc_kmajor(smem_u32(A + kb * step), lbo, sbo) | lt;
      const uint64_t bd = umma_desc_kmajor(smem_u32(Bm + kb * step), lbo, sbo) | lt;
      umma_tf32(tmem, ad, bd, idesc, kb != 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  if (warp < 4) {
    mbar_wait(bar, 0);
    tc_fence_after();
    float* dst = stg + warp * 1024;
    for (int cb = 0; cb < H; cb += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + cb, v);
      for (int j = 0; j < 32; ++j) d_direct[(32 * warp + lane) * H + cb + j] = v[j];
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(dst + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { tma_reduce_add_2d(&tm, cb, 32 * warp, dst); bulk_commit(); }
    }
    if (lane == 0) bulk_wait_read<0>();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128) : "memory");
}

int main() {
  std::vector<float> dz(ROWS * H), h(ROWS * H), ref(H * H, 0.f);
  for (int i = 0; i < ROWS * H; ++i) { dz[i] = (float)((i * 7 + 3) % 11 - 5); h[i] = (float)((i * 5 + 1) % 13 - 6); }
  for (int n = 0; n < H; ++n) for (int k = 0; k < H; ++k) { float s = 0; for (int b = 0; b < ROWS; ++b) s += dz[b * H + n] * h[b * H + k]; ref[n * H + k] = s; }
  float *d_dz, *d_h, *d_dir, *d_tma;
  cudaMalloc(&d_dz, ROWS * H * 4); cudaMalloc(&d_h, ROWS * H * 4); cudaMalloc(&d_dir, H * H * 4); cudaMalloc(&d_tma, H * H * 4);
  cudaMemcpy(d_dz, dz.data(), ROWS * H * 4, cudaMemcpyHostToDevice); cudaMemcpy(d_h, h.data(), ROWS * H * 4, cudaMemcpyHostToDevice);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                         CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap tm;
  const cuuint64_t dims[2] = {H, H}; const cuuint64_t strides[1] = {H * 4}; const cuuint32_t box[2] = {32, 32}; const cuuint32_t estr[2] = {1, 1};
  CUresult r = ((Fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_tma, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc %d\n", (int)r);
  const size_t smem = (2 * H * ROWS + 4096 + 16) * 4 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct V { uint32_t lbo, sbo, majors; int layout; const char* name; } vs[] = {
      {128, ROWS * 16, (1u << 15) | (1u << 16), 0, "MN/MN interleave lbo=128 sbo=1024"},
      {ROWS * 16, 128, 0, 0, "K/K interleave (sanity: nonzero)"},
      {(ROWS / 4) * 512, 512, (1u << 15) | (1u << 16), 1, "MN/MN SW128_32B lbo=8192 sbo=512"},
      {512, (ROWS / 4) * 512, (1u << 15) | (1u << 16), 1, "MN/MN SW128_32B lbo=512 sbo=8192"},
  };
  for (auto& v : vs) {
    cudaMemset(d_dir, 0, H * H * 4); cudaMemset(d_tma, 0, H * H * 4);
    probe<<<1, 160, smem>>>(d_dz, d_h, d_dir, v.lbo, v.sbo, v.majors, v.layout, tm);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> dir(H * H), tma(H * H);
    cudaMemcpy(dir.data(), d_dir, H * H * 4, cudaMemcpyDeviceToHost); cudaMemcpy(tma.data(), d_tma, H * H * 4, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0, e3 = 0; int nz1 = 0, nz2 = 0;
    for (int i = 0; i < H * H; ++i) { e1 = fmax(e1, fabs(dir[i] - ref[i])); e2 = fmax(e2, fabs(tma[i] - ref[i])); e3 = fmax(e3, fabs(tma[i] - dir[i])); nz1 += dir[i] != 0; nz2 += tma[i] != 0; }
    printf("%s: %s  direct err %.3g (nonzero %d)  tma err %.3g (nonzero %d)  tma-vs-direct %.3g\n", v.name, cudaGetErrorString(e), e1, nz1, e2, nz2, e3);
    printf("  ref[0][0..3] %g %g %g %g | direct %g %g %g %g | tma %g %g %g %g\n", ref[0], ref[1], ref[2], ref[3], dir[0], dir[1], dir[2], dir[3], tma[0], tma[1], tma[2], tma[3]);
    printf("  ref[1][0],ref[0][1] %g %g direct[1][0] %g\n", ref[H], ref[1], dir[H]);
  }
  return 0;
}
