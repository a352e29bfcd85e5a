// ReplayBuffer.sample's index draw (robot.py:111): np.random.choice(len, B, replace=False) on a numpy-legacy
// MT19937 stream = the first B entries of a legacy Fisher-Yates shuffle of arange(len) (RandomState.shuffle ->
// random_interval masked rejection).  Bit-exact with numpy; `count` consecutive samples are drawn in one launch
// (a TD3 update needs 150 of them: robot.py:272-285).
#include "rtd3_common.cuh"
#include "rtd3_mt.cuh"

namespace rtd3 {

// v1: the shuffle is inherently serial in the stream; one thread walks it with the MT state and the permutation in
// shared memory, the rest of the warp only helps to (re)initialise the permutation and to write the result.
__global__ void __launch_bounds__(128)
sample_indices_kernel(rtd3_mt_bank b, int64_t stream_id, int32_t n, int32_t batch, int32_t count, int32_t* __restrict__ out,
                      int32_t* __restrict__ perm_global) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* mt = reinterpret_cast<uint32_t*>(smem_raw);                 // [624]
  int32_t* perm = perm_global ? perm_global : reinterpret_cast<int32_t*>(smem_raw + RTD3_MT_N * 4);
  __shared__ int s_pos;
  for (int k = threadIdx.x; k < RTD3_MT_N; k += blockDim.x) mt[k] = b.mt[(int64_t)k * b.n + stream_id];
  if (threadIdx.x == 0) s_pos = b.pos[stream_id];
  __syncthreads();
  for (int c = 0; c < count; ++c) {
    for (int k = threadIdx.x; k < n; k += blockDim.x) perm[k] = k;
    __syncthreads();
    if (threadIdx.x == 0) {
      MtStream s{mt, 1, s_pos};
      for (int i = n - 1; i >= 1; --i) {
        const int j = (int)s.interval((uint32_t)i);
        const int t = perm[i];
        perm[i] = perm[j];
        perm[j] = t;
      }
      s_pos = s.pos;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < batch; k += blockDim.x) out[(int64_t)c * batch + k] = perm[k];
    __syncthreads();
  }
  for (int k = threadIdx.x; k < RTD3_MT_N; k += blockDim.x) b.mt[(int64_t)k * b.n + stream_id] = mt[k];
  if (threadIdx.x == 0) b.pos[stream_id] = s_pos;
}

}  // namespace rtd3

using namespace rtd3;

extern "C" int32_t rtd3_sample_indices_mt19937(const rtd3_mt_bank* bank, int64_t stream_id, int32_t n, int32_t batch, int32_t count,
                                               int32_t* out, int32_t* scratch, void* stream) {
  RTD3_CHECK_ARG(bank && bank->mt && bank->pos && out, "null argument");
  RTD3_CHECK_ARG(stream_id >= 0 && stream_id < bank->n, "stream id out of range");
  RTD3_CHECK_ARG(n >= 1 && batch >= 1 && batch <= n && count >= 0, "need 1 <= batch <= n");
  if (count == 0) return 0;
  const bool in_smem = n <= 48000;
  RTD3_CHECK_ARG(in_smem || scratch, "n > 48000 needs a scratch buffer of n int32");
  static bool attr_set = false;
  if (!attr_set) {
    RTD3_CUDA(cudaFuncSetAttribute(sample_indices_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RTD3_MT_N * 4 + 48000 * 4));
    attr_set = true;
  }
  const size_t smem = RTD3_MT_N * 4 + (in_smem ? (size_t)n * 4 : 0);
  sample_indices_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(*bank, stream_id, n, batch, count, out, in_smem ? nullptr : scratch);
  RTD3_LAUNCHED();
  return 0;
}
