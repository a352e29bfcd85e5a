// Error string, version and launch counter of librtd3.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

#include "rtd3_common.cuh"

namespace rtd3 {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::mutex g_dev_mutex;
static std::map<std::pair<const void*, int>, size_t> g_smem_attr;   // (kernel, device) -> bytes already granted
static std::map<int, int> g_num_sms;

cudaError_t ensure_dyn_smem(const void* kernel, size_t bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  size_t& have = g_smem_attr[std::make_pair(kernel, dev)];
  if (bytes <= have) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) have = bytes;
  return e;
}

cudaError_t current_num_sms(int* out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  auto it = g_num_sms.find(dev);
  if (it == g_num_sms.end()) {
    int n = 0;
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    it = g_num_sms.emplace(dev, n).first;
  }
  *out = it->second;
  return cudaSuccess;
}
}  // namespace rtd3

extern "C" {
int32_t rtd3_version(void) { return RTD3_VERSION; }
const char* rtd3_last_error(void) { return rtd3::g_err; }
int64_t rtd3_launch_count(void) { return rtd3::g_launches.load(); }
void rtd3_launch_count_reset(void) { rtd3::g_launches.store(0); }
void rtd3_launch_count_add(int64_t n) { rtd3::g_launches.fetch_add(n); }
}
