// Error string, version and launch counter of librtd3.
#include <atomic>
#include <cstdarg>
#include <cstring>

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
}  // namespace rtd3

extern "C" {
int32_t rtd3_version(void) { return RTD3_VERSION; }
const char* rtd3_last_error(void) { return rtd3::g_err; }
int64_t rtd3_launch_count(void) { return rtd3::g_launches.load(); }
void rtd3_launch_count_reset(void) { rtd3::g_launches.store(0); }
void rtd3_launch_count_add(int64_t n) { rtd3::g_launches.fetch_add(n); }
}
