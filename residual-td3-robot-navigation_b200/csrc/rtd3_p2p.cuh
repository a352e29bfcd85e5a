// Device-side pieces of the peer-memory exchange (rtd3_p2p.cu: the all-reduce kernels; rtd3_td3.cu: the weight-gradient kernel that
// exchanges its own tiles): system-scope flag stores / loads, the bounded spin, loads that do not come from a stale L1 line.
#pragma once
#include "rtd3_common.cuh"

namespace rtd3 {

constexpr int kP2pMaxWorld = RTD3_P2P_MAX_WORLD;
constexpr unsigned long long kSpinLimitNs = 20ull * 1000 * 1000 * 1000;

struct P2pPeers {
  float* recv[kP2pMaxWorld];                 // rank q's receive area as mapped here: [2][W][count] floats
  unsigned long long* flags[kP2pMaxWorld];   // rank q's flag array: flags[r] = last step rank r has pushed completely
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void wait_flag(const unsigned long long* p, unsigned long long seq) {
  if (ld_acquire_sys(p) >= seq) return;
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(p) < seq)
    if (global_ns() - t0 > kSpinLimitNs) asm volatile("trap;");
}
__device__ __forceinline__ float4 ld_fresh(const float4* p) {   // written by a peer during this launch: not from a stale line
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ float ld_fresh1(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// One flag per (rank, block) behind the 16 whole-kernel words of a rank's flag array: block b of every rank owns the same elements
// in the push and in the reduction, raises its own flag right after its own pushes and waits only for block b of its peers.
constexpr int kP2pBlockFlags = 256;          // per rank; a grid that hands over block-locally never has more blocks
constexpr int kP2pBlockFlagBase = 16;

}  // namespace rtd3
