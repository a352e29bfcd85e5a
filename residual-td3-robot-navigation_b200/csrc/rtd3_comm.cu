// NCCL communicator of the data-parallel learner behind the C ABI (`rtd3_allreduce_grads(comm, flat_grads, count, stream)`,
// the one collective of the path; no counterpart in the reference, which is single-process).
// librtd3.so has no link-time dependency on NCCL: the library is resolved at run time - first the copy the host process has already
// loaded (torch ships one), then the system one - so that single-GPU users never need it.  The communicator is OURS (created from a
// unique id the host exchanges over whatever channel it has, e.g. torch.distributed), which is what makes the call capturable in a
// CUDA graph together with the learner kernels: ncclAllReduce on a capturing stream becomes a graph node.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "rtd3_common.cuh"

namespace rtd3 {

// Minimal restatement of the nccl.h declarations used (NCCL 2.x ABI: ncclUniqueId is 128 opaque bytes passed by value).
struct NcclUniqueId { char internal[RTD3_COMM_ID_BYTES]; };
typedef void* NcclComm;
constexpr int kNcclSuccess = 0, kNcclSum = 0, kNcclFloat32 = 7;

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mutex;

static const NcclApi* nccl_api() {
  std::lock_guard<std::mutex> lock(g_nccl_mutex);
  if (g_nccl.handle) return &g_nccl;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy already in the process (torch's), if any
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("NCCL not found: %s", dlerror());
    return nullptr;
  }
  NcclApi api;
  api.handle = h;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
  api.GetVersion = (decltype(api.GetVersion))dlsym(h, "ncclGetVersion");
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.GetErrorString) {
    set_error("libnccl lacks a required symbol");
    return nullptr;
  }
  g_nccl = api;
  return &g_nccl;
}

}  // namespace rtd3

struct rtd3_comm {
  rtd3::NcclComm comm;
  int rank, world, device;
};

using namespace rtd3;

#define RTD3_NCCL(api, expr)                                                               \
  do {                                                                                     \
    const int _r = (expr);                                                                 \
    if (_r != kNcclSuccess) {                                                              \
      ::rtd3::set_error("%s: %s -> NCCL error %d (%s)", __func__, #expr, _r, (api)->GetErrorString(_r)); \
      return 1000 + _r;                                                                    \
    }                                                                                      \
  } while (0)

extern "C" {

int32_t rtd3_comm_nccl_version(void) {
  const NcclApi* api = nccl_api();
  int v = 0;
  if (!api || !api->GetVersion || api->GetVersion(&v) != kNcclSuccess) return -1;
  return v;
}

int32_t rtd3_comm_unique_id(uint8_t* id_out) {
  RTD3_CHECK_ARG(id_out, "null id buffer");
  const NcclApi* api = nccl_api();
  if (!api) return RTD3_ERR_STATE;
  NcclUniqueId id;
  RTD3_NCCL(api, api->GetUniqueId(&id));
  memcpy(id_out, id.internal, RTD3_COMM_ID_BYTES);
  return 0;
}

int32_t rtd3_comm_create(rtd3_comm** out, const uint8_t* id, int32_t rank, int32_t world, int32_t device) {
  RTD3_CHECK_ARG(out && id, "null argument");
  RTD3_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "bad rank / world");
  const NcclApi* api = nccl_api();
  if (!api) return RTD3_ERR_STATE;
  int prev = 0;
  RTD3_CUDA(cudaGetDevice(&prev));
  RTD3_CUDA(cudaSetDevice(device));
  NcclUniqueId uid;
  memcpy(uid.internal, id, RTD3_COMM_ID_BYTES);
  NcclComm c = nullptr;
  const int r = api->CommInitRank(&c, world, uid, rank);
  cudaSetDevice(prev);
  if (r != kNcclSuccess) {
    set_error("rtd3_comm_create: ncclCommInitRank -> NCCL error %d (%s)", r, api->GetErrorString(r));
    return 1000 + r;
  }
  *out = new rtd3_comm{c, rank, world, device};
  return 0;
}

int32_t rtd3_comm_destroy(rtd3_comm* comm) {
  if (!comm) return 0;
  const NcclApi* api = nccl_api();
  if (api && comm->comm) api->CommDestroy(comm->comm);
  delete comm;
  return 0;
}

int32_t rtd3_comm_world(const rtd3_comm* comm) { return comm ? comm->world : -1; }
int32_t rtd3_comm_rank(const rtd3_comm* comm) { return comm ? comm->rank : -1; }

int32_t rtd3_allreduce_grads(rtd3_comm* comm, float* flat_grads, int64_t count, void* stream) {
  RTD3_CHECK_ARG(comm && flat_grads, "null argument");
  RTD3_CHECK_ARG(count >= 0, "negative count");
  if (count == 0 || comm->world == 1) return 0;
  const NcclApi* api = nccl_api();
  if (!api) return RTD3_ERR_STATE;
  RTD3_NCCL(api, api->AllReduce(flat_grads, flat_grads, (size_t)count, kNcclFloat32, kNcclSum, comm->comm, (cudaStream_t)stream));
  count_launch();
  return 0;
}

}  // extern "C"
