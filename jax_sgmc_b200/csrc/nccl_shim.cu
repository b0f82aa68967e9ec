// NCCL entry points of libsgmc_b200 (reSGLD replica exchange, sharded-gradient
// all-reduce).  libnccl.so.2 is opened lazily with dlopen so the library has
// no link-time NCCL dependency and shares the already-loaded NCCL when the
// host process has one.
#include "common.cuh"

#include <dlfcn.h>
#include <mutex>

namespace sgmc {

typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
typedef int NcclResult;

struct NcclApi {
  void* handle = nullptr;
  NcclResult (*GetUniqueId)(NcclUniqueId*) = nullptr;
  NcclResult (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  NcclResult (*CommDestroy)(NcclComm) = nullptr;
  NcclResult (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  NcclResult (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(NcclResult) = nullptr;
  bool ok = false;
};

static NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) return;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
    a.AllReduce = (decltype(a.AllReduce))dlsym(a.handle, "ncclAllReduce");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather &&
           a.AllReduce && a.GetErrorString;
  });
  return a;
}

static int check_nccl(NcclResult r, const char* what) {
  if (r != 0) {
    set_error("%s: %s", what, api().GetErrorString ? api().GetErrorString(r) : "?");
    return 1;
  }
  return 0;
}

}  // namespace sgmc

using namespace sgmc;

extern "C" {

int sgmc_nccl_available(void) { return api().ok ? 1 : 0; }

int sgmc_nccl_unique_id(void* unique_id_128) {
  SGMC_REQUIRE(api().ok, "libnccl.so.2 not found");
  return check_nccl(api().GetUniqueId((NcclUniqueId*)unique_id_128),
                    "ncclGetUniqueId");
}

int sgmc_nccl_init(void** comm, const void* unique_id_128, int n_ranks, int rank) {
  SGMC_REQUIRE(api().ok, "libnccl.so.2 not found");
  NcclUniqueId id;
  memcpy(&id, unique_id_128, sizeof(id));
  NcclComm c = nullptr;
  if (check_nccl(api().CommInitRank(&c, n_ranks, id, rank), "ncclCommInitRank"))
    return 1;
  *comm = c;
  return 0;
}

int sgmc_nccl_destroy(void* comm) {
  SGMC_REQUIRE(api().ok, "libnccl.so.2 not found");
  return check_nccl(api().CommDestroy((NcclComm)comm), "ncclCommDestroy");
}

int sgmc_nccl_allgather(void* comm, void* stream, const void* send, void* recv,
                        size_t bytes_per_rank) {
  SGMC_REQUIRE(api().ok, "libnccl.so.2 not found");
  // ncclInt8 = 0
  return check_nccl(api().AllGather(send, recv, bytes_per_rank, 0, (NcclComm)comm,
                                    (cudaStream_t)stream), "ncclAllGather");
}

// One reSGLD exchange of the sharded ladder in a single call: snapshot of the
// local (U, var) rows on the sampling stream, all-gather + decision kernels on
// the exchange stream (or everything on the sampling stream when x_stream is
// NULL).  comm == NULL: single rank, the gather is a device copy.
int sgmc_resgld_sharded_exchange(void* comm, void* main_stream, void* x_stream,
                                 void* ready_event, void* done_event, const float* uv,
                                 float* uv_send, size_t uv_bytes, float* gathered,
                                 int32_t* holder, float* ssq, const float* F,
                                 const float* temps, const uint32_t* keys_in,
                                 uint32_t* keys_out, int32_t* exchange, int n_replicas,
                                 int64_t n_systems, int64_t step, int first_local_replica,
                                 int n_local_replicas, float* temp_per_chain,
                                 int32_t* temp_index, int prng_layout) {
  cudaStream_t ms = (cudaStream_t)main_stream, xs = (cudaStream_t)x_stream;
  const float* src = uv;
  cudaStream_t s = ms;
  if (xs != nullptr) {
    SGMC_REQUIRE(ready_event && done_event && uv_send, "overlapped exchange needs events");
    // the next potential overwrites (U, var): keep a copy for the exchange stream
    if (check_cuda(cudaMemcpyAsync(uv_send, uv, uv_bytes, cudaMemcpyDeviceToDevice, ms),
                   "snapshot")) return 1;
    if (check_cuda(cudaEventRecord((cudaEvent_t)ready_event, ms), "cudaEventRecord")) return 1;
    if (check_cuda(cudaStreamWaitEvent(xs, (cudaEvent_t)ready_event, 0), "cudaStreamWaitEvent"))
      return 1;
    src = uv_send;
    s = xs;
  }
  if (comm != nullptr) {
    if (int e = sgmc_nccl_allgather(comm, s, src, gathered, uv_bytes)) return e;
  } else if (check_cuda(cudaMemcpyAsync(gathered, src, uv_bytes, cudaMemcpyDeviceToDevice, s),
                        "gather copy")) {
    return 1;
  }
  if (int e = sgmc_resgld_ladder_step(s, gathered, holder, ssq, F, temps, keys_in, keys_out,
                                      exchange, n_replicas, n_systems, step,
                                      first_local_replica, n_local_replicas, temp_per_chain,
                                      temp_index, prng_layout))
    return e;
  if (xs != nullptr)
    return check_cuda(cudaEventRecord((cudaEvent_t)done_event, xs), "cudaEventRecord");
  return 0;
}

int sgmc_nccl_allreduce_sum_f32(void* comm, void* stream, const float* send,
                                float* recv, size_t count) {
  SGMC_REQUIRE(api().ok, "libnccl.so.2 not found");
  // ncclFloat32 = 7, ncclSum = 0
  return check_nccl(api().AllReduce(send, recv, count, 7, 0, (NcclComm)comm,
                                    (cudaStream_t)stream), "ncclAllReduce");
}

}  // extern "C"
