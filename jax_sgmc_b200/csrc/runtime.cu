// Runtime plumbing of libsgmc_b200: error state, device memory, streams,
// events.  Thin wrappers over the CUDA runtime so that a host written in
// Python (ctypes) or C needs no other CUDA binding.  Not on the hot path.
#include "common.cuh"

#include <cstring>

namespace sgmc {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

static std::atomic<int> g_options[SGMC_OPT_COUNT];
int option(int which) {
  return (which >= 0 && which < SGMC_OPT_COUNT) ? g_options[which].load() : 0;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace sgmc

using namespace sgmc;

extern "C" {

const char* sgmc_last_error(void) { return g_err; }
int sgmc_version(void) { return 100; }
int sgmc_set_option(int option, int value) {
  SGMC_REQUIRE(option >= 0 && option < SGMC_OPT_COUNT, "unknown option %d", option);
  g_options[option].store(value);
  return 0;
}
int sgmc_get_option(int option) { return sgmc::option(option); }
unsigned long long sgmc_launch_count(void) { return g_launches.load(); }

int sgmc_device_count(int* count) {
  return check_cuda(cudaGetDeviceCount(count), "cudaGetDeviceCount");
}
int sgmc_set_device(int device) {
  return check_cuda(cudaSetDevice(device), "cudaSetDevice");
}
int sgmc_device_info(int device, int* sms, int* major, int* minor,
                     size_t* total_mem) {
  cudaDeviceProp p;
  if (check_cuda(cudaGetDeviceProperties(&p, device), "cudaGetDeviceProperties"))
    return 1;
  if (sms) *sms = p.multiProcessorCount;
  if (major) *major = p.major;
  if (minor) *minor = p.minor;
  if (total_mem) *total_mem = p.totalGlobalMem;
  return 0;
}
int sgmc_malloc(void** dptr, size_t bytes) {
  return check_cuda(cudaMalloc(dptr, bytes ? bytes : 1), "cudaMalloc");
}
int sgmc_free(void* dptr) { return check_cuda(cudaFree(dptr), "cudaFree"); }
int sgmc_host_alloc(void** hptr, size_t bytes) {
  return check_cuda(cudaMallocHost(hptr, bytes ? bytes : 1), "cudaMallocHost");
}
int sgmc_host_free(void* hptr) {
  return check_cuda(cudaFreeHost(hptr), "cudaFreeHost");
}
int sgmc_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
  return check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice,
                                    (cudaStream_t)stream), "memcpy h2d");
}
int sgmc_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
  return check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost,
                                    (cudaStream_t)stream), "memcpy d2h");
}
int sgmc_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
  return check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice,
                                    (cudaStream_t)stream), "memcpy d2d");
}
int sgmc_memset(void* dst, int value, size_t bytes, void* stream) {
  return check_cuda(cudaMemsetAsync(dst, value, bytes, (cudaStream_t)stream),
                    "memset");
}
int sgmc_stream_create(void** stream) {
  cudaStream_t s;
  if (check_cuda(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking),
                 "cudaStreamCreate"))
    return 1;
  *stream = (void*)s;
  return 0;
}
int sgmc_stream_create_high_priority(void** stream) {
  int lo = 0, hi = 0;
  if (check_cuda(cudaDeviceGetStreamPriorityRange(&lo, &hi), "cudaDeviceGetStreamPriorityRange"))
    return 1;
  cudaStream_t s;
  if (check_cuda(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi),
                 "cudaStreamCreateWithPriority"))
    return 1;
  *stream = (void*)s;
  return 0;
}
int sgmc_stream_destroy(void* stream) {
  return check_cuda(cudaStreamDestroy((cudaStream_t)stream), "cudaStreamDestroy");
}
int sgmc_stream_sync(void* stream) {
  return check_cuda(cudaStreamSynchronize((cudaStream_t)stream),
                    "cudaStreamSynchronize");
}
int sgmc_device_sync(void) {
  return check_cuda(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
}
int sgmc_event_create(void** event) {
  cudaEvent_t e;
  if (check_cuda(cudaEventCreate(&e), "cudaEventCreate")) return 1;
  *event = (void*)e;
  return 0;
}
int sgmc_event_destroy(void* event) {
  return check_cuda(cudaEventDestroy((cudaEvent_t)event), "cudaEventDestroy");
}
int sgmc_event_record(void* event, void* stream) {
  return check_cuda(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream),
                    "cudaEventRecord");
}
int sgmc_event_sync(void* event) {
  return check_cuda(cudaEventSynchronize((cudaEvent_t)event),
                    "cudaEventSynchronize");
}
int sgmc_stream_wait_event(void* stream, void* event) {
  return check_cuda(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0),
                    "cudaStreamWaitEvent");
}
int sgmc_event_elapsed_ms(void* start, void* stop, float* ms) {
  return check_cuda(
      cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop),
      "cudaEventElapsedTime");
}

}  // extern "C"
