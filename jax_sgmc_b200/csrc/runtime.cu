// Runtime plumbing of libsgmc_b200: error state, device memory, streams,
// events.  Thin wrappers over the CUDA runtime so that a host written in
// Python (ctypes) or C needs no other CUDA binding.  Not on the hot path.
#include "common.cuh"

#include <algorithm>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <cstring>
#include <thread>
#include <vector>

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

namespace sgmc {

// Per-kernel timing of sgmc_glm_sgld_step (SGMC_OPT_STEP_PROFILE): CUDA events on the
// launching stream after each of the step's launches; the intervals are accumulated
// when the next step begins (which synchronises on the previous step -- profile mode
// only).  Marks: 0 step entry, 1 after the operand preparation, 2 after the potential
// kernel, 3 after the update.
static cudaEvent_t g_prof_ev[4];
static bool g_prof_init = false, g_prof_pending = false;
static double g_prof_acc[3] = {0, 0, 0};
static long long g_prof_steps = 0;

static void prof_fold() {
  if (!g_prof_pending) return;
  if (cudaEventSynchronize(g_prof_ev[3]) == cudaSuccess) {
    for (int i = 0; i < 3; ++i) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, g_prof_ev[i], g_prof_ev[i + 1]) == cudaSuccess)
        g_prof_acc[i] += (double)ms * 1e3;
    }
    ++g_prof_steps;
  }
  g_prof_pending = false;
}

void prof_mark(cudaStream_t stream, int slot) {
  if (!option(SGMC_OPT_STEP_PROFILE)) return;
  if (!g_prof_init) {
    for (int i = 0; i < 4; ++i) cudaEventCreate(&g_prof_ev[i]);
    g_prof_init = true;
  }
  if (slot == 0) prof_fold();
  cudaEventRecord(g_prof_ev[slot], stream);
  if (slot == 3) g_prof_pending = true;
}

}  // namespace sgmc

using namespace sgmc;

extern "C" {

// us per step spent in {operand preparation, potential kernel, update} since the last
// reset, and the number of steps folded; reset != 0 clears the accumulators.
int sgmc_debug_step_profile(double* us3, long long* steps, int reset) {
  prof_fold();
  for (int i = 0; i < 3; ++i) us3[i] = g_prof_steps ? g_prof_acc[i] / (double)g_prof_steps : 0.0;
  if (steps) *steps = g_prof_steps;
  if (reset) {
    for (int i = 0; i < 3; ++i) g_prof_acc[i] = 0;
    g_prof_steps = 0;
  }
  return 0;
}

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
int sgmc_host_alloc_wc(void** hptr, size_t bytes) {
  return check_cuda(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocWriteCombined),
                    "cudaHostAlloc");
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
// Host side of the streaming data loader (data/core.py:664-791: the reference refills
// its cache with NumPy fancy indexing inside an io_callback): gather `n_batches`
// minibatches of rows [row0, row0 + rows) of each index row into the staging layout
// sgmc_glm_sgld_scan_host consumes -- per batch `rows` feature rows of d floats, then
// all n labels -- with `n_threads` host threads (memcpy of whole rows; the staging
// buffer is page-locked memory the H2D copies read directly).
// One feature row -> staging.  The destination is page-locked memory only the DMA engine
// reads, so it is written with non-temporal stores (no read-for-ownership of the
// destination lines); the NEXT row of the gather is prefetched while this one streams
// (rows are scattered over a multi-GB array: every row starts with a TLB + DRAM miss).
static inline void copy_row_nt(float* dst, const float* src, const float* next, int64_t d) {
#if defined(__SSE2__)
  if (((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)(d * 4)) & 15u) == 0) {
    const int64_t bytes = d * 4;
    if (next != nullptr)
      for (int64_t o = 0; o < bytes; o += 64)
        __builtin_prefetch(reinterpret_cast<const char*>(next) + o, 0, 0);
    const __m128i* s = reinterpret_cast<const __m128i*>(src);
    __m128i* t = reinterpret_cast<__m128i*>(dst);
    const int64_t q = bytes / 16;
    int64_t i = 0;
    for (; i + 4 <= q; i += 4) {
      const __m128i a = _mm_loadu_si128(s + i), b = _mm_loadu_si128(s + i + 1);
      const __m128i c = _mm_loadu_si128(s + i + 2), e = _mm_loadu_si128(s + i + 3);
      _mm_stream_si128(t + i, a);
      _mm_stream_si128(t + i + 1, b);
      _mm_stream_si128(t + i + 2, c);
      _mm_stream_si128(t + i + 3, e);
    }
    for (; i < q; ++i) _mm_stream_si128(t + i, _mm_loadu_si128(s + i));
    return;
  }
#endif
  (void)next;
  std::memcpy(dst, src, (size_t)d * sizeof(float));
}

int sgmc_host_gather_batches(float* dst, const float* X, const float* y, const int32_t* idx,
                             int64_t n_batches, int64_t n, int64_t d, int64_t row0,
                             int64_t rows, int n_threads) {
  SGMC_REQUIRE(dst && X && y && idx && n_batches >= 0 && n > 0 && d > 0 && row0 >= 0 &&
               rows >= 0 && row0 + rows <= n, "bad gather arguments");
  const int64_t stride = rows * d + n;
  const int64_t items = n_batches * rows;              // one item = one feature row
  const int T = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, items / 64 + 1));
  auto work = [&](int t) {
    const int64_t lo = items * t / T, hi = items * (t + 1) / T;
    for (int64_t it = lo; it < hi; ++it) {
      const int64_t b = it / rows, r = it - b * rows;
      const int64_t src = (int64_t)idx[b * n + row0 + r];
      const float* next = nullptr;
      if (it + 1 < hi) {
        const int64_t b1 = (it + 1) / rows, r1 = it + 1 - b1 * rows;
        next = X + (int64_t)idx[b1 * n + row0 + r1] * d;
      }
      copy_row_nt(dst + b * stride + r * d, X + src * d, next, d);
    }
#if defined(__SSE2__)
    _mm_sfence();
#endif
    for (int64_t b = n_batches * t / T; b < n_batches * (t + 1) / T; ++b) {
      float* lab = dst + b * stride + rows * d;
      for (int64_t i = 0; i < n; ++i) lab[i] = y[idx[b * n + i]];
    }
  };
  if (T == 1) {
    work(0);
    return 0;
  }
  std::vector<std::thread> pool;
  pool.reserve(T - 1);
  for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  return 0;
}

int sgmc_event_elapsed_ms(void* start, void* stop, float* ms) {
  return check_cuda(
      cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop),
      "cudaEventElapsedTime");
}

}  // extern "C"
