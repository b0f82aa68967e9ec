// Peer-memory all-gather for the sharded reSGLD exchange.
//
// The payload of one exchange is tiny (2 floats per system per replica: 32 KB
// per rank at 4096 systems), so ncclAllGather is pure latency (15-20 us on 8
// GPUs).  Here every rank owns a *window* in its HBM,
//     [2 parities][R ranks][bytes_per_rank]  data   +   [2][R] u32 flags,
// exported once with cudaIpcGetMemHandle and mapped by all peers.  One kernel
// per exchange: block d stores this rank's rows straight into rank d's window
// over NVLink / NVSwitch, fences system-wide, posts the sequence number into
// rank d's flag slot for this rank, and then waits for rank d's flag in the
// LOCAL window -- when the kernel retires, all R rows of this exchange are in
// local HBM.  Two parities make the window safe to refill: a peer can be at
// most one exchange ahead (it waited for our flag of the previous one, which
// we posted after our previous decision kernels in stream order).
#include "common.cuh"

namespace sgmc {

__device__ unsigned int g_p2p_timeouts = 0;

struct P2pArgs {
  uint8_t* const* peers;   // device array of R window base pointers (own included)
  int rank, R;
  const uint8_t* send;
  uint32_t bytes;          // per rank, multiple of 16
  uint32_t parity, seq;
  unsigned long long spin_limit;   // clock64 cycles before giving up
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(256) k_p2p_allgather(const P2pArgs a) {
  pdl_launch_dependents();
  pdl_wait();                                             // the rows to send are final
  const int d = blockIdx.x;                               // destination / source rank
  const size_t data_bytes = (size_t)2 * a.R * a.bytes;
  uint8_t* win = a.peers[d];
  uint4* dst = reinterpret_cast<uint4*>(win + ((size_t)a.parity * a.R + a.rank) * a.bytes);
  const uint4* src = reinterpret_cast<const uint4*>(a.send);
  for (uint32_t i = threadIdx.x; i < a.bytes / 16; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t* flags_d = reinterpret_cast<uint32_t*>(win + data_bytes);
    st_release_sys(flags_d + a.parity * a.R + a.rank, a.seq);          // "rank's rows are in"
    const uint32_t* mine =
        reinterpret_cast<const uint32_t*>(a.peers[a.rank] + data_bytes) + a.parity * a.R + d;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) != a.seq) {
      if ((unsigned long long)(clock64() - t0) > a.spin_limit) {       // a peer is gone
        atomicAdd(&g_p2p_timeouts, 1u);
        break;
      }
    }
  }
}

}  // namespace sgmc

using namespace sgmc;

extern "C" {

size_t sgmc_p2p_window_bytes(int n_ranks, size_t bytes_per_rank) {
  return (size_t)2 * n_ranks * bytes_per_rank + (size_t)2 * n_ranks * 4;
}

int sgmc_p2p_export(const void* window, void* handle_64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  return check_cuda(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle_64, (void*)window),
                    "cudaIpcGetMemHandle");
}

int sgmc_p2p_open(const void* handle_64, void** peer_window) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_64, sizeof(h));
  return check_cuda(cudaIpcOpenMemHandle(peer_window, h, cudaIpcMemLazyEnablePeerAccess),
                    "cudaIpcOpenMemHandle");
}

int sgmc_p2p_close(void* peer_window) {
  return check_cuda(cudaIpcCloseMemHandle(peer_window), "cudaIpcCloseMemHandle");
}

int sgmc_p2p_allgather(void* stream, void* const* peer_windows_dev, int rank, int n_ranks,
                       const void* send, size_t bytes_per_rank, unsigned int seq) {
  SGMC_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "bad rank");
  SGMC_REQUIRE(bytes_per_rank % 16 == 0 && bytes_per_rank < (1u << 31), "bytes_per_rank %% 16");
  SGMC_REQUIRE(seq != 0, "sequence numbers start at 1");
  P2pArgs a{(uint8_t* const*)peer_windows_dev, rank, n_ranks, (const uint8_t*)send,
            (uint32_t)bytes_per_rank, seq & 1u, seq, 40000000000ull /* ~20 s */};
  launch_pdl(k_p2p_allgather, dim3(n_ranks), dim3(256), 0, (cudaStream_t)stream, a);
  return post_launch("sgmc_p2p_allgather");
}

int sgmc_p2p_timeouts(unsigned int* count) {
  return check_cuda(cudaMemcpyFromSymbol(count, g_p2p_timeouts, sizeof(unsigned int)),
                    "cudaMemcpyFromSymbol");
}

}  // extern "C"
