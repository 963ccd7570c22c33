// exchange.cu -- peer-memory plumbing of the fused shard-merge exchange (see ggnn_b200.h): the query kernel's epilogue
// stores every query's top-K list straight into the gathered buffer of every destination GPU and bumps a flag word there;
// the receiving side waits for the flag on its own stream and merges the lists in place (merge_topk.cu).
// The reference has no device-side exchange: every GPU copies its lists to the host and the CPU heap-merges them
// (src/ggnn/base/gpu_instance.cu:714-742, src/ggnn/base/result_merger.cpp:51-149).
#include "common.cuh"
#include "host_util.h"
#include "../../include/ggnn_b200.h"

#include <cstring>

namespace g200 {

__global__ void __launch_bounds__(32) wait_flag_kernel(const uint32_t* flag, uint32_t expected, unsigned long long timeout_ns,
                                                       uint32_t* timed_out)
{
  if (threadIdx.x != 0) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (true) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (static_cast<int32_t>(v - expected) >= 0) return;  // wrap-safe ">="
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) {
      if (timed_out) *timed_out = 1u;
      return;
    }
    __nanosleep(200);
  }
}

}  // namespace g200

using namespace g200;

extern "C" int ggnn_b200_ipc_alloc(size_t bytes, void** d_ptr, unsigned char* handle64)
{
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!d_ptr || !bytes) return set_error(GGNN_B200_ERR_INVALID, "bad argument");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaMalloc(exchange buffer)");
  e = cudaMemset(p, 0, bytes);
  if (e != cudaSuccess) {
    cudaFree(p);
    return set_cuda_error(e, "cudaMemset(exchange buffer)");
  }
  if (handle64) {
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
      cudaFree(p);
      return set_cuda_error(e, "cudaIpcGetMemHandle");
    }
    memcpy(handle64, &h, sizeof(h));
  }
  *d_ptr = p;
  return 0;
}

extern "C" int ggnn_b200_ipc_open(const unsigned char* handle64, void** d_ptr)
{
  if (!handle64 || !d_ptr) return set_error(GGNN_B200_ERR_INVALID, "bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  return set_cuda_error(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}

extern "C" int ggnn_b200_ipc_close(void* d_ptr) { return set_cuda_error(cudaIpcCloseMemHandle(d_ptr), "cudaIpcCloseMemHandle"); }

extern "C" int ggnn_b200_ipc_free(void* d_ptr) { return set_cuda_error(cudaFree(d_ptr), "cudaFree(exchange buffer)"); }

extern "C" int ggnn_b200_peer_enable(int peer_device)
{
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev == peer_device) return 0;
  int can = 0;
  cudaError_t e = cudaDeviceCanAccessPeer(&can, dev, peer_device);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaDeviceCanAccessPeer");
  if (!can) return set_error(GGNN_B200_ERR_UNSUPPORTED, "no peer access between these devices");
  e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return 0;
  }
  return set_cuda_error(e, "cudaDeviceEnablePeerAccess");
}

extern "C" int ggnn_b200_wait_flag(const uint32_t* d_flag, uint32_t expected, uint32_t timeout_ms, uint32_t* d_timed_out,
                                   ggnn_b200_stream_t stream_)
{
  if (!d_flag) return set_error(GGNN_B200_ERR_INVALID, "null flag pointer");
  const unsigned long long ns = (timeout_ms ? timeout_ms : 10000u) * 1000000ull;
  wait_flag_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream_)>>>(d_flag, expected, ns, d_timed_out);
  return set_cuda_error(cudaGetLastError(), "wait_flag_kernel launch");
}
