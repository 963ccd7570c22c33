#include "host_util.h"
#include "../../include/ggnn_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace g200 {

static thread_local char g_err[512] = "";

int set_error(int code, const char* msg)
{
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

int set_cuda_error(cudaError_t e, const char* where)
{
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
  return static_cast<int>(e);
}

const DeviceInfo& device_info()
{
  static DeviceInfo infos[64];
  static bool have[64] = {};
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  std::lock_guard<std::mutex> lock(mu);
  if (!have[dev]) {
    DeviceInfo d{};
    d.device = dev;
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    d.num_sms = v > 0 ? v : 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    d.smem_per_sm = v > 0 ? v : 233472;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    d.smem_per_block_optin = v > 0 ? v : 232448;
    cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
    infos[dev] = d;
    have[dev] = true;
  }
  return infos[dev];
}

uint32_t env_u32(const char* name, uint32_t dflt)
{
  const char* s = getenv(name);
  if (!s || !*s) return dflt;
  return static_cast<uint32_t>(strtoul(s, nullptr, 10));
}

}  // namespace g200

extern "C" const char* ggnn_b200_last_error(void) { return g200::g_err; }
extern "C" const char* ggnn_b200_version(void) { return "ggnn_b200 0.1 (sm_100a; cp.async.bulk staged traversal)"; }
