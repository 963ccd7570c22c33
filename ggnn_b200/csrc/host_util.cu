#include "host_util.h"

#include <cuda.h>
#include "../../include/ggnn_b200.h"

#include <cstdio>
#include <cstdlib>
#include <algorithm>
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

static int make_row_gather_tensor_map_any(TensorMapStorage* out, const void* d_base, uint64_t rows, uint32_t cols, bool u8);
int make_row_gather_tensor_map(TensorMapStorage* out, const float* d_base, uint64_t rows, uint32_t cols)
{
  return make_row_gather_tensor_map_any(out, d_base, rows, cols, false);
}
int make_row_gather_tensor_map_u8(TensorMapStorage* out, const uint8_t* d_base, uint64_t rows, uint32_t cols)
{
  return make_row_gather_tensor_map_any(out, d_base, rows, cols, true);
}
static int make_row_gather_tensor_map_any(TensorMapStorage* out, const void* d_base, uint64_t rows, uint32_t cols, bool u8)
{
  static_assert(sizeof(CUtensorMap) == sizeof(TensorMapStorage), "CUtensorMap is 128 bytes");
  using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<EncodeFn>(fn);
  });
  if (!encode) return set_error(GGNN_B200_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  if (!u8 && (cols == 0 || cols > 256 || cols % 4)) return set_error(GGNN_B200_ERR_UNSUPPORTED, "row gather tensor map needs D % 4 == 0 and D <= 256");
  if (u8 && (cols == 0 || cols > 256 || cols % 16)) return set_error(GGNN_B200_ERR_UNSUPPORTED, "uint8 row gather tensor map needs D % 16 == 0 and D <= 256");
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * (u8 ? 1 : sizeof(float))};
  const cuuint32_t box[2] = {cols, 1};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUresult r = encode(reinterpret_cast<CUtensorMap*>(out), u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(d_base), dims,
                            strides, box, elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[96];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
    return set_error(GGNN_B200_ERR_UNSUPPORTED, msg);
  }
  return 0;
}

uint32_t env_u32(const char* name, uint32_t dflt)
{
  const char* s = getenv(name);
  if (!s || !*s) return dflt;
  return static_cast<uint32_t>(strtoul(s, nullptr, 10));
}

}  // namespace g200

extern "C" const char* ggnn_b200_last_error(void) { return g200::g_err; }
extern "C" const char* ggnn_b200_version(void) { return "ggnn_b200 0.2 (sm_100a; TMA gather4 staged traversal, native uint8 rows, fused shard-merge exchange, tcgen05 brute force)"; }

// uint8 -> fp32 widening (see ggnn_b200.h); 16 values per thread: one 16-byte load, four 16-byte stores
namespace g200 {
__global__ void __launch_bounds__(256) widen_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t count)
{
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t n16 = count / 16;
  const bool aligned = (reinterpret_cast<uintptr_t>(src) % 16 == 0) && (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
  size_t done = 0;
  if (aligned) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
      const uint4 v = reinterpret_cast<const uint4*>(src)[i];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      float4* o = reinterpret_cast<float4*>(dst) + 4 * i;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o[j] = make_float4(static_cast<float>(w[j] & 0xffu), static_cast<float>((w[j] >> 8) & 0xffu),
                           static_cast<float>((w[j] >> 16) & 0xffu), static_cast<float>(w[j] >> 24));
    }
    done = n16 * 16;
  }
  for (size_t i = done + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride)
    dst[i] = static_cast<float>(src[i]);
}
}  // namespace g200

extern "C" int ggnn_b200_widen_u8(const uint8_t* d_src, float* d_dst, size_t count, ggnn_b200_stream_t stream_)
{
  if (!count) return 0;
  if (!d_src || !d_dst) return g200::set_error(GGNN_B200_ERR_INVALID, "null device pointer");
  const g200::DeviceInfo& dev = g200::device_info();
  const size_t want = (count / 16 + 255) / 256 + 1;
  const unsigned grid = static_cast<unsigned>(std::min<size_t>(want, static_cast<size_t>(dev.num_sms) * 8));
  g200::widen_u8_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(d_src, d_dst, count);
  return g200::set_cuda_error(cudaGetLastError(), "widen_u8_kernel launch");
}

// interleaved copy of an fp32 base (see ggnn_b200.h): one warp per row, coalesced loads, D/32 adjacent values per lane
namespace g200 {
__global__ void __launch_bounds__(256) interleave_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, uint32_t N, uint32_t D32)
{
  const uint32_t lane = threadIdx.x & 31;
  for (size_t n = static_cast<size_t>(blockIdx.x) * 8 + (threadIdx.x >> 5); n < N; n += static_cast<size_t>(gridDim.x) * 8) {
    const float* r = src + n * 32 * D32;
    float* w = dst + n * 32 * D32 + static_cast<size_t>(D32) * lane;
    for (uint32_t c = 0; c < D32; ++c) w[c] = r[32 * c + lane];
  }
}
}  // namespace g200

extern "C" int ggnn_b200_interleave_rows(const float* d_src, float* d_dst, uint32_t N, uint32_t D, ggnn_b200_stream_t stream_)
{
  if (!d_src || !d_dst) return g200::set_error(GGNN_B200_ERR_INVALID, "null device pointer");
  if (D == 0 || D % 32 || D > 128) return g200::set_error(GGNN_B200_ERR_UNSUPPORTED, "interleaved rows need D in {32, 64, 96, 128}");
  if (!N) return 0;
  const unsigned grid = static_cast<unsigned>(std::min<size_t>((static_cast<size_t>(N) + 7) / 8, static_cast<size_t>(g200::device_info().num_sms) * 16));
  g200::interleave_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(d_src, d_dst, N, D / 32);
  return g200::set_cuda_error(cudaGetLastError(), "interleave_rows_kernel launch");
}
