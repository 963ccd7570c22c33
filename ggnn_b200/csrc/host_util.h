// host_util.h -- small host-side helpers shared by the launchers (error reporting, device info).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace g200 {

int set_error(int code, const char* msg);                 // stores a thread-local message, returns code
int set_cuda_error(cudaError_t e, const char* where);     // returns 0 on cudaSuccess

struct DeviceInfo {
  int device;
  uint32_t num_sms;
  uint32_t smem_per_sm;
  uint32_t smem_per_block_optin;
  int cc_major, cc_minor;
};
const DeviceInfo& device_info();  // of the current device (cached per device)

uint32_t env_u32(const char* name, uint32_t dflt);

// CUtensorMap (128 bytes, 64-byte aligned) of a row-major fp32 matrix [rows, cols] whose box is one whole row:
// the operand of TMA tile::gather4 row gathers.  Returns 0 or an error code (message set).
struct alignas(64) TensorMapStorage {
  unsigned char bytes[128];
};
int make_row_gather_tensor_map(TensorMapStorage* out, const float* d_base, uint64_t rows, uint32_t cols);
// the same for uint8 rows of `cols` bytes (cols % 16 == 0, cols <= 256)
int make_row_gather_tensor_map_u8(TensorMapStorage* out, const uint8_t* d_base, uint64_t rows, uint32_t cols);

inline uint32_t bit_ceil_u32(uint32_t v)
{
  if (v <= 1) return 1;
  v--;
  v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
  return v + 1;
}
inline uint32_t next_multiple32(uint32_t v) { return v % 32 == 0 ? v : 32 * (v / 32 + 1); }
inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }
inline size_t align8(size_t s) { return (s + 7) / 8 * 8; }

}  // namespace g200
