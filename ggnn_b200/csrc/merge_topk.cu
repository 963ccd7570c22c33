// merge_topk.cu -- k-way merge of per-shard / per-GPU sorted result lists on the device.
// Replaces cub::DeviceSegmentedRadixSort in GPUInstance::sortQueryResults
// (src/ggnn/base/gpu_instance.cu:745-790) and the CPU heap merge ResultMerger::merge
// (src/ggnn/base/result_merger.cpp:51-149): the inputs are already sorted runs, so one pass suffices.
#include "common.cuh"
#include "host_util.h"
#include "../../include/ggnn_b200.h"

namespace g200 {

constexpr int MERGE_MAX_LISTS = 64;

__global__ void __launch_bounds__(128) merge_topk_kernel(const int32_t* __restrict__ ids, const float* __restrict__ dists,
                                                         uint32_t n_lists, size_t list_stride, size_t query_stride,
                                                         uint32_t K_in, uint32_t N_query, uint32_t K,
                                                         long long id_offset_per_list, int32_t* __restrict__ out_ids,
                                                         float* __restrict__ out_dists)
{
  const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N_query) return;
  uint8_t pos[MERGE_MAX_LISTS];
  float headd[MERGE_MAX_LISTS];
  for (uint32_t l = 0; l < n_lists; ++l) {
    pos[l] = 0;
    headd[l] = dists[l * list_stride + n * query_stride];
  }
  for (uint32_t k = 0; k < K; ++k) {
    uint32_t bl = 0;
    float bd = G200_INF;
    bool have = false;
    for (uint32_t l = 0; l < n_lists; ++l) {
      if (pos[l] >= K_in) continue;
      if (!have || headd[l] < bd) {  // ties: lower list index first
        bd = headd[l];
        bl = l;
        have = true;
      }
    }
    const size_t src = bl * list_stride + n * query_stride + pos[bl];
    const int32_t id = ids[src];
    out_ids[static_cast<size_t>(n) * K + k] = (id >= 0) ? static_cast<int32_t>(id + id_offset_per_list * bl) : id;
    if (out_dists) out_dists[static_cast<size_t>(n) * K + k] = bd;
    pos[bl]++;
    if (pos[bl] < K_in) headd[bl] = dists[src + 1];
  }
}

}  // namespace g200

using namespace g200;

extern "C" int ggnn_b200_merge_topk(const int32_t* d_ids, const float* d_dists, uint32_t n_lists, size_t list_stride,
                                    size_t query_stride, uint32_t K_in, uint32_t N_query, uint32_t K,
                                    int64_t id_offset_per_list, int32_t* d_out_ids, float* d_out_dists,
                                    ggnn_b200_stream_t stream_)
{
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d_ids || !d_dists || !d_out_ids) return set_error(GGNN_B200_ERR_INVALID, "null device pointer");
  if (n_lists == 0 || n_lists > MERGE_MAX_LISTS) return set_error(GGNN_B200_ERR_INVALID, "n_lists must be in [1, 64]");
  if (K_in == 0 || K_in > 255 * 1u + 0u) return set_error(GGNN_B200_ERR_UNSUPPORTED, "K_in must be in [1, 255]");
  if (K == 0 || static_cast<uint64_t>(K) > static_cast<uint64_t>(K_in) * n_lists)
    return set_error(GGNN_B200_ERR_INVALID, "need 1 <= K <= K_in * n_lists");
  if (N_query == 0) return 0;
  const int block = 128;
  const int grid = (N_query + block - 1) / block;
  merge_topk_kernel<<<grid, block, 0, stream>>>(d_ids, d_dists, n_lists, list_stride, query_stride, K_in, N_query, K,
                                                static_cast<long long>(id_offset_per_list), d_out_ids, d_out_dists);
  return set_cuda_error(cudaGetLastError(), "merge_topk_kernel launch");
}
