// merge_topk.cu -- k-way merge of per-shard / per-GPU sorted result lists on the device, one WARP per query.
// Replaces cub::DeviceSegmentedRadixSort in GPUInstance::sortQueryResults
// (src/ggnn/base/gpu_instance.cu:745-790) and the CPU heap merge ResultMerger::merge
// (src/ggnn/base/result_merger.cpp:51-149): the inputs are already sorted runs, so one pass suffices.
// Like the reference there is no limit on the list length (KQuery <= 6000, query_kernels.cu:63-69) or on the
// number of lists (shards per GPU x GPUs) beyond the shared memory of a CTA (4096 lists).
//
// Lane l owns lists l, l+32, ...; every step the warp takes the smallest head (ties: lower list index, then position)
// with one redux + one ballot; the winner's next entry was prefetched one step earlier, so the chain of dependent
// global loads is hidden.  Up to 32 lists the whole state lives in registers.
#include "common.cuh"
#include "host_util.h"
#include "../../include/ggnn_b200.h"

namespace g200 {

constexpr uint32_t MERGE_MAX_LISTS = 4096;
constexpr int MERGE_WARPS = 4;

// order-preserving float -> uint map (negative values and -0.0 included); 0xffffffff = "list exhausted"
__device__ __forceinline__ uint32_t order_key(float f)
{
  const uint32_t b = __float_as_uint(f);
  const uint32_t k = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return k == 0xffffffffu ? 0xfffffffeu : k;
}

struct MergeTopkArgs {
  const int32_t* ids;
  const float* dists;
  uint32_t n_lists;
  size_t list_stride, query_stride;
  uint32_t K_in, N_query, K;
  long long id_offset_per_list;
  int32_t* out_ids;
  float* out_dists;
};

// n_lists <= 32: one list per lane, state in registers
__global__ void __launch_bounds__(MERGE_WARPS * 32) merge_topk_warp_kernel(const MergeTopkArgs a)
{
  const int lane = lane_id();
  const uint32_t n = blockIdx.x * MERGE_WARPS + (threadIdx.x >> 5);
  if (n >= a.N_query) return;
  const bool have_list = static_cast<uint32_t>(lane) < a.n_lists;
  const size_t off = have_list ? lane * a.list_stride + n * a.query_stride : 0;
  const int32_t* li = a.ids + off;
  const float* ld = a.dists + off;
  uint32_t pos = 0;
  float cur_d = 0.f, nxt_d = 0.f;
  int32_t cur_i = EMPTY_KEY, nxt_i = EMPTY_KEY;
  if (have_list) {
    cur_d = ld[0];
    cur_i = li[0];
    if (a.K_in > 1) {
      nxt_d = ld[1];
      nxt_i = li[1];
    }
  }
  uint32_t key = have_list ? order_key(cur_d) : 0xffffffffu;
  const long long my_off = a.id_offset_per_list * lane;
  int32_t* oi = a.out_ids + static_cast<size_t>(n) * a.K;
  float* od = a.out_dists ? a.out_dists + static_cast<size_t>(n) * a.K : nullptr;
  int32_t keep_i = EMPTY_KEY;
  float keep_d = G200_INF;
  for (uint32_t k = 0; k < a.K; ++k) {
    const uint32_t m = __reduce_min_sync(FULL, key);
    const int w = __ffs(__ballot_sync(FULL, key == m)) - 1;  // lowest list index among the ties
    const int32_t wi = __shfl_sync(FULL, (cur_i >= 0) ? static_cast<int32_t>(cur_i + my_off) : cur_i, w);
    const float wd = __shfl_sync(FULL, cur_d, w);
    if (lane == static_cast<int>(k & 31u)) {
      keep_i = wi;
      keep_d = wd;
    }
    if ((k & 31u) == 31u || k + 1 == a.K) {  // coalesced flush of up to 32 results
      const uint32_t k0 = k & ~31u;
      if (k0 + lane <= k) {
        oi[k0 + lane] = keep_i;
        if (od) od[k0 + lane] = keep_d;
      }
    }
    if (lane == w) {
      ++pos;
      cur_d = nxt_d;
      cur_i = nxt_i;
      key = pos < a.K_in ? order_key(cur_d) : 0xffffffffu;
      if (pos + 1 < a.K_in) {
        nxt_d = ld[pos + 1];
        nxt_i = li[pos + 1];
      }
    }
  }
}

// any number of lists: positions and head keys of all lists in shared memory, each lane caches the best of its lists
__global__ void __launch_bounds__(MERGE_WARPS * 32) merge_topk_many_kernel(const MergeTopkArgs a)
{
  extern __shared__ uint32_t s_merge[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const uint32_t n = blockIdx.x * MERGE_WARPS + warp;
  if (n >= a.N_query) return;
  uint32_t* pos = s_merge + static_cast<size_t>(warp) * 2 * a.n_lists;
  uint32_t* hkey = pos + a.n_lists;
  const size_t qoff = n * a.query_stride;
  uint32_t best_key = 0xffffffffu, best_l = 0xffffffffu;
  for (uint32_t l = lane; l < a.n_lists; l += 32) {
    pos[l] = 0;
    const uint32_t k = order_key(a.dists[l * a.list_stride + qoff]);
    hkey[l] = k;
    if (k < best_key) {
      best_key = k;
      best_l = l;
    }
  }
  int32_t* oi = a.out_ids + static_cast<size_t>(n) * a.K;
  float* od = a.out_dists ? a.out_dists + static_cast<size_t>(n) * a.K : nullptr;
  for (uint32_t k = 0; k < a.K; ++k) {
    const uint32_t m = __reduce_min_sync(FULL, best_key);
    const uint32_t wl = __reduce_min_sync(FULL, best_key == m ? best_l : 0xffffffffu);  // lowest list index among the ties
    if (best_l == wl && best_key == m) {  // exactly one lane: the owner of list wl
      const size_t src = wl * a.list_stride + qoff + pos[wl];
      const int32_t id = a.ids[src];
      oi[k] = (id >= 0) ? static_cast<int32_t>(id + a.id_offset_per_list * wl) : id;
      if (od) od[k] = a.dists[src];
      const uint32_t p = ++pos[wl];
      hkey[wl] = p < a.K_in ? order_key(a.dists[src + 1]) : 0xffffffffu;
      best_key = 0xffffffffu;
      best_l = 0xffffffffu;
      for (uint32_t l = lane; l < a.n_lists; l += 32) {
        const uint32_t kk = hkey[l];
        if (kk < best_key) {
          best_key = kk;
          best_l = l;
        }
      }
    }
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
  if (n_lists == 0 || n_lists > MERGE_MAX_LISTS) return set_error(GGNN_B200_ERR_INVALID, "n_lists must be in [1, 4096]");
  if (K_in == 0) return set_error(GGNN_B200_ERR_INVALID, "K_in must be >= 1");
  if (K == 0 || static_cast<uint64_t>(K) > static_cast<uint64_t>(K_in) * n_lists)
    return set_error(GGNN_B200_ERR_INVALID, "need 1 <= K <= K_in * n_lists");
  if (N_query == 0) return 0;
  MergeTopkArgs a{d_ids, d_dists, n_lists, list_stride, query_stride, K_in, N_query, K,
                  static_cast<long long>(id_offset_per_list), d_out_ids, d_out_dists};
  const int grid = (N_query + MERGE_WARPS - 1) / MERGE_WARPS;
  if (n_lists <= 32) {
    merge_topk_warp_kernel<<<grid, MERGE_WARPS * 32, 0, stream>>>(a);
  }
  else {
    const size_t smem = static_cast<size_t>(MERGE_WARPS) * 2 * n_lists * sizeof(uint32_t);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(merge_topk_many_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(merge_topk_many_kernel)");
    }
    merge_topk_many_kernel<<<grid, MERGE_WARPS * 32, smem, stream>>>(a);
  }
  return set_cuda_error(cudaGetLastError(), "merge_topk kernel launch");
}
