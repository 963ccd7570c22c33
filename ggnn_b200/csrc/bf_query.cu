// bf_query.cu -- exact brute-force kNN (ground truth generation).
// Replaces src/ggnn/query/bf_query_layer.cu:39-65 + include/ggnn/cuda_utils/k_best_list.cuh:77-109
// and the launcher src/ggnn/query/query_kernels.cu:188-264.
//
// The reference streams the whole base once per query (one block per query, one vector in flight).
// Here a CTA owns 32 queries (8 warps x 4 queries, query vectors in registers), the base is
// streamed once per CTA through a double-buffered shared-memory tile filled by ONE bulk async copy
// (TMA engine) per tile, and every distance is still produced in the reference's exact fp32
// summation order, so ids AND distances are bit-identical to the reference's.
#include "traverse.cuh"
#include "host_util.h"
#include "../../include/ggnn_b200.h"

#include <algorithm>

namespace g200 {

// bf_tc.cu
bool tc_supported(uint32_t D, uint32_t K, int measure);
int tc_bf_query(const ggnn_b200_bf_query_params& p, uint32_t Nq, void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t tc_workspace_bytes(uint32_t N, uint32_t Nq, uint32_t D, uint32_t K);

constexpr int BF_WARPS = 8;
constexpr int BF_QW = 4;  // queries per warp

// sorted K-best list in registers: slot p = 32*j + lane (k_best_list.cuh:29-109)
template <int NSK>
struct WarpKBest {
  int id[NSK];
  float dist[NSK];
  __device__ __forceinline__ void init()
  {
#pragma unroll
    for (int j = 0; j < NSK; ++j) {
      id[j] = EMPTY_KEY;
      dist[j] = G200_INF;
    }
  }
  __device__ __forceinline__ float dist_at(uint32_t p) const
  {
    float v = dist[0];
#pragma unroll
    for (int j = 1; j < NSK; ++j) v = (p >> 5) == j ? dist[j] : v;
    return __shfl_sync(FULL, v, p & 31);
  }
  // k_best_list.cuh:77-109: entries with dist > d move right, the new entry goes after all
  // entries with dist <= d (ties keep the earlier-inserted first)
  __device__ __forceinline__ void add(float d, int i)
  {
    const int lane = lane_id();
    int nid[NSK];
    float nd[NSK];
#pragma unroll
    for (int j = 0; j < NSK; ++j) {
      int pi = __shfl_up_sync(FULL, id[j], 1);
      float pd = __shfl_up_sync(FULL, dist[j], 1);
      if (j > 0) {
        const int ci = __shfl_sync(FULL, id[j - 1], 31);
        const float cd = __shfl_sync(FULL, dist[j - 1], 31);
        if (lane == 0) {
          pi = ci;
          pd = cd;
        }
      }
      const bool first = (j == 0 && lane == 0);
      const bool shift_in = !first && (d < pd);            // left neighbour moves into this slot
      const bool ins = (d < dist[j]) && (first || pd <= d);  // :100
      nid[j] = ins ? i : (shift_in ? pi : id[j]);
      nd[j] = ins ? d : (shift_in ? pd : dist[j]);
    }
#pragma unroll
    for (int j = 0; j < NSK; ++j) {
      id[j] = nid[j];
      dist[j] = nd[j];
    }
  }
};

struct BfArgs {
  ggnn_b200_bf_query_params p;
  uint32_t N_query;
  uint32_t tile_rows;
  uint32_t block_dim_x;
};

// ---- fast path: D == 32*NI, VB == 32 ----
template <int NI, int NSK, int MEASURE>
__global__ void __launch_bounds__(BF_WARPS * 32) bf_kernel_fast(const BfArgs a)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int D = 32 * NI;
  const ggnn_b200_bf_query_params& p = a.p;
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const uint32_t TB = a.tile_rows;
  float* tile[2] = {reinterpret_cast<float*>(smem_raw), reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(TB) * D};
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * static_cast<size_t>(TB) * D * 4);

  const uint32_t N = static_cast<uint32_t>(p.N_base);
  const uint32_t ntiles = (N + TB - 1) / TB;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](uint32_t t) {
    const uint32_t rows = min(TB, N - t * TB);
    const uint32_t bytes = rows * D * 4u;
    mbar_expect_tx(&bars[t & 1], bytes);
    bulk_g2s(tile[t & 1], p.d_base + static_cast<size_t>(t) * TB * D, bytes, &bars[t & 1]);
  };
  if (threadIdx.x == 0) {
    issue(0);
    if (ntiles > 1) issue(1);
  }

  // this warp's queries
  const uint32_t q0 = (blockIdx.x * BF_WARPS + warp) * BF_QW;
  float q[BF_QW][NI];
  float qn[BF_QW];
  WarpKBest<NSK> best[BF_QW];
#pragma unroll
  for (int qi = 0; qi < BF_QW; ++qi) {
    const uint32_t n = min(q0 + qi, a.N_query - 1);
#pragma unroll
    for (int it = 0; it < NI; ++it) q[qi][it] = p.d_query[static_cast<size_t>(n) * D + lane + 32 * it];
    qn[qi] = 0.f;
    if (MEASURE != 0) {
      float s = 0.f;
#pragma unroll
      for (int it = 0; it < NI; ++it) s = fmaf(q[qi][it], q[qi][it], s);
      qn[qi] = warp_tree_sum(s);
    }
    best[qi].init();
  }
  const uint32_t K = p.KQuery;

  for (uint32_t t = 0; t < ntiles; ++t) {
    mbar_wait(&bars[t & 1], (t >> 1) & 1);
    const float* rows = tile[t & 1];
    const uint32_t nrows = min(TB, N - t * TB);
    for (uint32_t g = 0; g < nrows; g += 8) {
      const int nr = min(8u, nrows - g);
      const float* rg = rows + static_cast<size_t>(g) * D + lane;
      float dq[BF_QW];
      if (MEASURE == 0) {
        float v[BF_QW][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float b[NI];
#pragma unroll
          for (int it = 0; it < NI; ++it) b[it] = (i < nr) ? rg[i * D + 32 * it] : 0.f;
#pragma unroll
          for (int qi = 0; qi < BF_QW; ++qi) {
            float acc = 0.f;
#pragma unroll
            for (int it = 0; it < NI; ++it) {
              const float diff = b[it] - q[qi][it];
              acc = fmaf(diff, diff, acc);
            }
            v[qi][i] = acc;
          }
        }
#pragma unroll
        for (int qi = 0; qi < BF_QW; ++qi) dq[qi] = warp_tree_sum8(v[qi]);
      }
      else {
#pragma unroll
        for (int qi = 0; qi < BF_QW; ++qi) dq[qi] = dist8_fast<NI, 1>(rows + static_cast<size_t>(g) * D, nr, 1, q[qi], qn[qi]);
      }
      // bf_query_layer.cu:52-57: ascending base index, `if (dist < worst) add_unique`
      const unsigned rowmask = (nr >= 8 ? 0xffu : ((1u << nr) - 1u));
#pragma unroll
      for (int qi = 0; qi < BF_QW; ++qi) {
        unsigned rem = rowmask;
        while (true) {
          const float worst = best[qi].dist_at(K - 1);
          const unsigned pm = __ballot_sync(FULL, dq[qi] < worst) & rem;
          if (!pm) break;
          const int r = __ffs(pm) - 1;
          const float d = __shfl_sync(FULL, dq[qi], r);
          best[qi].add(d, static_cast<int>(t * TB + g + r));
          rem &= ~((2u << r) - 1u);
        }
      }
    }
    __syncthreads();  // everyone is done with this buffer
    if (threadIdx.x == 0 && t + 2 < ntiles) issue(t + 2);
  }

#pragma unroll
  for (int qi = 0; qi < BF_QW; ++qi) {
    const uint32_t n = q0 + qi;
    if (n >= a.N_query) continue;
#pragma unroll
    for (int j = 0; j < NSK; ++j) {
      const uint32_t k = 32u * j + lane;
      if (k < K) {
        p.d_query_results[static_cast<size_t>(n) * K + k] = best[qi].id[j];
        if (p.d_query_results_dists) p.d_query_results_dists[static_cast<size_t>(n) * K + k] = best[qi].dist[j];
      }
    }
  }
}

// ---- generic path: any D / block_dim_x, one warp per query, rows read from global memory ----
template <int NSK>
__global__ void __launch_bounds__(BF_WARPS * 32) bf_kernel_generic(const BfArgs a)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const ggnn_b200_bf_query_params& p = a.p;
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const uint32_t n = blockIdx.x * BF_WARPS + warp;
  if (n >= a.N_query) return;
  float* s_q = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(warp) * p.D;
  const DistCfg dc{p.D, a.block_dim_x, 4u, p.measure};
  QueryVec<false, 1, 1> qv;
  qv.load(dc, p.d_query + static_cast<size_t>(n) * p.D, s_q);
  WarpKBest<NSK> best;
  best.init();
  const uint32_t K = p.KQuery;
  for (int i = 0; i < p.N_base; ++i) {
    float x, y;
    dist_partials_generic(dc, p.d_base + static_cast<size_t>(i) * p.D, s_q, x, y);
    const float d = p.measure == 0 ? x : cosine_finish(x, y, qv.q_norm);
    if (d < best.dist_at(K - 1)) best.add(d, i);
  }
#pragma unroll
  for (int j = 0; j < NSK; ++j) {
    const uint32_t k = 32u * j + lane;
    if (k < K) {
      p.d_query_results[static_cast<size_t>(n) * K + k] = best.id[j];
      if (p.d_query_results_dists) p.d_query_results_dists[static_cast<size_t>(n) * K + k] = best.dist[j];
    }
  }
}

// ---- any K (129 .. 6000, src/ggnn/query/query_kernels.cu:204-218): the K-best list lives in shared memory ----
// One warp per query; the list is kept sorted by KBestList::add_unique's rule (k_best_list.cuh:77-109: entries with a
// larger distance move right, the new entry goes after all entries with dist <= d), 32 slots per step from the right.
constexpr int BF_BIGK_WARPS = 4;
__global__ void __launch_bounds__(BF_BIGK_WARPS * 32) bf_kernel_bigk(const BfArgs a)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const ggnn_b200_bf_query_params& p = a.p;
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const uint32_t n = blockIdx.x * BF_BIGK_WARPS + warp;
  if (n >= a.N_query) return;
  const uint32_t K = p.KQuery;
  float* s_q = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(warp) * p.D;
  float* l_d = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(BF_BIGK_WARPS) * p.D + static_cast<size_t>(warp) * 2 * K;
  int* l_i = reinterpret_cast<int*>(l_d + K);
  const DistCfg dc{p.D, a.block_dim_x, 4u, p.measure};
  QueryVec<false, 1, 1> qv;
  qv.load(dc, p.d_query + static_cast<size_t>(n) * p.D, s_q);
  for (uint32_t k = lane; k < K; k += 32) {
    l_d[k] = G200_INF;
    l_i[k] = EMPTY_KEY;
  }
  __syncwarp();
  for (int i = 0; i < p.N_base; ++i) {
    float x, y;
    dist_partials_generic(dc, p.d_base + static_cast<size_t>(i) * p.D, s_q, x, y);
    const float d = p.measure == 0 ? x : cosine_finish(x, y, qv.q_norm);
    if (!(d < l_d[K - 1])) continue;
    uint32_t pos = 0;
    for (int b = static_cast<int>((K - 1) / 32 * 32); b >= 0; b -= 32) {
      const uint32_t k = b + lane;
      const bool in = k < K;
      const float od = in ? l_d[k] : 0.f;
      const int oi = in ? l_i[k] : 0;
      __syncwarp();
      if (in && d < od && k + 1 < K) {
        l_d[k + 1] = od;
        l_i[k + 1] = oi;
      }
      pos += __popc(__ballot_sync(FULL, in && od <= d));
      __syncwarp();
    }
    if (lane == 0) {
      l_d[pos] = d;
      l_i[pos] = i;
    }
    __syncwarp();
  }
  for (uint32_t k = lane; k < K; k += 32) {
    p.d_query_results[static_cast<size_t>(n) * K + k] = l_i[k];
    if (p.d_query_results_dists) p.d_query_results_dists[static_cast<size_t>(n) * K + k] = l_d[k];
  }
}

template <int NI, int NSK>
static int launch_fast(const BfArgs& a, cudaStream_t stream)
{
  const size_t smem = 2 * static_cast<size_t>(a.tile_rows) * 32 * NI * 4 + 16;
  const int grid = (a.N_query + BF_WARPS * BF_QW - 1) / (BF_WARPS * BF_QW);
  auto run = [&](auto kern) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(bf_kernel_fast)");
    kern<<<grid, BF_WARPS * 32, smem, stream>>>(a);
    return set_cuda_error(cudaGetLastError(), "bf_kernel_fast launch");
  };
  return a.p.measure == 0 ? run(bf_kernel_fast<NI, NSK, 0>) : run(bf_kernel_fast<NI, NSK, 1>);
}

template <int NSK>
static int launch_generic(const BfArgs& a, cudaStream_t stream)
{
  const size_t smem = static_cast<size_t>(BF_WARPS) * a.p.D * 4;
  const int grid = (a.N_query + BF_WARPS - 1) / BF_WARPS;
  auto kern = bf_kernel_generic<NSK>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(bf_kernel_generic)");
  kern<<<grid, BF_WARPS * 32, smem, stream>>>(a);
  return set_cuda_error(cudaGetLastError(), "bf_kernel_generic launch");
}

}  // namespace g200

using namespace g200;

extern "C" size_t ggnn_b200_bf_query_workspace_bytes(uint32_t D, int32_t measure, uint32_t KQuery, uint32_t N_base,
                                                     uint32_t N_query)
{
  if (!tc_supported(D, KQuery, measure) || N_base < 128 || N_query == 0) return 0;
  return tc_workspace_bytes(N_base, N_query, D, KQuery);
}

extern "C" int ggnn_b200_bf_query(const ggnn_b200_bf_query_params* pin, uint32_t N_query, ggnn_b200_stream_t stream_)
{
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!pin) return set_error(GGNN_B200_ERR_INVALID, "null params");
  BfArgs a{};
  a.p = *pin;
  const ggnn_b200_bf_query_params& p = a.p;
  if (!p.d_base || !p.d_query || !p.d_query_results) return set_error(GGNN_B200_ERR_INVALID, "null device pointer");
  if (p.measure != GGNN_B200_EUCLIDEAN && p.measure != GGNN_B200_COSINE)
    return set_error(GGNN_B200_ERR_INVALID, "unknown distance measure");
  // query_kernels.cu:204-218
  if (p.KQuery == 0 || p.KQuery > 6000) return set_error(GGNN_B200_ERR_INVALID, "KQuery must be in [1, 6000]");
  if (p.D == 0 || p.D > 4096) return set_error(GGNN_B200_ERR_INVALID, "D must be in [1, 4096]");
  if (p.N_base <= 0) return set_error(GGNN_B200_ERR_INVALID, "N_base must be positive");
  if (N_query == 0) return 0;
  if (p.d_workspace && tc_supported(p.D, p.KQuery, p.measure) && p.N_base >= 128 && env_u32("GGNN_B200_BF_TC", 1))
    return tc_bf_query(p, N_query, p.d_workspace, p.workspace_bytes, stream);
  a.N_query = N_query;
  a.block_dim_x = std::max(32u, bit_ceil_u32((p.D + 3) / 4));
  const int NSK = (p.KQuery + 31) / 32;
  if (NSK > 4) {  // up to the reference's MAX_K_QUERY = 6000: list in shared memory
    const size_t smem = static_cast<size_t>(BF_BIGK_WARPS) * (p.D * 4 + static_cast<size_t>(p.KQuery) * 8);
    if (smem > device_info().smem_per_block_optin) return set_error(GGNN_B200_ERR_UNSUPPORTED, "bf_query: KQuery * D too large for shared memory");
    cudaError_t e = cudaFuncSetAttribute(bf_kernel_bigk, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(bf_kernel_bigk)");
    bf_kernel_bigk<<<(N_query + BF_BIGK_WARPS - 1) / BF_BIGK_WARPS, BF_BIGK_WARPS * 32, smem, stream>>>(a);
    return set_cuda_error(cudaGetLastError(), "bf_kernel_bigk launch");
  }
  const bool fast = (p.D % 32 == 0) && (p.D <= 128);
  if (fast) {
    a.tile_rows = env_u32("GGNN_B200_BF_TILE_ROWS", 64);
#define G200_BF(NI_, NSK_) return launch_fast<NI_, NSK_>(a, stream)
    switch ((p.D / 32) * 10 + NSK) {
      case 11: G200_BF(1, 1);
      case 12: G200_BF(1, 2);
      case 13: G200_BF(1, 3);
      case 14: G200_BF(1, 4);
      case 21: G200_BF(2, 1);
      case 22: G200_BF(2, 2);
      case 23: G200_BF(2, 3);
      case 24: G200_BF(2, 4);
      case 31: G200_BF(3, 1);
      case 32: G200_BF(3, 2);
      case 33: G200_BF(3, 3);
      case 34: G200_BF(3, 4);
      case 41: G200_BF(4, 1);
      case 42: G200_BF(4, 2);
      case 43: G200_BF(4, 3);
      case 44: G200_BF(4, 4);
    }
#undef G200_BF
  }
  switch (NSK) {
    case 1: return launch_generic<1>(a, stream);
    case 2: return launch_generic<2>(a, stream);
    case 3: return launch_generic<3>(a, stream);
    case 4: return launch_generic<4>(a, stream);
  }
  return set_error(GGNN_B200_ERR_UNSUPPORTED, "no kernel variant");
}
