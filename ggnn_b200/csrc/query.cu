// query.cu -- batched ANN query: best-first traversal of the layer-0 kNN graph, one warp per query.
// Replaces src/ggnn/query/query_layer.cu:39-97 + include/ggnn/cuda_utils/simple_knn_cache.cuh of the
// reference; launcher replaces src/ggnn/query/query_kernels.cu:50-186.
#include "traverse.cuh"
#include "host_util.h"
#include "../../include/ggnn_b200.h"

#include <algorithm>
#include <cstdlib>

#ifndef G200_QUERY_MB
#define G200_QUERY_MB 6  // min resident CTAs (4 warps each) per SM the register allocation must allow (80 registers)
#endif

#ifndef G200_QUERY_MB_G4
#define G200_QUERY_MB_G4 5  // the gather4 variants keep 16 stage rows per warp: shared memory allows 5 CTAs per SM
#endif

namespace g200 {

struct QueryArgs {
  ggnn_b200_query_params p;
  uint32_t N_query;
  uint32_t warps_per_cta;
  uint32_t warp_smem_bytes;
  uint32_t stage_rows;
  uint32_t stage_mode;
  uint32_t prefetch;   // 0 off; 1: L2-prefetch the adjacency rows of promising candidates; 2: speculative next-anchor load
  uint32_t hsize;      // visited hash slots (power of two)
  uint32_t ring_cap;   // 0 = ring mirror not needed
  uint32_t off_sq, off_sorted, off_hash, off_ring, off_bar;  // byte offsets in the per-warp block
  uint32_t off_lists;  // SmemLists backing store (sorted_size > 256 only)
  int32_t pad_row;     // gather4 staging: row index that pads a group of four (see traverse.cuh)
  TensorMapStorage tmap;  // gather4 staging (stage_mode 3): tensor map of the base
};

// query_layer.cu:81-90 (+ simple_knn_cache.cuh:344-352): the K best of query n into the (interleaved) result buffer and /
// or, for the fused shard-merge exchange, straight into every destination GPU's gathered buffer (peer-mapped memory: the
// stores travel over NVLink while the other warps keep traversing)
template <class LT>
__device__ __forceinline__ void write_query_results(const LT& L, const ggnn_b200_query_params& p, uint32_t n)
{
  if (p.d_query_results) {
    const size_t row = (static_cast<size_t>(n) * p.shards_per_gpu + p.on_gpu_shard_id) * p.KQuery;
    const int id_off = static_cast<int>(p.on_gpu_shard_id) * p.N_base;
    L.write_results(p.d_query_results + row, p.d_query_results_dists ? p.d_query_results_dists + row : nullptr, p.KQuery,
                    id_off);
  }
  if (p.n_scatter) {
    const size_t srow = (static_cast<size_t>(p.scatter_slot) * p.scatter_rows + n) * p.KQuery;
    for (uint32_t t = 0; t < p.n_scatter; ++t) {
      char* dst = static_cast<char*>(p.d_scatter_dst[t]);
      L.write_results(reinterpret_cast<int*>(dst) + srow, reinterpret_cast<float*>(dst + p.scatter_dists_offset) + srow,
                      p.KQuery, 0);
    }
  }
}

// exchange: the last warp of the launch to get here bumps the flag word of every destination; every warp's result
// stores are ordered before its count (system-scope fence), the count before the flags
__device__ __forceinline__ void signal_exchange(const ggnn_b200_query_params& p, uint32_t total_warps)
{
  if (p.n_scatter && p.d_scatter_done) {
    const int lane = lane_id();
    __threadfence_system();
    unsigned old = 0;
    if (lane == 0) old = atomicAdd(p.d_scatter_done, 1u);
    old = __shfl_sync(FULL, old, 0);
    if (old == total_warps - 1) {
      if (lane == 0) *p.d_scatter_done = 0;  // ready for the next launch on this stream
      __threadfence_system();
      if (p.d_scatter_flags)
        for (uint32_t t = lane; t < p.n_scatter; t += 32) atomicAdd_system(p.d_scatter_flags[t], 1u);
    }
  }
}

// LT = WarpLists<NS> (best list + prioQ in registers, sorted_size <= 256) or SmemLists (anything larger)
template <class LT, bool FAST, int NI, bool G4 = false, bool IL = false>
__global__ void __launch_bounds__(128, G4 ? G200_QUERY_MB_G4 : G200_QUERY_MB) query_kernel(const __grid_constant__ QueryArgs a)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const ggnn_b200_query_params& p = a.p;

  unsigned char* wbase = smem_raw + static_cast<size_t>(warp) * a.warp_smem_bytes;
  WarpSmem ws;
  ws.stage = reinterpret_cast<float*>(wbase);
  ws.s_q = reinterpret_cast<float*>(wbase + a.off_sq);
  ws.s_sorted = reinterpret_cast<int*>(wbase + a.off_sorted);
  ws.bar = reinterpret_cast<uint64_t*>(wbase + a.off_bar);
  ws.parity = 0;
  ws.stage_rows = a.stage_rows;
  ws.stage_mode = a.stage_mode;
  ws.tmap = G4 ? &a.tmap : nullptr;
  ws.pad_row = G4 ? a.pad_row : 0;
  if (lane < 4) mbar_init(&ws.bar[lane], 1);
  mbar_fence_init();
  __syncwarp();

  VisitedSet V;
  V.tab = reinterpret_cast<int*>(wbase + a.off_hash);
  V.hmask = a.hsize - 1;
  V.hshift = 32 - (31 - __clz(a.hsize));
  V.ring = a.ring_cap ? reinterpret_cast<int*>(wbase + a.off_ring) : nullptr;
  V.vcap = p.cache_size - p.sorted_size;
  V.vpos = 0;

  const DistCfg dc{p.D, p.block_dim_x, 4u, p.measure};
  // query_layer.cu:48-50
  const float max_nn1 = p.d_nn1_stats[1];
  const float xi = (p.measure == 0) ? __fmul_rn(__fmul_rn(__fmul_rn(max_nn1, max_nn1), p.tau_query), p.tau_query)
                                    : __fmul_rn(max_nn1, p.tau_query);

  const uint32_t total_warps = gridDim.x * a.warps_per_cta;
  uint32_t n = blockIdx.x * a.warps_per_cta + warp;
  if (p.d_work_counter) {
    if (lane == 0) n = atomicAdd(p.d_work_counter, 1u);
    n = __shfl_sync(FULL, n, 0);
  }

  while (n < a.N_query) {
    QueryVec<FAST, NI, 1> qv;
    qv.load(dc, p.d_query + static_cast<size_t>(n) * p.D, ws.s_q);
    LT L;
    L.init(p.KQuery, wbase + a.off_lists, p.sorted_size);
    V.clear();
    Stats st{0, 0};

    // query_layer.cu:55 fetch_unfiltered(d_starting_points, nullptr, S)
    for (uint32_t i = 0; i < p.num_starting_points; i += 32) {
      const int ck = (i + lane < p.num_starting_points) ? p.d_starting_points[i + lane] : EMPTY_KEY;
      fetch<LT, FAST, NI, 1, false, G4, IL>(L, V, ws, qv, p.d_base, nullptr, ck, xi, st, a.prefetch ? p.d_graph : nullptr,
                                    p.KBuild);
    }

    SpecRow spec{EMPTY_KEY, EMPTY_KEY};
    const bool use_spec = a.prefetch >= 2 && p.KBuild <= 32;
    for (uint32_t ite = 0; ite < p.max_iterations; ++ite) {
      // :58-63
      const float best0 = L.dist_at(0);
      const float r_xi = (p.measure == 0) ? fminf(xi, __fmul_rn(__fmul_rn(best0, p.tau_query), p.tau_query))
                                          : fminf(xi, __fmul_rn(best0, p.tau_query));
      const float crit = L.dist_at(L.BEST - 1) + r_xi;
      const int anchor = L.pop(crit);
      if (anchor == EMPTY_KEY) break;
      V.insert(anchor);
      st.pops++;
      // :69-76
      for (uint32_t i = 0; i < p.KBuild; i += 32) {
        int ck;
        if (use_spec && spec.key == anchor)
          ck = spec.row;  // the speculative load issued before the previous push loop was right
        else
          ck = (i + lane < p.KBuild) ? __ldg(p.d_graph + static_cast<size_t>(anchor) * p.KBuild + i + lane) : EMPTY_KEY;
        spec.key = EMPTY_KEY;
        fetch<LT, FAST, NI, 1, true, G4, IL>(L, V, ws, qv, p.d_base, nullptr, ck, r_xi, st, a.prefetch ? p.d_graph : nullptr,
                                     p.KBuild, use_spec ? &spec : nullptr);
      }
    }

    write_query_results(L, p, n);
    if (p.d_stats && lane == 0) {
      p.d_stats[2 * static_cast<size_t>(n)] = st.pops;
      p.d_stats[2 * static_cast<size_t>(n) + 1] = st.dists;
    }

    if (p.d_work_counter) {
      if (lane == 0) n = atomicAdd(p.d_work_counter, 1u);
      n = __shfl_sync(FULL, n, 0);
    }
    else {
      n += total_warps;
    }
  }

  signal_exchange(p, total_warps);
}

template <class LT, bool FAST, int NI, bool G4 = false, bool IL = false>
static int launch(const QueryArgs& a, int grid, size_t smem, cudaStream_t stream)
{
  auto kern = query_kernel<LT, FAST, NI, G4, IL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(query_kernel)");
  kern<<<grid, a.warps_per_cta * 32, smem, stream>>>(a);
  return set_cuda_error(cudaGetLastError(), "query_kernel launch");
}


// ================================================================================================
// native uint8 rows (BaseT = uint8_t, include/ggnn/base/lib.h:26-28): rows of D bytes staged by TMA gather4 (32 rows of
// 128 bytes fit where 8 fp32 rows did), distances in integer arithmetic -- per row and lane one LDS.32 (4 dims), one
// VABSDIFF4, one IDP4A, and ONE warp-wide integer REDUX instead of a shuffle tree.  Exact: every partial sum is an integer
// below 2^24 for D <= 256, so the reference's fp32 accumulation of static_cast<float>(value) terms
// (distance.cuh:104-148) yields the same numbers in any order.
// ================================================================================================
// D32 = D / 32 is a compile-time constant: the row offsets are immediates of the LDS instructions.  Per row and lane:
// LDS.32 + VABSDIFF4 + IDP4A + REDUX + (compare, select) -- the conversion to float and the cosine finish happen once per
// lane after all rows.
template <class LT, int D32, bool FILTER>
__device__ __forceinline__ void fetch_u8(LT& L, const VisitedSet& V, WarpSmem& ws, const uint32_t (&q)[(D32 + 3) / 4], float q_norm,
                                         int measure, int ck, float xi, Stats& st, const int* __restrict__ pf_graph,
                                         uint32_t pf_stride, SpecRow* spec)
{
  constexpr int WORDS = 8 * D32;        // 32-bit words per row
  constexpr int W = (D32 + 3) / 4;      // words per lane
  constexpr bool FULLW = (D32 % 4) == 0;
  const int lane = lane_id();
  bool valid = ck != EMPTY_KEY;
  if constexpr (FILTER) {
    __syncwarp();
    L.store_keys(ws.s_sorted);
    __syncwarp();
    if (valid) valid = !L.in_sorted(ws.s_sorted, ck) && !V.contains(ck);
  }
  const unsigned mask = __ballot_sync(FULL, valid);
  const int cnt = __popc(mask);
  if (cnt == 0) return;
  st.dists += cnt;
  __syncwarp();
  if (valid) ws.s_sorted[__popc(mask & ((1u << lane) - 1u))] = ck;
  __syncwarp();
  const int key_r = ws.s_sorted[lane];

  // all rows of the fetch in flight at once: lane 4j fetches candidates 4j .. 4j+3 with ONE instruction
  constexpr uint32_t ROW_BYTES = WORDS * 4u;
  const int mp = lane < cnt ? key_r : ws.pad_row;
  const int m1 = __shfl_down_sync(FULL, mp, 1);
  const int m2 = __shfl_down_sync(FULL, mp, 2);
  const int m3 = __shfl_down_sync(FULL, mp, 3);
  const uint32_t bar_s = smem_u32(ws.bar);
  if (lane == 0) mbar_expect_tx_s(bar_s, static_cast<uint32_t>((cnt + 3) & ~3) * ROW_BYTES);
  if ((lane & 3) == 0 && lane < cnt)
    tma_gather4_s(smem_u32(ws.stage) + static_cast<uint32_t>(lane) * ROW_BYTES, ws.tmap, mp, m1, m2, m3, bar_s);
  mbar_wait_s(bar_s, ws.parity & 1u);
  ws.parity ^= 1u;

  // lane t owns words t (and t + 32) of every row; lanes past the end of a short row contribute nothing
  const bool in0 = FULLW || W > 1 || lane < WORDS;
  const bool in1 = W > 1 && (FULLW || lane + 32 < WORDS);
  const uint32_t* rp = reinterpret_cast<const uint32_t*>(ws.stage) + (in0 ? lane : 0);
  const uint32_t q0 = in0 ? q[0] : 0u, q1 = (W > 1 && in1) ? q[W - 1] : 0u;
  unsigned mine_a = 0u, mine_n = 0u;  // this lane's candidate: sum (b-q)^2 | dot, and |b|^2 (cosine)
  if (measure == 0) {
    for (int r0 = 0; r0 < cnt; r0 += 8) {  // (rows past cnt hold zero fill or stale bytes: their results are never used)
      const int li = lane - r0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint32_t b0 = rp[i * WORDS];
        if (!in0) b0 = 0u;
        const uint32_t s0 = __vabsdiffu4(b0, q0);
        unsigned acc = __dp4a(s0, s0, 0u);
        if constexpr (W > 1) {
          uint32_t b1 = in1 ? rp[i * WORDS + 32] : 0u;
          const uint32_t s1 = __vabsdiffu4(b1, q1);
          acc = __dp4a(s1, s1, acc);
        }
        const unsigned tot = __reduce_add_sync(FULL, acc);
        if (li == i) mine_a = tot;
      }
      rp += 8 * WORDS;
    }
  }
  else {
    for (int r0 = 0; r0 < cnt; r0 += 8) {
      const int li = lane - r0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint32_t b0 = rp[i * WORDS];
        if (!in0) b0 = 0u;
        unsigned dot = __dp4a(b0, q0, 0u), nrm = __dp4a(b0, b0, 0u);
        if constexpr (W > 1) {
          uint32_t b1 = in1 ? rp[i * WORDS + 32] : 0u;
          dot = __dp4a(b1, q1, dot);
          nrm = __dp4a(b1, b1, nrm);
        }
        const unsigned td = __reduce_add_sync(FULL, dot), tn = __reduce_add_sync(FULL, nrm);
        if (li == i) {
          mine_a = td;
          mine_n = tn;
        }
      }
      rp += 8 * WORDS;
    }
  }
  float mine = G200_INF;
  if (lane < cnt)
    mine = measure == 0 ? static_cast<float>(mine_a) : cosine_finish(static_cast<float>(mine_a), static_cast<float>(mine_n), q_norm);
  __syncwarp();  // all reads of the stage are done before the next fetch overwrites it
  finish_fetch<LT, FILTER>(L, key_r, mine, cnt, xi, pf_graph, pf_stride, spec);
}

#ifndef G200_QUERY_MB_U8
#define G200_QUERY_MB_U8 8  // 32 warps per SM (64 registers): 0.344 vs 0.392 ms per batch with 24
#endif
template <class LT, int D32>
__global__ void __launch_bounds__(128, G200_QUERY_MB_U8) query_kernel_u8(const __grid_constant__ QueryArgs a)
{
  constexpr int W = (D32 + 3) / 4;
  constexpr bool FULLW = (D32 % 4) == 0;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const ggnn_b200_query_params& p = a.p;
  unsigned char* wbase = smem_raw + static_cast<size_t>(warp) * a.warp_smem_bytes;
  WarpSmem ws;
  ws.stage = reinterpret_cast<float*>(wbase);
  ws.s_q = nullptr;
  ws.s_sorted = reinterpret_cast<int*>(wbase + a.off_sorted);
  ws.bar = reinterpret_cast<uint64_t*>(wbase + a.off_bar);
  ws.parity = 0;
  ws.stage_rows = 32;
  ws.stage_mode = 3;
  ws.tmap = &a.tmap;
  ws.pad_row = a.pad_row;
  if (lane == 0) mbar_init(&ws.bar[0], 1);
  mbar_fence_init();
  __syncwarp();

  VisitedSet V;
  V.tab = reinterpret_cast<int*>(wbase + a.off_hash);
  V.hmask = a.hsize - 1;
  V.hshift = 32 - (31 - __clz(a.hsize));
  V.ring = a.ring_cap ? reinterpret_cast<int*>(wbase + a.off_ring) : nullptr;
  V.vcap = p.cache_size - p.sorted_size;
  V.vpos = 0;

  const uint32_t words = p.D / 4;
  const uint8_t* base_q = reinterpret_cast<const uint8_t*>(p.d_query);
  const float max_nn1 = p.d_nn1_stats[1];
  const float xi = (p.measure == 0) ? __fmul_rn(__fmul_rn(__fmul_rn(max_nn1, max_nn1), p.tau_query), p.tau_query)
                                    : __fmul_rn(max_nn1, p.tau_query);
  const uint32_t total_warps = gridDim.x * a.warps_per_cta;
  uint32_t n = blockIdx.x * a.warps_per_cta + warp;
  if (p.d_work_counter) {
    if (lane == 0) n = atomicAdd(p.d_work_counter, 1u);
    n = __shfl_sync(FULL, n, 0);
  }
  const bool use_spec = a.prefetch >= 2 && p.KBuild <= 32;

  while (n < a.N_query) {
    uint32_t q[W];
    const uint32_t* gq = reinterpret_cast<const uint32_t*>(base_q + static_cast<size_t>(n) * p.D);
    unsigned qn = 0u;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const uint32_t idx = lane + 32u * w;
      q[w] = (FULLW || idx < words) ? gq[idx] : 0u;
      qn = __dp4a(q[w], q[w], qn);
    }
    const float q_norm = static_cast<float>(__reduce_add_sync(FULL, qn));  // distance.cuh:104-117 (exact)
    LT L;
    L.init(p.KQuery, nullptr, p.sorted_size);
    V.clear();
    Stats st{0, 0};

    for (uint32_t i = 0; i < p.num_starting_points; i += 32) {  // query_layer.cu:55
      const int ck = (i + lane < p.num_starting_points) ? p.d_starting_points[i + lane] : EMPTY_KEY;
      fetch_u8<LT, D32, false>(L, V, ws, q, q_norm, p.measure, ck, xi, st, a.prefetch ? p.d_graph : nullptr, p.KBuild, nullptr);
    }
    SpecRow spec{EMPTY_KEY, EMPTY_KEY};
    for (uint32_t ite = 0; ite < p.max_iterations; ++ite) {  // :58-76
      const float best0 = L.dist_at(0);
      const float r_xi = (p.measure == 0) ? fminf(xi, __fmul_rn(__fmul_rn(best0, p.tau_query), p.tau_query))
                                          : fminf(xi, __fmul_rn(best0, p.tau_query));
      const float crit = L.dist_at(L.BEST - 1) + r_xi;
      const int anchor = L.pop(crit);
      if (anchor == EMPTY_KEY) break;
      V.insert(anchor);
      st.pops++;
      for (uint32_t i = 0; i < p.KBuild; i += 32) {
        int ck;
        if (use_spec && spec.key == anchor) ck = spec.row;
        else ck = (i + lane < p.KBuild) ? __ldg(p.d_graph + static_cast<size_t>(anchor) * p.KBuild + i + lane) : EMPTY_KEY;
        spec.key = EMPTY_KEY;
        fetch_u8<LT, D32, true>(L, V, ws, q, q_norm, p.measure, ck, r_xi, st, a.prefetch ? p.d_graph : nullptr, p.KBuild,
                                use_spec ? &spec : nullptr);
      }
    }
    write_query_results(L, p, n);
    if (p.d_stats && lane == 0) {
      p.d_stats[2 * static_cast<size_t>(n)] = st.pops;
      p.d_stats[2 * static_cast<size_t>(n) + 1] = st.dists;
    }
    if (p.d_work_counter) {
      if (lane == 0) n = atomicAdd(p.d_work_counter, 1u);
      n = __shfl_sync(FULL, n, 0);
    }
    else n += total_warps;
  }
  signal_exchange(p, total_warps);
}

template <class LT, int D32>
static int launch_u8(const QueryArgs& a, int grid, size_t smem, cudaStream_t stream)
{
  auto kern = query_kernel_u8<LT, D32>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(query_kernel_u8)");
  kern<<<grid, a.warps_per_cta * 32, smem, stream>>>(a);
  return set_cuda_error(cudaGetLastError(), "query_kernel_u8 launch");
}

}  // namespace g200

using namespace g200;

extern "C" int ggnn_b200_query_shape_init(ggnn_b200_query_shape* s, uint32_t D, uint32_t KQuery,
                                          uint32_t max_iterations)
{
  // src/ggnn/query/query_kernels.cu:63-110
  if (KQuery == 0 || KQuery > 6000) return set_error(GGNN_B200_ERR_INVALID, "KQuery must be in [1, 6000]");
  const uint32_t required_sorted = next_multiple32(KQuery + 1 + 16);
  const uint32_t cache = std::max({256u, required_sorted + 32u, bit_ceil_u32(max_iterations)});
  const uint32_t cache_block = bit_ceil_u32((cache + 15) / 16);
  const uint32_t dim_block = bit_ceil_u32((D + 3) / 4);
  const uint32_t block = std::max({32u, cache_block, dim_block});
  if (max_iterations > 8192) return set_error(GGNN_B200_ERR_INVALID, "max_iterations must be <= 8192");
  if (D == 0 || D > 4096) return set_error(GGNN_B200_ERR_INVALID, "D must be in [1, 4096]");
  if (cache > 8192 || block > 1024) return set_error(GGNN_B200_ERR_INVALID, "cache size / block size out of range");
  s->cache_size = cache;
  s->block_dim_x = block;
  s->sorted_size = std::max(cache < 512u ? 64u : 32u, required_sorted);
  return 0;
}

extern "C" int ggnn_b200_query(const ggnn_b200_query_params* pin, uint32_t N_query, ggnn_b200_stream_t stream_)
{
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!pin) return set_error(GGNN_B200_ERR_INVALID, "null params");
  QueryArgs a{};
  a.p = *pin;
  ggnn_b200_query_params& p = a.p;
  if (!p.d_base || !p.d_query || !p.d_graph || !p.d_starting_points || !p.d_nn1_stats || (!p.d_query_results && !p.n_scatter))
    return set_error(GGNN_B200_ERR_INVALID, "null device pointer");
  if (p.n_scatter && (!p.d_scatter_dst || p.scatter_rows < N_query))
    return set_error(GGNN_B200_ERR_INVALID, "scatter: need destination pointers and scatter_rows >= N_query");
  if (p.measure != GGNN_B200_EUCLIDEAN && p.measure != GGNN_B200_COSINE)
    return set_error(GGNN_B200_ERR_INVALID, "unknown distance measure");
  if (p.shards_per_gpu == 0) p.shards_per_gpu = 1;
  if (p.on_gpu_shard_id >= p.shards_per_gpu) return set_error(GGNN_B200_ERR_INVALID, "on_gpu_shard_id out of range");
  ggnn_b200_query_shape shape;
  if (int rc = ggnn_b200_query_shape_init(&shape, p.D, p.KQuery, p.max_iterations)) return rc;
  if (!p.cache_size) p.cache_size = shape.cache_size;
  if (!p.sorted_size) p.sorted_size = shape.sorted_size;
  if (!p.block_dim_x) p.block_dim_x = shape.block_dim_x;
  // query_layer.cuh:45-49
  if (!(p.KQuery < p.sorted_size && p.sorted_size < p.cache_size))
    return set_error(GGNN_B200_ERR_INVALID, "need KQuery < sorted_size < cache_size");
  if (p.D > p.block_dim_x * 4) return set_error(GGNN_B200_ERR_INVALID, "D > block_dim_x * 4");
  if (p.sorted_size % 32 || p.block_dim_x % 32) return set_error(GGNN_B200_ERR_INVALID, "sorted_size/block_dim_x must be multiples of 32");
  if (p.KBuild == 0 || p.N_base <= 0) return set_error(GGNN_B200_ERR_INVALID, "bad KBuild / N_base");
  if (N_query == 0) return 0;
  const int NS = p.sorted_size / 32;
  // register-resident lists up to 256 slots (KQuery <= 239); beyond that the lists live in shared memory
  const bool smem_lists = NS > 8;
  if (p.d_work_counter) {
    cudaError_t e = cudaMemsetAsync(p.d_work_counter, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(work counter)");
  }

  if (p.base_type == GGNN_B200_BASE_U8) {
    // native uint8 rows: register lists (sorted_size <= 64), D a multiple of 16 (TMA rows) and <= 256 (exactness bound)
    // (a gather4 lands 4 rows = 4*D bytes at a 128-byte aligned shared-memory address: D % 32 == 0)
    if (smem_lists || NS > 2 || !(p.D == 32 || p.D == 64 || p.D == 96 || p.D == 128 || p.D == 256))
      return set_error(GGNN_B200_ERR_UNSUPPORTED, "native uint8 query: needs D in {32, 64, 96, 128, 256} and KQuery <= 47 (widen to fp32 otherwise)");
    if (make_row_gather_tensor_map_u8(&a.tmap, reinterpret_cast<const uint8_t*>(p.d_base), static_cast<uint64_t>(p.N_base), p.D))
      return GGNN_B200_ERR_UNSUPPORTED;
    const DeviceInfo& dev = device_info();
    const uint32_t vcap = p.cache_size - p.sorted_size;
    a.ring_cap = (vcap < p.max_iterations) ? vcap : 0;
    a.hsize = std::max(64u, bit_ceil_u32(p.max_iterations + p.max_iterations / 4 + 1));
    a.pad_row = p.N_base;  // out of bounds: zero fill, no memory traffic
    a.stage_rows = 32;
    a.stage_mode = 3;
    a.prefetch = env_u32("GGNN_B200_QUERY_PREFETCH", 2);
    a.warps_per_cta = 4;
    uint32_t off = align_up(32u * p.D, 128);
    a.off_sq = off;
    a.off_sorted = off;
    off += p.sorted_size * 4;
    a.off_lists = off;
    a.off_hash = off;
    off += a.hsize * 4;
    a.off_ring = off;
    off += align_up(a.ring_cap * 4, 16);
    a.off_bar = off;
    off += 32;
    a.warp_smem_bytes = align_up(off, 128);
    const size_t smem = static_cast<size_t>(a.warp_smem_bytes) * a.warps_per_cta;
    if (smem > dev.smem_per_block_optin) return set_error(GGNN_B200_ERR_UNSUPPORTED, "per-CTA shared memory exceeds the device limit");
    a.N_query = N_query;
    const uint32_t ctas_needed = (N_query + a.warps_per_cta - 1) / a.warps_per_cta;
    uint32_t grid = ctas_needed;
    if (p.d_work_counter) {
      const uint32_t per_sm = std::max<uint32_t>(1, std::min<uint32_t>(G200_QUERY_MB_U8, dev.smem_per_sm / (smem + 1024)));
      grid = std::min(ctas_needed, per_sm * dev.num_sms);
    }
    switch (NS * 10 + p.D / 32) {  // the instantiated row lengths: 32, 64, 96, 128 and 256 bytes
      case 11: return launch_u8<WarpLists<1>, 1>(a, grid, smem, stream);
      case 12: return launch_u8<WarpLists<1>, 2>(a, grid, smem, stream);
      case 13: return launch_u8<WarpLists<1>, 3>(a, grid, smem, stream);
      case 14: return launch_u8<WarpLists<1>, 4>(a, grid, smem, stream);
      case 18: return launch_u8<WarpLists<1>, 8>(a, grid, smem, stream);
      case 21: return launch_u8<WarpLists<2>, 1>(a, grid, smem, stream);
      case 22: return launch_u8<WarpLists<2>, 2>(a, grid, smem, stream);
      case 23: return launch_u8<WarpLists<2>, 3>(a, grid, smem, stream);
      case 24: return launch_u8<WarpLists<2>, 4>(a, grid, smem, stream);
      case 28: return launch_u8<WarpLists<2>, 8>(a, grid, smem, stream);
    }
    return set_error(GGNN_B200_ERR_UNSUPPORTED, "no uint8 kernel variant");
  }

  // the FAST variants (query vector in registers) exist for NS <= 2; everything else takes the generic kernel
  const bool fast = (p.block_dim_x == 32) && (p.D % 32 == 0) && (p.D <= 128) && (NS <= 2);
  const int NI = p.D / 32;

  // per-warp shared memory plan
  const DeviceInfo& dev = device_info();
  const uint32_t row_bytes = p.D * 4;
  const uint32_t vcap = p.cache_size - p.sorted_size;
  a.ring_cap = (vcap < p.max_iterations) ? vcap : 0;
  // visited hash: at most max_iterations keys are ever inserted; a power of two >= 1.25x that keeps probe sequences
  // short in the worst case and the table small (more resident warps per SM: 0.70 -> 0.64 ms on the bench workload)
  a.hsize = std::max(64u, bit_ceil_u32(p.max_iterations + p.max_iterations / 4 + 1));
  {  // tuning override: any power of two that can never fill up (at most max_iterations keys are ever inserted)
    const uint32_t hs = env_u32("GGNN_B200_QUERY_HASH_SLOTS", 0);
    if (hs >= 64 && (hs & (hs - 1)) == 0 && hs > p.max_iterations) a.hsize = hs;
  }
  const uint32_t sorted_mirror = smem_lists ? 128u : p.sorted_size * 4;  // SmemLists only need 32 ints of scratch
  const uint32_t lists_bytes = smem_lists ? p.sorted_size * 8 : 0;
  const uint32_t fixed = (fast ? 0 : align_up(row_bytes, 16)) + sorted_mirror + lists_bytes + a.hsize * 4 +
                         align_up(a.ring_cap * 4, 16) + 32;
  a.warps_per_cta = std::min(4u, std::max(1u, env_u32("GGNN_B200_QUERY_WARPS", 4)));
  const uint32_t target_warps_per_sm = env_u32("GGNN_B200_QUERY_WARPS_PER_SM", 16);
  const uint32_t budget = (dev.smem_per_sm - 1024 * (target_warps_per_sm / a.warps_per_cta + 1)) / target_warps_per_sm;
  uint32_t rows = budget > fixed ? (budget - fixed) / row_bytes : 0;
  rows = std::min(32u, rows / 8 * 8);
  rows = std::max(rows, 8u);
  rows = env_u32("GGNN_B200_QUERY_STAGE_ROWS", rows);
  if (rows % 8 || rows == 0 || rows > 32) return set_error(GGNN_B200_ERR_INVALID, "stage rows must be 8, 16, 24 or 32");
  // rows must be 16-byte multiples to be staged; default: TMA gather4 (3) where a variant exists, else one bulk copy per row (0)
  bool use_il = false;
  a.stage_mode = (p.D % 4) ? 2u : env_u32("GGNN_B200_STAGE_MODE", 3);
  if (a.stage_mode == 3) {  // TMA tile::gather4 row staging: register-resident fast kernels with pipelined 8-row groups only
    if (fast && rows >= 16 && (NI == 3 || NI == 4)) {  // the instantiated gather4 variants (two 8-row buffers)
      // (a driver without cuTensorMapEncodeTiled: stay with one bulk copy per row)
      // (interleaved copy of the base, if the caller has one: same shape, rows permuted internally)
      use_il = p.d_base_interleaved && NI == 4 && p.measure == GGNN_B200_EUCLIDEAN;
      if (make_row_gather_tensor_map(&a.tmap, use_il ? p.d_base_interleaved : p.d_base, static_cast<uint64_t>(p.N_base), p.D) == 0) {
        rows = 16;
        a.pad_row = env_u32("GGNN_B200_GATHER4_PAD_VALID", 0) ? 0 : p.N_base;  // out of bounds: zero fill, no memory traffic
      }
      else a.stage_mode = 0;
    }
    else a.stage_mode = 0;
  }
  a.stage_rows = rows;
  a.prefetch = env_u32("GGNN_B200_QUERY_PREFETCH", 2);  // 0 off, 1 L2 prefetch of candidate rows, 2 speculative next-anchor row load
  uint32_t off = align_up(rows * row_bytes, 16);
  a.off_sq = off;
  off += fast ? 0 : align_up(row_bytes, 16);
  a.off_sorted = off;
  off += sorted_mirror;
  a.off_lists = off;
  off += lists_bytes;
  a.off_hash = off;
  off += a.hsize * 4;
  a.off_ring = off;
  off += align_up(a.ring_cap * 4, 16);
  a.off_bar = off;
  off += 32;
  a.warp_smem_bytes = align_up(off, 128);
  // very long rows (D up to 4096 = 16 KB): fewer warps per CTA until the CTA fits
  while (static_cast<size_t>(a.warp_smem_bytes) * a.warps_per_cta > dev.smem_per_block_optin && a.warps_per_cta > 1) a.warps_per_cta /= 2;
  const size_t smem = static_cast<size_t>(a.warp_smem_bytes) * a.warps_per_cta;
  if (smem > dev.smem_per_block_optin)
    return set_error(GGNN_B200_ERR_UNSUPPORTED, "per-CTA shared memory exceeds the device limit for this D");
  a.N_query = N_query;

  const uint32_t ctas_needed = (N_query + a.warps_per_cta - 1) / a.warps_per_cta;
  uint32_t grid = ctas_needed;
  if (p.d_work_counter) {
    const uint32_t per_sm = std::max<uint32_t>(1, std::min<uint32_t>(32, dev.smem_per_sm / (smem + 1024)));
    grid = std::min(ctas_needed, per_sm * dev.num_sms);
  }

#define G200_LAUNCH(NS_, FAST_, NI_) return launch<WarpLists<NS_>, FAST_, NI_>(a, grid, smem, stream)
  if (smem_lists) return launch<SmemLists, false, 1>(a, grid, smem, stream);
  if (a.stage_mode == 3 && use_il) {
    if (NS == 1) return launch<WarpLists<1>, true, 4, true, true>(a, grid, smem, stream);
    return launch<WarpLists<2>, true, 4, true, true>(a, grid, smem, stream);
  }
  if (a.stage_mode == 3) {
    switch (NS * 10 + NI) {
      case 13: return launch<WarpLists<1>, true, 3, true>(a, grid, smem, stream);
      case 14: return launch<WarpLists<1>, true, 4, true>(a, grid, smem, stream);
      case 23: return launch<WarpLists<2>, true, 3, true>(a, grid, smem, stream);
      case 24: return launch<WarpLists<2>, true, 4, true>(a, grid, smem, stream);
      default: return set_error(GGNN_B200_ERR_UNSUPPORTED, "no gather4 kernel variant");
    }
  }
  if (fast) {
    switch (NS * 10 + NI) {
      case 11: G200_LAUNCH(1, true, 1);
      case 12: G200_LAUNCH(1, true, 2);
      case 13: G200_LAUNCH(1, true, 3);
      case 14: G200_LAUNCH(1, true, 4);
      case 21: G200_LAUNCH(2, true, 1);
      case 22: G200_LAUNCH(2, true, 2);
      case 23: G200_LAUNCH(2, true, 3);
      case 24: G200_LAUNCH(2, true, 4);
      default: break;
    }
  }
  switch (NS) {
    case 1: G200_LAUNCH(1, false, 1);
    case 2: G200_LAUNCH(2, false, 1);
    case 3: G200_LAUNCH(3, false, 1);
    case 4: G200_LAUNCH(4, false, 1);
    case 5: G200_LAUNCH(5, false, 1);
    case 6: G200_LAUNCH(6, false, 1);
    case 7: G200_LAUNCH(7, false, 1);
    case 8: G200_LAUNCH(8, false, 1);
  }
#undef G200_LAUNCH
  return set_error(GGNN_B200_ERR_UNSUPPORTED, "no kernel variant");
}
