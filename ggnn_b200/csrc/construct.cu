// construct.cu -- graph construction kernels (one warp per graph point) and the build schedule.
// Replaces, kernel by kernel:
//   top              src/ggnn/construction/top_merge_layer.cu:40-82
//   nn1 stats        src/ggnn/construction/graph_construction.cu:381-393 (+ divide :79-83)
//   select           src/ggnn/construction/wrs_select_layer.cu:41-102
//   merge            src/ggnn/construction/merge_layer.cu:39-158
//   sym              src/ggnn/construction/sym_query_layer.cu:39-145 + include/ggnn/cuda_utils/simple_knn_sym_cache.cuh
//   sym_buffer_merge src/ggnn/construction/sym_buffer_merge_layer.cu:36-99
//   schedule         src/ggnn/construction/graph_construction.cu:128-147
#include "traverse.cuh"
#include "host_util.h"
#include "../../include/ggnn_b200.h"

#include <curand.h>

#include <algorithm>

namespace g200 {

constexpr int CW = 4;  // warps per CTA in the construction kernels

struct GraphPtrs {
  int32_t* graph;        // [N_all, K]
  int32_t* translation;  // [ST_all]
  int32_t* selection;    // [ST_all]
  float* nn1_stats;      // [2]
};
static GraphPtrs graph_ptrs(const ggnn_b200_graph_config& c, void* blob)
{
  ggnn_b200_graph_offsets o;
  ggnn_b200_graph_blob_offsets(&c, &o);
  char* b = static_cast<char*>(blob);
  return {reinterpret_cast<int32_t*>(b + o.graph), reinterpret_cast<int32_t*>(b + o.translation),
          reinterpret_cast<int32_t*>(b + o.selection), reinterpret_cast<float*>(b + o.nn1_stats)};
}

// per-warp shared memory plan shared by top / merge / sym
struct WarpPlan {
  uint32_t warp_smem_bytes, stage_rows, hsize, ring_cap, stage_mode;
  uint32_t off_sq, off_half, off_sorted, off_hash, off_ring, off_bar;
};
// overlay_sorted: the sorted-key mirror / compaction scratch of `fetch` lives in the first bytes of the stage buffer (it is
// only used between the end of one fetch's distance phase and the next fetch's row copies) -- 256 bytes per warp that
// decide between 5 and 6 resident CTAs per SM for the merge kernel
static int make_plan(WarpPlan& pl, uint32_t D, bool need_sq, bool need_half, uint32_t sorted, uint32_t cache,
                     uint32_t max_pops, uint32_t target_warps_per_sm, uint32_t force_rows = 0, bool overlay_sorted = false)
{
  const DeviceInfo& dev = device_info();
  const uint32_t row_bytes = D * 4;
  const uint32_t vcap = cache > sorted ? cache - sorted : 0;
  pl.ring_cap = (vcap && vcap < max_pops) ? vcap : 0;
  pl.hsize = cache ? std::max(64u, bit_ceil_u32(max_pops + max_pops / 4 + 1)) : 0;  // see query.cu
  const uint32_t fixed = (need_sq ? align_up(row_bytes, 16) : 0) + (need_half ? align_up(row_bytes, 16) : 0) +
                         (overlay_sorted ? 0 : align_up(sorted * 4, 16)) + pl.hsize * 4 + align_up(pl.ring_cap * 4, 16) + 32;
  const uint32_t budget = (dev.smem_per_sm - 1024 * (target_warps_per_sm / CW + 1)) / target_warps_per_sm;
  uint32_t rows = budget > fixed ? (budget - fixed) / row_bytes : 0;
  rows = std::max(8u, std::min(32u, rows / 8 * 8));
  pl.stage_rows = env_u32("GGNN_B200_BUILD_STAGE_ROWS", rows);
  if (pl.stage_rows % 8 || pl.stage_rows == 0 || pl.stage_rows > 32) pl.stage_rows = rows;
  if (force_rows && pl.stage_rows >= force_rows) pl.stage_rows = force_rows;
  rows = pl.stage_rows;
  pl.stage_mode = (D % 4) ? 2u : env_u32("GGNN_B200_STAGE_MODE", 0);  // rows must be 16-byte multiples to be staged
  if (pl.stage_mode == 3) pl.stage_mode = 0;  // gather4 staging is wired up for the query kernel only
  uint32_t off = align_up(rows * row_bytes, 16);
  pl.off_sq = off;
  off += need_sq ? align_up(row_bytes, 16) : 0;
  pl.off_half = off;
  off += need_half ? align_up(row_bytes, 16) : 0;
  pl.off_sorted = overlay_sorted ? 0 : off;
  off += overlay_sorted ? 0 : align_up(sorted * 4, 16);
  pl.off_hash = off;
  off += pl.hsize * 4;
  pl.off_ring = off;
  off += align_up(pl.ring_cap * 4, 16);
  pl.off_bar = off;
  off += 32;
  pl.warp_smem_bytes = align_up(off, 128);
  if (static_cast<size_t>(pl.warp_smem_bytes) * CW > dev.smem_per_block_optin)
    return set_error(GGNN_B200_ERR_UNSUPPORTED, "per-CTA shared memory exceeds the device limit for this D");
  return 0;
}

__device__ __forceinline__ void init_warp_smem(WarpSmem& ws, VisitedSet& V, unsigned char* wbase, const WarpPlan& pl,
                                               uint32_t vcap)
{
  ws.stage = reinterpret_cast<float*>(wbase);
  ws.s_q = reinterpret_cast<float*>(wbase + pl.off_sq);
  ws.s_sorted = reinterpret_cast<int*>(wbase + pl.off_sorted);
  ws.bar = reinterpret_cast<uint64_t*>(wbase + pl.off_bar);
  ws.parity = 0;
  ws.stage_rows = pl.stage_rows;
  ws.stage_mode = pl.stage_mode;
  ws.tmap = nullptr;
  ws.pad_row = 0;
  if (lane_id() < 4) mbar_init(&ws.bar[lane_id()], 1);
  mbar_fence_init();
  __syncwarp();
  V.tab = reinterpret_cast<int*>(wbase + pl.off_hash);
  V.hmask = pl.hsize ? pl.hsize - 1 : 0;
  V.hshift = pl.hsize ? 32 - (31 - __clz(pl.hsize)) : 0;
  V.ring = pl.ring_cap ? reinterpret_cast<int*>(wbase + pl.off_ring) : nullptr;
  V.vcap = vcap;
  V.vpos = 0;
}

// ================================================================================================
// top
// ================================================================================================
struct TopArgs {
  uint32_t D, KBuild, N_layer, layer, S, S_offset, VB, items;
  int32_t measure;
  const float* base;
  const int32_t* translation;  // layer translation or nullptr
  int32_t* graph;              // graph[layer]
  float* nn1;
  WarpPlan pl;
};

// sorted K-best list in registers (k_best_list.cuh:29-109), same as bf_query's
template <int NSK>
struct KBestRegs {
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
      const bool shift_in = !first && (d < pd);
      const bool ins = (d < dist[j]) && (first || pd <= d);
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

template <int NSK, bool FAST, int D32, int NW>
__global__ void __launch_bounds__(CW * 32) top_kernel(const TopArgs a)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const uint32_t n = blockIdx.x * CW + warp;
  if (n >= a.N_layer) return;
  WarpSmem ws;
  VisitedSet V;
  init_warp_smem(ws, V, smem_raw + static_cast<size_t>(warp) * a.pl.warp_smem_bytes, a.pl, 0);

  const int m = a.layer ? a.translation[n] : static_cast<int>(n);
  const DistCfg dc{a.D, a.VB, a.items, a.measure};
  QueryVec<FAST, D32, NW> qv;
  qv.load(dc, a.base + static_cast<size_t>(m) * a.D, ws.s_q);
  KBestRegs<NSK> best;
  best.init();

  // top_merge_layer.cu:54-60 segment bounds
  const uint32_t S_plus_offset = a.S_offset * (a.S + 1);
  const uint32_t S_actual = (!a.layer && n < S_plus_offset) ? a.S + 1 : a.S;
  const uint32_t start = (a.layer || n < S_plus_offset) ? (n / S_actual) * S_actual
                                                         : S_plus_offset + ((n - S_plus_offset) / S_actual) * S_actual;
  const uint32_t end = start + S_actual;

  for (uint32_t c0 = start; c0 < end; c0 += 32) {
    const uint32_t other_n = c0 + lane;
    int other_m = EMPTY_KEY;
    if (other_n < end) {
      other_m = a.layer ? a.translation[other_n] : static_cast<int>(other_n);
      if (other_m == m) other_m = EMPTY_KEY;  // :65-66
    }
    const unsigned mask = __ballot_sync(FULL, other_m != EMPTY_KEY);
    const int cnt = __popc(mask);
    if (!cnt) continue;
    const unsigned src = __fns(mask, 0, lane + 1);
    const int mm = __shfl_sync(FULL, other_m, src & 31);
    const int nn = __shfl_sync(FULL, static_cast<int>(other_n), src & 31);
    float mine = G200_INF;
    for (int b0 = 0; b0 < cnt; b0 += ws.stage_rows) {
      const int nb = min(static_cast<int>(ws.stage_rows), cnt - b0);
      const float d = stage_and_dist<FAST, D32, NW>(ws, qv, a.base, mm, b0, nb);
      if (lane >= b0 && lane < b0 + nb) mine = d;
    }
    // add_unique for every candidate in order; only those with dist < some entry change the list
    unsigned rem = cnt >= 32 ? FULL : ((1u << cnt) - 1u);
    while (true) {
      const float worst = best.dist_at(a.KBuild - 1);
      const unsigned pm = __ballot_sync(FULL, mine < worst) & rem;
      if (!pm) break;
      const int r = __ffs(pm) - 1;
      best.add(__shfl_sync(FULL, mine, r), __shfl_sync(FULL, nn, r));
      rem &= ~((2u << r) - 1u);
    }
  }
#pragma unroll
  for (int j = 0; j < NSK; ++j) {
    const uint32_t k = 32u * j + lane;
    if (k < a.KBuild) a.graph[static_cast<size_t>(n) * a.KBuild + k] = best.id[j];
  }
  float nn1 = best.dist_at(1);  // :76-81 second nearest (first is not self here: self is skipped)
  if (a.measure == 0) nn1 = __fsqrt_rn(nn1);
  if (lane == 0) a.nn1[n] = nn1;
}

// ================================================================================================
// nn1 statistics: {mean, max}
// ================================================================================================
constexpr int STATS_BLOCKS = 512;
__global__ void __launch_bounds__(256) stats_partial_kernel(const float* __restrict__ v, uint32_t N, float* partial)
{
  __shared__ float s_sum[8], s_max[8];
  float sum = 0.f, mx = -G200_INF;
  for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < N; i += static_cast<size_t>(STATS_BLOCKS) * 256) {
    const float x = v[i];
    sum += x;
    mx = fmaxf(mx, x);
  }
  for (int o = 16; o; o >>= 1) {
    sum += __shfl_xor_sync(FULL, sum, o);
    mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    s_sum[threadIdx.x >> 5] = sum;
    s_max[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tm = -G200_INF;
    for (int w = 0; w < 8; ++w) {
      ts += s_sum[w];
      tm = fmaxf(tm, s_max[w]);
    }
    partial[blockIdx.x] = ts;
    partial[STATS_BLOCKS + blockIdx.x] = tm;
  }
}
__global__ void __launch_bounds__(32) stats_final_kernel(const float* __restrict__ partial, uint32_t N, float* nn1_stats)
{
  // double accumulation of the 512 partials, fixed order -> deterministic
  double s = 0.0;
  float mx = -G200_INF;
  for (int i = threadIdx.x; i < STATS_BLOCKS; i += 32) {
    s += static_cast<double>(partial[i]);
    mx = fmaxf(mx, partial[STATS_BLOCKS + i]);
  }
  for (int o = 16; o; o >>= 1) {
    s += __shfl_xor_sync(FULL, s, o);
    mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
  }
  if (threadIdx.x == 0) {
    nn1_stats[0] = __fdiv_rn(static_cast<float>(s), static_cast<float>(N));  // graph_construction.cu:79-83
    nn1_stats[1] = mx;
  }
}

// ================================================================================================
// select (weighted reservoir sampling)
// ================================================================================================
struct SelectArgs {
  uint32_t layer, n_blocks, Sglob, S, S_offset, G, SG, SG_offset;
  const float* nn1;
  const float* rng;
  int32_t* selection;                // selection[layer+1]
  int32_t* translation;              // translation[layer+1]
  const int32_t* translation_layer;  // translation[layer] or nullptr
};

__device__ __forceinline__ uint32_t radix_key(float f)
{
  // order-preserving map used by radix sorts; -0.0 is canonicalised to +0.0 (CUB does the same)
  uint32_t b = __float_as_uint(f);
  if (b == 0x80000000u) b = 0;
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(CW * 32) select_kernel(const SelectArgs a)
{
  __shared__ uint32_t s_keys[CW][256];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const uint32_t b = blockIdx.x * CW + warp;
  if (b >= a.n_blocks) return;
  // wrs_select_layer.cu:49-50
  const uint32_t S_current = a.S + (b < a.S_offset);
  const uint32_t start = b * a.S + min(b, a.S_offset);
  uint32_t* keys = s_keys[warp];
  for (uint32_t i = lane; i < 256; i += 32) {
    uint32_t k = radix_key(-1.f);
    if (i < S_current) {
      const uint32_t n = start + i;
      // :59-62
      const float e = __fdiv_rn(-1 * logf(a.rng[n]), a.nn1[n] + 1.1920928955078125e-07f);
      k = radix_key(e);
    }
    keys[i] = k;
  }
  __syncwarp();
  const uint32_t upper_segment = b / a.G;
  const uint32_t nth = b - upper_segment * a.G;
  const uint32_t num_selected = a.SG + (nth < a.SG_offset);
  const uint32_t dest = upper_segment * a.Sglob + nth * a.SG + min(nth, a.SG_offset);
  // rank of item i in a stable descending sort = #{j: key_j > key_i} + #{j < i: key_j == key_i}
  for (uint32_t i = lane; i < S_current; i += 32) {
    const uint32_t ki = keys[i];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < S_current; ++j) {
      const uint32_t kj = keys[j];
      rank += (kj > ki) || (kj == ki && j < i);
    }
    if (rank < num_selected) {
      const int32_t n = static_cast<int32_t>(start + i);
      a.selection[dest + rank] = n;
      a.translation[dest + rank] = a.layer ? a.translation_layer[n] : n;
    }
  }
}

// ================================================================================================
// merge
// ================================================================================================
struct MergeArgs {
  uint32_t D, KBuild, S, layer_top, layer_btm, G, S0, S0_offset, N_btm, VB, items;
  uint32_t Ns_offsets[GGNN_B200_L], STs_offsets[GGNN_B200_L];
  int32_t measure;
  float tau_build;
  const float* base;
  const int32_t* selection;    // whole selection array (layer l at STs_offsets[l])
  const int32_t* translation;  // whole translation array
  const int32_t* graph;        // whole neighbourhood array
  int32_t* graph_buffer;       // [N_btm, K]
  const float* nn1_stats;
  float* nn1;
  uint32_t sorted, cache, max_iterations;
  WarpPlan pl;
  int32_t pad_row;        // gather4 staging (see traverse.cuh)
  unsigned long long* counters;  // optional {pops, distance evaluations, points} of this launch (diagnostics)
  TensorMapStorage tmap;  // tensor map of the base for the gather4 variants
};

// simple_knn_cache.cuh:297-333
template <int NS>
__device__ __forceinline__ void lists_transform(WarpLists<NS>& L, const int32_t* __restrict__ transform)
{
  const int lane = lane_id();
  const uint32_t BEST = L.BEST;
  int tk[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const uint32_t p = 32u * j + lane;
    tk[j] = (p < BEST && L.key[j] != EMPTY_KEY) ? transform[L.key[j]] : EMPTY_KEY;
  }
  int nk[NS];
  float nd[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const uint32_t p = 32u * j + lane;
    nk[j] = EMPTY_KEY;
    nd[j] = G200_INF;
    // slot p in [BEST, 2*BEST) takes the transformed best entry p-BEST
    const uint32_t sp = p - BEST;  // wraps for p < BEST (unused then)
#pragma unroll
    for (int js = 0; js < NS; ++js) {
      const int sk = __shfl_sync(FULL, tk[js], sp & 31);
      const float sd = __shfl_sync(FULL, L.dist[js], sp & 31);
      if (p >= BEST && p < 2 * BEST && p < L.SORTED && (sp >> 5) == static_cast<uint32_t>(js)) {
        nk[j] = sk;
        nd[j] = sd;
      }
    }
    if (p < BEST) {
      nk[j] = tk[j];
      nd[j] = L.dist[j];
    }
  }
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    L.key[j] = nk[j];
    L.dist[j] = nd[j];
  }
  L.head = BEST;
  L.update_flags();
}

#ifndef G200_MERGE_MB
#define G200_MERGE_MB 5  // resident CTAs per SM the gather4 variants are compiled for (register cap)
#endif
#ifndef G200_SYM_MB
#define G200_SYM_MB 6  // sym: 24 warps per SM (80 registers, 16 stage rows) instead of 16: 1M build 0.863 -> 0.800 s
#endif
template <int NS, bool FAST, int D32, int NW, bool G4 = false>
__global__ void __launch_bounds__(CW * 32, G4 ? G200_MERGE_MB : 1) merge_kernel(const __grid_constant__ MergeArgs a)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  WarpSmem ws;
  VisitedSet V;
  init_warp_smem(ws, V, smem_raw + static_cast<size_t>(warp) * a.pl.warp_smem_bytes, a.pl, a.cache - a.sorted);
  if constexpr (G4) {
    ws.tmap = &a.tmap;
    ws.pad_row = a.pad_row;
  }

  const uint32_t K = a.KBuild;
  const float mean_nn1 = a.nn1_stats[0];
  const float xi = (a.measure == 0) ? __fmul_rn(__fmul_rn(__fmul_rn(mean_nn1, mean_nn1), a.tau_build), a.tau_build)
                                    : __fmul_rn(mean_nn1, a.tau_build);
  unsigned long long tot_pops = 0, tot_dists = 0, tot_points = 0;
  // persistent warps: warp w of the grid handles points w, w + W, w + 2W, ... -- a warp slot is never held idle by the
  // slower warps of its CTA, and at any time the resident warps work on one contiguous block of points (L2 locality)
  for (uint32_t n = blockIdx.x * CW + warp; n < a.N_btm; n += gridDim.x * CW) {
  const int m = a.layer_btm ? a.translation[a.STs_offsets[a.layer_btm] + n] : static_cast<int>(n);
  const DistCfg dc{a.D, a.VB, a.items, a.measure};
  QueryVec<FAST, D32, NW> qv;
  qv.load(dc, a.base + static_cast<size_t>(m) * a.D, ws.s_q);
  WarpLists<NS> L;
  L.init(K + 1);
  V.clear();
  Stats st{0, 0};

  {  // merge_layer.cu:41-62, 86-97
    uint32_t seg_btm = n / a.S;
    if (!a.layer_btm) {
      const uint32_t offset_points = a.S0_offset * (a.S0 + 1);
      seg_btm = (n < offset_points) ? n / (a.S0 + 1) : a.S0_offset + (n - offset_points) / a.S0;
    }
    uint32_t powG = a.G;
    for (uint32_t i = 1; i < a.layer_top - a.layer_btm; ++i) powG *= a.G;
    const uint32_t s_offset = (seg_btm / powG) * a.S;
    for (uint32_t i = 0; i < a.S; i += 32) {
      const int ck = (i + lane < a.S) ? static_cast<int>(s_offset + i + lane) : EMPTY_KEY;
      fetch<WarpLists<NS>, FAST, D32, NW, false, G4>(L, V, ws, qv, a.base, a.translation + a.STs_offsets[a.layer_top], ck, xi, st);
    }
  }

  for (uint32_t layer = a.layer_top - 1; layer >= a.layer_btm && layer != 0xffffffffu; layer--) {
    lists_transform<NS>(L, a.selection + a.STs_offsets[layer + 1]);  // :101
    V.clear();
    const int32_t* tr = layer ? a.translation + a.STs_offsets[layer] : nullptr;
    if (layer == a.layer_btm) {  // :103-104
      const int ck = lane == 0 ? static_cast<int>(n) : EMPTY_KEY;
      fetch<WarpLists<NS>, FAST, D32, NW, false, G4>(L, V, ws, qv, a.base, tr, ck, xi, st);
    }
    const int32_t* layer_graph = a.graph + static_cast<size_t>(a.Ns_offsets[layer]) * K;
    SpecRow spec{EMPTY_KEY, EMPTY_KEY};
    const bool use_spec = K <= 32;
    for (uint32_t ite = 0; ite < a.max_iterations; ++ite) {
      const float crit = L.dist_at(L.BEST - 1) + xi;
      const int anchor = L.pop(crit);
      if (anchor == EMPTY_KEY) break;
      V.insert(anchor);
      st.pops++;
      for (uint32_t j = 0; j < K; j += 32) {
        int ck;
        if (use_spec && spec.key == anchor) ck = spec.row;  // speculative load issued before the previous push loop
        else ck = (j + lane < K) ? __ldg(layer_graph + static_cast<size_t>(anchor) * K + j + lane) : EMPTY_KEY;
        spec.key = EMPTY_KEY;
        fetch<WarpLists<NS>, FAST, D32, NW, true, G4>(L, V, ws, qv, a.base, tr, ck, xi, st, use_spec ? layer_graph : nullptr, K,
                                       use_spec ? &spec : nullptr);
      }
    }
  }

  // :122-145 strip the self link
  bool is_self = false;
  int self_slot = 0;
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const uint32_t p = 32u * j + lane;
    if (p < K && L.key[j] == static_cast<int>(n)) {
      is_self = true;
      self_slot = p;
    }
  }
  const unsigned selfmask = __ballot_sync(FULL, is_self);
  int own = -1;
  if (selfmask) own = __shfl_sync(FULL, self_slot, __ffs(selfmask) - 1);
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const uint32_t p = 32u * j + lane;
    // value of slot p+1
    int nxt = __shfl_down_sync(FULL, L.key[j], 1);
    if (j + 1 < NS) {
      const int c = __shfl_sync(FULL, L.key[j + 1 < NS ? j + 1 : j], 0);
      if (lane == 31) nxt = c;
    }
    if (p < K) {
      const int idx = (static_cast<int>(p) >= own) ? nxt : L.key[j];
      a.graph_buffer[static_cast<size_t>(n) * K + p] = (idx != EMPTY_KEY) ? idx : static_cast<int>(n);
    }
  }
  if (!a.layer_btm) {  // :147-157
    uint32_t i = static_cast<uint32_t>(own + 1);
    float dist;
    do {
      dist = L.dist_at(i);
      ++i;
    } while (dist == 0.0f && i < L.BEST);
    if (a.measure == 0) dist = __fsqrt_rn(dist);
    if (lane == 0) a.nn1[n] = dist;
  }
  __syncwarp();
  tot_pops += st.pops;
  tot_dists += st.dists;
  ++tot_points;
  }  // points of this warp
  if (a.counters && lane == 0) {
    atomicAdd(&a.counters[0], tot_pops);
    atomicAdd(&a.counters[1], tot_dists);
    atomicAdd(&a.counters[2], tot_points);
  }
}

// ================================================================================================
// sym
// ================================================================================================
struct SymArgs {
  uint32_t D, KBuild, N_layer, VB, items;
  int32_t measure;
  float tau_build;
  const float* base;
  const int32_t* graph;        // graph[layer]
  const int32_t* translation;  // translation[layer] or nullptr
  const float* nn1_stats;
  int32_t* sym_buffer;
  uint32_t* sym_atomic;
  uint32_t sorted, cache, max_iterations;
  WarpPlan pl;
  int32_t pad_row;        // gather4 staging (see traverse.cuh); stage_mode 3 in `pl` switches it on
  uint32_t serial;        // 1: one warp, points in order (deterministic; testing)
  unsigned long long* counters;  // optional {pops, distance evaluations (two distances each), points} (diagnostics)
  TensorMapStorage tmap;
};

// two distances (to the point itself and to the half-way point) of up to 8 staged rows, in the
// arithmetic order of simple_knn_sym_cache.cuh:214-283 (FFMA chains, also for cosine)
template <int D32, int NW>
__device__ __forceinline__ void sym_dist8(const float* __restrict__ rows, int nrows, int measure,
                                          const float (&q)[D32], const float (&h)[D32], float q_norm, float h_norm,
                                          float& out_q, float& out_h)
{
  const int lane = lane_id();
  constexpr int D = 32 * D32;
  float tq = 0.f, th = 0.f, tn = 0.f;
  if (measure == 0 && G200_PACKED_DIST) {
    // Euclidean: rows (2p, 2p+1) advance together in packed fp32 (sub.rn + fma.rn per element, like the scalar chain);
    // always the whole group -- rows >= nrows hold stale shared memory, their results are never used
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      uint64_t aq[4] = {0ull, 0ull, 0ull, 0ull}, ah[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int c = it * NW + w;
        if (c < D32) {
          const uint64_t q2 = pack2(q[c], q[c]), h2 = pack2(h[c], h[c]);
#pragma unroll
          for (int pr = 0; pr < 4; ++pr) {
            const float* rp = rows + (2 * pr) * D + 32 * c + lane;
            const uint64_t o2 = pack2(rp[0], rp[D]);
            const uint64_t dq = sub2(q2, o2), dh = sub2(h2, o2);
            aq[pr] = fma2(dq, dq, aq[pr]);
            ah[pr] = fma2(dh, dh, ah[pr]);
          }
        }
      }
      float vq[8], vh[8];
#pragma unroll
      for (int pr = 0; pr < 4; ++pr) {
        unpack2(aq[pr], vq[2 * pr], vq[2 * pr + 1]);
        unpack2(ah[pr], vh[2 * pr], vh[2 * pr + 1]);
      }
      const float sq = warp_tree_sum8(vq), sh = warp_tree_sum8(vh);
      tq = (w == 0) ? sq : tq + sq;
      th = (w == 0) ? sh : th + sh;
    }
    out_q = tq;
    out_h = th;
    return;
  }
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    float vq[8], vh[8], vn[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float aq = 0.f, ah = 0.f, an = 0.f;
      if (i < nrows) {
        const float* rp = rows + i * D + lane;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int c = it * NW + w;
          if (c < D32) {
            const float o = rp[32 * c];
            if (measure == 0) {
              const float dq = q[c] - o;
              aq = fmaf(dq, dq, aq);
              const float dh = h[c] - o;
              ah = fmaf(dh, dh, ah);
            }
            else {
              aq = fmaf(q[c], o, aq);
              ah = fmaf(h[c], o, ah);
              an = fmaf(o, o, an);
            }
          }
        }
      }
      vq[i] = aq;
      vh[i] = ah;
      vn[i] = an;
    }
    const float sq = warp_tree_sum8(vq), sh = warp_tree_sum8(vh);
    tq = (w == 0) ? sq : tq + sq;
    th = (w == 0) ? sh : th + sh;
    if (measure != 0) {
      const float sn = warp_tree_sum8(vn);
      tn = (w == 0) ? sn : tn + sn;
    }
  }
  if (measure != 0) {  // :255-272
    const float qn = __fmul_rn(tn, q_norm), hn = __fmul_rn(tn, h_norm);
    tq = (qn > 0.0f) ? fabsf(1.0f - __fdiv_rn(tq, __fsqrt_rn(qn))) : 1.0f;
    th = (hn > 0.0f) ? fabsf(1.0f - __fdiv_rn(th, __fsqrt_rn(hn))) : 1.0f;
  }
  out_q = tq;
  out_h = th;
}

// generic (any D / VB / items): one row, query + half vectors in shared memory
__device__ __forceinline__ void sym_dist_generic(const DistCfg& c, const float* __restrict__ row,
                                                 const float* __restrict__ s_q, const float* __restrict__ s_h,
                                                 float q_norm, float h_norm, float& out_q, float& out_h)
{
  const int lane = lane_id();
  float tq = 0.f, th = 0.f, tn = 0.f;
  for (uint32_t w = 0; w < c.VB / 32; ++w) {
    float aq = 0.f, ah = 0.f, an = 0.f;
    for (uint32_t it = 0; it < c.items; ++it) {
      const uint32_t d = it * c.VB + 32 * w + lane;
      if (d < c.D) {
        const float o = row[d];
        if (c.measure == 0) {
          const float dq = s_q[d] - o;
          aq = fmaf(dq, dq, aq);
          const float dh = s_h[d] - o;
          ah = fmaf(dh, dh, ah);
        }
        else {
          aq = fmaf(s_q[d], o, aq);
          ah = fmaf(s_h[d], o, ah);
          an = fmaf(o, o, an);
        }
      }
    }
    aq = warp_tree_sum(aq);
    ah = warp_tree_sum(ah);
    tq = (w == 0) ? aq : tq + aq;
    th = (w == 0) ? ah : th + ah;
    if (c.measure != 0) {
      an = warp_tree_sum(an);
      tn = (w == 0) ? an : tn + an;
    }
  }
  if (c.measure != 0) {
    const float qn = __fmul_rn(tn, q_norm), hn = __fmul_rn(tn, h_norm);
    tq = (qn > 0.0f) ? fabsf(1.0f - __fdiv_rn(tq, __fsqrt_rn(qn))) : 1.0f;
    th = (hn > 0.0f) ? fabsf(1.0f - __fdiv_rn(th, __fsqrt_rn(hn))) : 1.0f;
  }
  out_q = tq;
  out_h = th;
}

template <bool FAST, int D32, int NW>
struct SymVec {
  float q[FAST ? D32 : 1], h[FAST ? D32 : 1];
  float q_norm, h_norm;
  DistCfg cfg;
  float *s_q, *s_h;
};

// stage rows of candidates [b0,b0+nb) and compute both distances
template <bool FAST, int D32, int NW>
__device__ __forceinline__ void sym_stage_and_dist(WarpSmem& ws, const SymVec<FAST, D32, NW>& sv,
                                                   const float* __restrict__ base, int m, int b0, int nb,
                                                   float& mine_q, float& mine_h)
{
  const int lane = lane_id();
  const uint32_t D = sv.cfg.D;
  const int r = lane - b0;
  if constexpr (!FAST) {
    if (ws.stage_mode == 2) {  // unaligned rows: straight from global memory
      for (int i = 0; i < nb; ++i) {
        const int mi = __shfl_sync(FULL, m, b0 + i);
        float dq, dh;
        sym_dist_generic(sv.cfg, base + static_cast<size_t>(mi) * D, sv.s_q, sv.s_h, sv.q_norm, sv.h_norm, dq, dh);
        if (r == i) {
          mine_q = dq;
          mine_h = dh;
        }
      }
      return;
    }
  }
  stage_rows_g2s(ws, base, D, m, b0, nb);
  if constexpr (FAST) {
    for (int g = 0; g * 8 < nb; ++g) {
      float dq, dh;
      sym_dist8<D32, NW>(ws.stage + g * 8 * D, nb - g * 8, sv.cfg.measure, sv.q, sv.h, sv.q_norm, sv.h_norm, dq, dh);
      if ((r >> 3) == g) {
        mine_q = dq;
        mine_h = dh;
      }
    }
  }
  else {
    for (int i = 0; i < nb; ++i) {
      float dq, dh;
      sym_dist_generic(sv.cfg, ws.stage + static_cast<size_t>(i) * D, sv.s_q, sv.s_h, sv.q_norm, sv.h_norm, dq, dh);
      if (r == i) {
        mine_q = dq;
        mine_h = dh;
      }
    }
  }
  __syncwarp();
}

template <int NS, bool FAST, int D32, int NW>
__global__ void __launch_bounds__(CW * 32, FAST ? G200_SYM_MB : 1) sym_kernel(const __grid_constant__ SymArgs a)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  // serial mode (testing, GGNN_B200_SYM_SERIAL=1): ONE warp walks all points in order -- the deterministic schedule of
  // the CPU oracle (the reference's own order is a race: cross-block atomics and reads of a buffer being written)
  uint32_t n = a.serial ? (blockIdx.x == 0 && warp == 0 ? 0u : a.N_layer) : blockIdx.x * CW + warp;
  if (n >= a.N_layer) return;
  unsigned char* wbase = smem_raw + static_cast<size_t>(warp) * a.pl.warp_smem_bytes;
  WarpSmem ws;
  VisitedSet V;
  init_warp_smem(ws, V, wbase, a.pl, a.cache - a.sorted);
  if (a.pl.stage_mode == 3) {
    ws.tmap = &a.tmap;
    ws.pad_row = a.pad_row;
  }

  const uint32_t K = a.KBuild, KF = K / 2, KL = K - KF;
  const float mean_nn1 = a.nn1_stats[0];
  const float xi = (a.measure == 0) ? __fmul_rn(__fmul_rn(__fmul_rn(mean_nn1, mean_nn1), a.tau_build), a.tau_build)
                                    : __fmul_rn(mean_nn1, a.tau_build);
  unsigned long long tot_pops = 0, tot_dists = 0, tot_points = 0;
  for (; n < a.N_layer; n += a.serial ? 1u : a.N_layer) {
  ++tot_points;
  if (a.serial) {
    __threadfence();  // the previous point's sym_buffer / sym_atomic writes (lane 0) are visible to every lane's loads
    __syncwarp();
  }
  const int m = a.translation ? a.translation[n] : static_cast<int>(n);
  SymVec<FAST, D32, NW> sv;
  sv.cfg = DistCfg{a.D, a.VB, a.items, a.measure};
  sv.s_q = ws.s_q;
  sv.s_h = reinterpret_cast<float*>(wbase + a.pl.off_half);
  sv.q_norm = 0.f;
  sv.h_norm = 0.f;
  const float* gq = a.base + static_cast<size_t>(m) * a.D;
  if constexpr (FAST) {
#pragma unroll
    for (int c = 0; c < D32; ++c) sv.q[c] = gq[lane + 32 * c];
  }
  else {
    for (uint32_t d = lane; d < a.D; d += 32) sv.s_q[d] = gq[d];
    __syncwarp();
  }
  constexpr float HALF_W = 0.5f - 0.1f;  // simple_knn_sym_cache.cuh:39,171

  WarpLists<NS> L;
  for (uint32_t k = 0; k < KL; ++k) {
    const int start_n = __ldg(a.graph + static_cast<size_t>(n) * K + k);
    if (start_n < 0) continue;  // (the reference would read out of bounds)
    const int start_m = a.translation ? a.translation[start_n] : start_n;
    const float* gs = a.base + static_cast<size_t>(start_m) * a.D;
    // init_start_point :159-201
    if constexpr (FAST) {
      float tq = 0.f, th = 0.f;
#pragma unroll
      for (int c = 0; c < D32; ++c) sv.h[c] = fmaf(gs[lane + 32 * c] - sv.q[c], HALF_W, sv.q[c]);
      if (a.measure != 0) {
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          float aq = 0.f, ah = 0.f;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int c = it * NW + w;
            if (c < D32) {
              aq = fmaf(sv.q[c], sv.q[c], aq);
              ah = fmaf(sv.h[c], sv.h[c], ah);
            }
          }
          aq = warp_tree_sum(aq);
          ah = warp_tree_sum(ah);
          tq = (w == 0) ? aq : tq + aq;
          th = (w == 0) ? ah : th + ah;
        }
        sv.q_norm = tq;
        sv.h_norm = th;
      }
    }
    else {
      __syncwarp();
      for (uint32_t d = lane; d < a.D; d += 32) sv.s_h[d] = fmaf(gs[d] - sv.s_q[d], HALF_W, sv.s_q[d]);
      __syncwarp();
      if (a.measure != 0) {
        float tq = 0.f, th = 0.f;
        for (uint32_t w = 0; w < a.VB / 32; ++w) {
          float aq = 0.f, ah = 0.f;
          for (uint32_t it = 0; it < a.items; ++it) {
            const uint32_t d = it * a.VB + 32 * w + lane;
            if (d < a.D) {
              aq = fmaf(sv.s_q[d], sv.s_q[d], aq);
              ah = fmaf(sv.s_h[d], sv.s_h[d], ah);
            }
          }
          aq = warp_tree_sum(aq);
          ah = warp_tree_sum(ah);
          tq = (w == 0) ? aq : tq + aq;
          th = (w == 0) ? ah : th + ah;
        }
        sv.q_norm = tq;
        sv.h_norm = th;
      }
    }
    float dq0 = 0.f, dh0 = 0.f;
    {
      float mq = 0.f, mh = 0.f;
      sym_stage_and_dist<FAST, D32, NW>(ws, sv, a.base, start_m, 0, 1, mq, mh);
      dq0 = __shfl_sync(FULL, mq, 0);
      dh0 = __shfl_sync(FULL, mh, 0);
    }
    const float crit_half = dh0 + xi;
    L.init(KF);
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const uint32_t p = 32u * j + lane;
      if (p == 0 || p == KF) {
        L.key[j] = start_n;
        L.dist[j] = dq0;
      }
    }
    V.clear();

    bool found = false;
    for (uint32_t ite = 0; ite < a.max_iterations && !found; ++ite) {
      const float crit = L.dist_at(0) + xi;  // criteria_sym :285-288
      const int anchor = L.pop(crit);
      if (anchor == EMPTY_KEY) break;
      V.insert(anchor);
      ++tot_pops;
      for (uint32_t i = 0; i < K; i += 32) {
        const uint32_t kk = i + lane;
        int ck = EMPTY_KEY;
        if (kk < K)
          ck = (kk < KL) ? __ldg(a.graph + static_cast<size_t>(anchor) * K + kk)
                         : __ldcg(a.sym_buffer + static_cast<size_t>(anchor) * KF + kk - KL);
        if (__any_sync(FULL, ck == static_cast<int>(n))) {  // sym_query_layer.cu:105-119
          found = true;
          break;
        }
        // fetch (simple_knn_sym_cache.cuh:405-436)
        bool valid = ck != EMPTY_KEY;
        __syncwarp();
        L.store_keys(ws.s_sorted);
        __syncwarp();
        if (valid) valid = !WarpLists<NS>::in_sorted(ws.s_sorted, ck) && !V.contains(ck);
        const unsigned mask = __ballot_sync(FULL, valid);
        const int cnt = __popc(mask);
        if (!cnt) continue;
        tot_dists += cnt;
        __syncwarp();
        if (valid) ws.s_sorted[__popc(mask & ((1u << lane) - 1u))] = ck;
        __syncwarp();
        const int key_r = ws.s_sorted[lane];
        int mm = 0;
        if (lane < cnt) mm = a.translation ? a.translation[key_r] : key_r;
        float mine_q = G200_INF, mine_h = G200_INF;
        for (int b0 = 0; b0 < cnt; b0 += ws.stage_rows) {
          const int nb = min(static_cast<int>(ws.stage_rows), cnt - b0);
          float dq = G200_INF, dh = G200_INF;
          sym_stage_and_dist<FAST, D32, NW>(ws, sv, a.base, mm, b0, nb, dq, dh);
          if (lane >= b0 && lane < b0 + nb) {
            mine_q = dq;
            mine_h = dh;
          }
        }
        unsigned rem = cnt >= 32 ? FULL : ((1u << cnt) - 1u);
        while (true) {
          const float c0 = L.dist_at(0) + xi;
          const unsigned pm = __ballot_sync(FULL, mine_q < c0 && mine_h < crit_half) & rem;
          if (!pm) break;
          const int r = __ffs(pm) - 1;
          L.push(__shfl_sync(FULL, key_r, r), __shfl_sync(FULL, mine_q, r));
          rem &= ~((2u << r) - 1u);
        }
      }
    }
    if (!found) {  // sym_query_layer.cu:121-141
      for (uint32_t i = 0; i < KF; ++i) {
        const int other_n = L.key_at(i);
        if (other_n == EMPTY_KEY) break;
        uint32_t pos = 0;
        if (lane == 0) pos = atomicAdd(&a.sym_atomic[other_n], 1u);
        pos = __shfl_sync(FULL, pos, 0);
        if (pos < KF) {
          if (lane == 0) a.sym_buffer[static_cast<size_t>(other_n) * KF + pos] = static_cast<int>(n);
          break;
        }
      }
      if (a.serial) {
        __threadfence();
        __syncwarp();
      }
    }
  }
  }  // points of this warp (one, unless serial)
  if (a.counters && lane == 0) {
    atomicAdd(&a.counters[0], tot_pops);
    atomicAdd(&a.counters[1], tot_dists + tot_points * (K - K / 2));  // + the start point of every local link's search
    atomicAdd(&a.counters[2], tot_points);
  }
}

// ================================================================================================
// sym_buffer_merge: one thread per point
// ================================================================================================
__global__ void __launch_bounds__(128) sym_buffer_merge_kernel(uint32_t N, uint32_t K, const int32_t* __restrict__ sym_buffer,
                                                               const uint32_t* __restrict__ sym_atomic, int32_t* graph)
{
  const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const uint32_t KF = K / 2, KL = K - KF;
  uint32_t num_links = sym_atomic[n];
  const int32_t* sb = sym_buffer + static_cast<size_t>(n) * KF;
  int32_t* gr = graph + static_cast<size_t>(n) * K + KL;
  // Requested inverse links occupy sb[0 .. min(num_links,KF)); existing foreign links gr[i] that are
  // not among the KF buffer entries are appended while there is room (:61-86).  The appended entries
  // are written straight to their final slot: slots < original count keep the requested links.
  const uint32_t n_req = min(num_links, KF);
  int32_t out[256];
  for (uint32_t kf = 0; kf < KF; ++kf) out[kf] = sb[kf];
  for (uint32_t i = 0; i < KF; ++i) {
    bool found = num_links >= KF;
    const int32_t r_graph = gr[i];
    if (!found)
      for (uint32_t kf = 0; kf < KF; ++kf) found |= (out[kf] == r_graph);
    if (!found) {
      out[num_links] = r_graph;
      ++num_links;
    }
  }
  (void)n_req;
  for (uint32_t kf = 0; kf < KF; ++kf) gr[kf] = (out[kf] >= 0) ? out[kf] : static_cast<int32_t>(n);
}

// ================================================================================================
// host launchers
// ================================================================================================
struct FastSel {
  bool fast;
  int d32, nw;
};
static FastSel fast_sel(uint32_t D, uint32_t VB, uint32_t items)
{
  FastSel f{false, 1, 1};
  if (D % 32 == 0 && D <= 128 && items == 4 && (VB == 32 || VB == 64 || VB == 128) && D <= VB * 4) {
    f.fast = true;
    f.d32 = D / 32;
    f.nw = VB / 32;
  }
  return f;
}
static void construction_config(uint32_t D, uint32_t min_block, uint32_t& VB, uint32_t& items)
{  // graph_construction.cu:154-161
  items = D <= 1024 ? 4 : 8;
  VB = std::max(min_block, bit_ceil_u32((D + items - 1) / items));
}

template <typename Kern, typename Args>
static int launch_warp_kernel(Kern kern, const Args& a, uint32_t n_items, size_t smem, cudaStream_t stream, const char* what,
                              bool persistent = false)
{
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return set_cuda_error(e, what);
  uint32_t grid = (n_items + CW - 1) / CW;
  if (persistent && env_u32("GGNN_B200_BUILD_PERSISTENT", 1)) {  // one wave of resident CTAs, each warp loops over its points
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CW * 32, smem);
    if (e != cudaSuccess) return set_cuda_error(e, what);
    grid = std::min<uint32_t>(grid, static_cast<uint32_t>(std::max(per_sm, 1)) * device_info().num_sms);
  }
  kern<<<grid, CW * 32, smem, stream>>>(a);
  return set_cuda_error(cudaGetLastError(), what);
}

}  // namespace g200

using namespace g200;

// ---- diagnostics: per-launch traversal counters + times of the construction kernels (thread-local recorder) ----
namespace {
struct BuildRecorder {
  static constexpr uint32_t MAX = GGNN_B200_MAX_BUILD_PASSES;
  unsigned long long* d_counters{nullptr};  // [MAX][4]
  cudaEvent_t ev[MAX][2];
  ggnn_b200_build_pass_stats meta[MAX];
  uint32_t n{0};
};
thread_local BuildRecorder* g_rec = nullptr;

// -> device counters of the next recorded pass (nullptr when not recording); the caller launches between begin / end
unsigned long long* rec_begin(uint32_t kernel, uint32_t layer_top, uint32_t layer_btm, uint32_t points, cudaStream_t stream)
{
  BuildRecorder* r = g_rec;
  if (!r || r->n >= BuildRecorder::MAX) return nullptr;
  ggnn_b200_build_pass_stats& m = r->meta[r->n];
  m = ggnn_b200_build_pass_stats{};
  m.kernel = kernel;
  m.layer_top = layer_top;
  m.layer_btm = layer_btm;
  m.points = points;
  cudaEventRecord(r->ev[r->n][0], stream);
  return r->d_counters + 4 * static_cast<size_t>(r->n);
}
void rec_end(unsigned long long* slot, cudaStream_t stream)
{
  if (!slot) return;
  BuildRecorder* r = g_rec;
  cudaEventRecord(r->ev[r->n][1], stream);
  ++r->n;
}
}  // namespace

extern "C" int ggnn_b200_build_stats_begin(void)
{
  if (g_rec) return set_error(GGNN_B200_ERR_INVALID, "build statistics are already being recorded on this thread");
  BuildRecorder* r = new BuildRecorder();
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&r->d_counters), BuildRecorder::MAX * 4 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(r->d_counters, 0, BuildRecorder::MAX * 4 * sizeof(unsigned long long));
  for (uint32_t i = 0; i < BuildRecorder::MAX && e == cudaSuccess; ++i)
    for (int j = 0; j < 2 && e == cudaSuccess; ++j) e = cudaEventCreate(&r->ev[i][j]);
  if (e != cudaSuccess) {
    delete r;
    return set_cuda_error(e, "build statistics set-up");
  }
  g_rec = r;
  return 0;
}

extern "C" int ggnn_b200_build_stats_end(ggnn_b200_build_pass_stats* out, uint32_t max_passes, uint32_t* n_passes)
{
  BuildRecorder* r = g_rec;
  if (!r) return set_error(GGNN_B200_ERR_INVALID, "build statistics are not being recorded on this thread");
  g_rec = nullptr;
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[BuildRecorder::MAX * 4];
  if (e == cudaSuccess) e = cudaMemcpy(h, r->d_counters, sizeof(h), cudaMemcpyDeviceToHost);
  const uint32_t n = std::min(r->n, max_passes);
  for (uint32_t i = 0; i < n && e == cudaSuccess && out; ++i) {
    out[i] = r->meta[i];
    out[i].pops = h[4 * i];
    out[i].dists = h[4 * i + 1];
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r->ev[i][0], r->ev[i][1]);
    out[i].ms = ms;
  }
  if (n_passes) *n_passes = n;
  for (uint32_t i = 0; i < BuildRecorder::MAX; ++i)
    for (int j = 0; j < 2; ++j) cudaEventDestroy(r->ev[i][j]);
  cudaFree(r->d_counters);
  delete r;
  return set_cuda_error(e, "build statistics read-back");
}

static int check_cfg(const ggnn_b200_graph_config* cfg)
{
  if (!cfg) return set_error(GGNN_B200_ERR_INVALID, "null graph config");
  // KF + 16 must stay below the sym cache (128 slots): the reference aborts on CHECK_LT(sorted_size, CACHE_SIZE)
  // (include/ggnn/construction/sym_query_layer.cuh:46,58-59) for KBuild > 161
  if (cfg->KBuild > 161)
    return set_error(GGNN_B200_ERR_INVALID, "KBuild > 161: sym sorted_size >= CACHE_SIZE (the reference CHECK-aborts)");
  return 0;
}

extern "C" int ggnn_b200_top(const ggnn_b200_graph_config* cfg, const float* d_base, int32_t measure, uint32_t layer,
                             void* d_graph_blob, float* d_nn1, ggnn_b200_stream_t stream_)
{
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = check_cfg(cfg)) return rc;
  if (layer >= GGNN_B200_L || !d_base || !d_graph_blob || !d_nn1) return set_error(GGNN_B200_ERR_INVALID, "bad argument");
  const GraphPtrs g = graph_ptrs(*cfg, d_graph_blob);
  TopArgs a{};
  a.D = cfg->D;
  a.KBuild = cfg->KBuild;
  a.N_layer = cfg->Ns[layer];
  a.layer = layer;
  a.S = layer ? cfg->S : cfg->S0;
  a.S_offset = layer ? 0 : cfg->S0_off;
  construction_config(cfg->D, 128, a.VB, a.items);
  a.measure = measure;
  a.base = d_base;
  a.translation = layer ? g.translation + cfg->STs_offsets[layer] : nullptr;
  a.graph = g.graph + static_cast<size_t>(cfg->Ns_offsets[layer]) * cfg->KBuild;
  a.nn1 = d_nn1;
  FastSel f = fast_sel(a.D, a.VB, a.items);
  const int NSK = (a.KBuild + 31) / 32;
  f.fast = f.fast && f.nw == 4 && NSK <= 2;
  if (int rc = make_plan(a.pl, a.D, !f.fast, false, 0, 0, 0, 16)) return rc;
  const size_t smem = static_cast<size_t>(a.pl.warp_smem_bytes) * CW;
#define G200_TOP(NSK_, FAST_, D32_, NW_) \
  return launch_warp_kernel(top_kernel<NSK_, FAST_, D32_, NW_>, a, a.N_layer, smem, stream, "top_kernel")
  if (f.fast) {
    switch (NSK * 10 + f.d32) {
      case 11: G200_TOP(1, true, 1, 4);
      case 12: G200_TOP(1, true, 2, 4);
      case 13: G200_TOP(1, true, 3, 4);
      case 14: G200_TOP(1, true, 4, 4);
      case 21: G200_TOP(2, true, 1, 4);
      case 22: G200_TOP(2, true, 2, 4);
      case 23: G200_TOP(2, true, 3, 4);
      case 24: G200_TOP(2, true, 4, 4);
    }
  }
  switch (NSK) {
    case 1: G200_TOP(1, false, 1, 1);
    case 2: G200_TOP(2, false, 1, 1);
    case 3: G200_TOP(3, false, 1, 1);
    case 4: G200_TOP(4, false, 1, 1);
    case 5: G200_TOP(5, false, 1, 1);
    case 6: G200_TOP(6, false, 1, 1);
  }
#undef G200_TOP
  return set_error(GGNN_B200_ERR_UNSUPPORTED, "no top kernel variant");
}

extern "C" int ggnn_b200_nn1_stats(const float* d_nn1, uint32_t N, float* d_nn1_stats, void* d_scratch,
                                   ggnn_b200_stream_t stream_)
{
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d_nn1 || !d_nn1_stats || !d_scratch || !N) return set_error(GGNN_B200_ERR_INVALID, "bad argument");
  float* partial = static_cast<float*>(d_scratch);
  stats_partial_kernel<<<STATS_BLOCKS, 256, 0, stream>>>(d_nn1, N, partial);
  stats_final_kernel<<<1, 32, 0, stream>>>(partial, N, d_nn1_stats);
  return set_cuda_error(cudaGetLastError(), "nn1 stats kernels");
}

extern "C" int ggnn_b200_select(const ggnn_b200_graph_config* cfg, uint32_t layer, const float* d_nn1, const float* d_rng,
                                void* d_graph_blob, ggnn_b200_stream_t stream_)
{
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = check_cfg(cfg)) return rc;
  if (layer >= GGNN_B200_L - 1 || !d_nn1 || !d_rng || !d_graph_blob) return set_error(GGNN_B200_ERR_INVALID, "bad argument");
  const GraphPtrs g = graph_ptrs(*cfg, d_graph_blob);
  SelectArgs a{};
  a.layer = layer;
  a.n_blocks = cfg->Bs[layer];
  a.Sglob = cfg->S;
  a.S = layer ? cfg->S : cfg->S0;
  a.S_offset = layer ? 0 : cfg->S0_off;
  a.G = cfg->G;
  a.SG = cfg->SG;
  a.SG_offset = cfg->SG_off;
  a.nn1 = d_nn1;
  a.rng = d_rng;
  a.selection = g.selection + cfg->STs_offsets[layer + 1];
  a.translation = g.translation + cfg->STs_offsets[layer + 1];
  a.translation_layer = layer ? g.translation + cfg->STs_offsets[layer] : nullptr;
  if (a.S + 1 > 256) return set_error(GGNN_B200_ERR_UNSUPPORTED, "segment size > 255 (the reference's select sorts at most 256 items)");
  const uint32_t grid = (a.n_blocks + CW - 1) / CW;
  select_kernel<<<grid, CW * 32, 0, stream>>>(a);
  return set_cuda_error(cudaGetLastError(), "select_kernel launch");
}

extern "C" int ggnn_b200_merge(const ggnn_b200_graph_config* cfg, const float* d_base, int32_t measure, float tau_build,
                               uint32_t layer_top, uint32_t layer_btm, void* d_graph_blob, int32_t* d_graph_buffer,
                               float* d_nn1, ggnn_b200_stream_t stream_)
{
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = check_cfg(cfg)) return rc;
  if (layer_top >= GGNN_B200_L || layer_btm >= layer_top || !d_base || !d_graph_blob || !d_graph_buffer || !d_nn1)
    return set_error(GGNN_B200_ERR_INVALID, "bad argument (need layer_top > layer_btm, merge_layer.cuh:46)");
  const GraphPtrs g = graph_ptrs(*cfg, d_graph_blob);
  MergeArgs a{};
  a.D = cfg->D;
  a.KBuild = cfg->KBuild;
  a.S = cfg->S;
  a.layer_top = layer_top;
  a.layer_btm = layer_btm;
  a.G = cfg->G;
  a.S0 = cfg->S0;
  a.S0_offset = cfg->S0_off;
  a.N_btm = cfg->Ns[layer_btm];
  construction_config(cfg->D, 32, a.VB, a.items);
  for (int l = 0; l < GGNN_B200_L; ++l) {
    a.Ns_offsets[l] = cfg->Ns_offsets[l];
    a.STs_offsets[l] = cfg->STs_offsets[l];
  }
  a.measure = measure;
  a.tau_build = tau_build;
  a.base = d_base;
  a.selection = g.selection;
  a.translation = g.translation;
  a.graph = g.graph;
  a.graph_buffer = d_graph_buffer;
  a.nn1_stats = g.nn1_stats;
  a.nn1 = d_nn1;
  a.cache = 256;  // merge_layer.cuh:40-42
  a.max_iterations = 200;
  a.sorted = std::max(64u, next_multiple32(cfg->KBuild + 1 + 16));  // :64-65
  if (a.sorted >= a.cache) return set_error(GGNN_B200_ERR_INVALID, "SORTED_SIZE >= CACHE_SIZE");
  FastSel f = fast_sel(a.D, a.VB, a.items);
  const int NS = a.sorted / 32;
  f.fast = f.fast && f.nw == 1 && NS == 2;  // the register-resident variants that are instantiated
  // gather4 staging (two 8-row buffers) for the register-resident variants with 384/512-byte rows
  const bool g4 = f.fast && (f.d32 == 3 || f.d32 == 4) && env_u32("GGNN_B200_BUILD_STAGE_MODE", 3) == 3;
  if (int rc = make_plan(a.pl, a.D, !f.fast, false, a.sorted, a.cache, a.max_iterations, g4 ? 4 * G200_MERGE_MB : 16, g4 ? 16 : 0,
                         g4 && G200_MERGE_MB > 5))
    return rc;
  const size_t smem = static_cast<size_t>(a.pl.warp_smem_bytes) * CW;
  a.counters = rec_begin(0, layer_top, layer_btm, a.N_btm, stream);
  int rc = -1;
#define G200_MERGE(NS_, FAST_, D32_, NW_) \
  rc = launch_warp_kernel(merge_kernel<NS_, FAST_, D32_, NW_>, a, a.N_btm, smem, stream, "merge_kernel", true)
  if (g4 && a.pl.stage_rows == 16 && a.pl.stage_mode == 0 && make_row_gather_tensor_map(&a.tmap, d_base, cfg->N, cfg->D) == 0) {
    a.pad_row = static_cast<int32_t>(cfg->N);  // out of bounds: zero fill, no memory traffic
    if (f.d32 == 3) rc = launch_warp_kernel(merge_kernel<2, true, 3, 1, true>, a, a.N_btm, smem, stream, "merge_kernel", true);
    else rc = launch_warp_kernel(merge_kernel<2, true, 4, 1, true>, a, a.N_btm, smem, stream, "merge_kernel", true);
  }
  else if (f.fast) {
    switch (f.d32) {
      case 1: G200_MERGE(2, true, 1, 1); break;
      case 2: G200_MERGE(2, true, 2, 1); break;
      case 3: G200_MERGE(2, true, 3, 1); break;
      case 4: G200_MERGE(2, true, 4, 1); break;
    }
  }
  else {
    switch (NS) {
      case 2: G200_MERGE(2, false, 1, 1); break;
      case 3: G200_MERGE(3, false, 1, 1); break;
      case 4: G200_MERGE(4, false, 1, 1); break;
      case 5: G200_MERGE(5, false, 1, 1); break;
      case 6: G200_MERGE(6, false, 1, 1); break;
      default: return set_error(GGNN_B200_ERR_UNSUPPORTED, "no merge kernel variant");
    }
  }
#undef G200_MERGE
  rec_end(a.counters, stream);
  if (rc) return rc;
  // publish: graph_construction.cu:290-295
  cudaError_t e = cudaMemcpyAsync(g.graph + static_cast<size_t>(cfg->Ns_offsets[layer_btm]) * cfg->KBuild, d_graph_buffer,
                                  static_cast<size_t>(a.N_btm) * cfg->KBuild * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                                  stream);
  return set_cuda_error(e, "merge publish copy");
}

extern "C" int ggnn_b200_sym(const ggnn_b200_graph_config* cfg, const float* d_base, int32_t measure, float tau_build,
                             uint32_t layer, void* d_graph_blob, int32_t* d_sym_buffer, uint32_t* d_sym_atomic,
                             ggnn_b200_stream_t stream_)
{
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = check_cfg(cfg)) return rc;
  if (layer >= GGNN_B200_L || !d_base || !d_graph_blob || !d_sym_buffer || !d_sym_atomic)
    return set_error(GGNN_B200_ERR_INVALID, "bad argument");
  const GraphPtrs g = graph_ptrs(*cfg, d_graph_blob);
  SymArgs a{};
  a.D = cfg->D;
  a.KBuild = cfg->KBuild;
  a.N_layer = cfg->Ns[layer];
  construction_config(cfg->D, 64, a.VB, a.items);
  a.measure = measure;
  a.tau_build = tau_build;
  a.base = d_base;
  a.graph = g.graph + static_cast<size_t>(cfg->Ns_offsets[layer]) * cfg->KBuild;
  a.translation = layer ? g.translation + cfg->STs_offsets[layer] : nullptr;
  a.nn1_stats = g.nn1_stats;
  a.sym_buffer = d_sym_buffer;
  a.sym_atomic = d_sym_atomic;
  a.cache = 128;  // sym_query_layer.cuh:37-39
  a.max_iterations = 20;
  a.sorted = std::max(64u, next_multiple32(cfg->KBuild / 2 + 16));  // :58-59
  if (a.sorted >= a.cache) return set_error(GGNN_B200_ERR_INVALID, "sorted_size >= CACHE_SIZE");
  const uint32_t KF = cfg->KBuild / 2;
  // graph_construction.cu:303-307
  cudaError_t e = cudaMemsetAsync(d_sym_buffer, 0xff, static_cast<size_t>(a.N_layer) * KF * sizeof(int32_t), stream);
  if (e != cudaSuccess) return set_cuda_error(e, "memset sym_buffer");
  e = cudaMemsetAsync(d_sym_atomic, 0, static_cast<size_t>(a.N_layer) * sizeof(uint32_t), stream);
  if (e != cudaSuccess) return set_cuda_error(e, "memset sym_atomic");
  FastSel f = fast_sel(a.D, a.VB, a.items);
  const int NS = a.sorted / 32;
  f.fast = f.fast && f.nw == 2 && NS == 2;
  if (int rc = make_plan(a.pl, a.D, !f.fast, !f.fast, a.sorted, a.cache, a.max_iterations, env_u32("GGNN_B200_SYM_WARPS_PER_SM", 4 * G200_SYM_MB))) return rc;
  if (f.fast && a.pl.stage_mode == 0 && a.D <= 256 && env_u32("GGNN_B200_BUILD_STAGE_MODE", 3) == 3 &&
      make_row_gather_tensor_map(&a.tmap, d_base, cfg->N, cfg->D) == 0) {
    a.pad_row = static_cast<int32_t>(cfg->N);  // out of bounds: zero fill, no memory traffic
    a.pl.stage_mode = 3;
  }
  const size_t smem = static_cast<size_t>(a.pl.warp_smem_bytes) * CW;
  a.serial = env_u32("GGNN_B200_SYM_SERIAL", 0) ? 1u : 0u;
  a.counters = rec_begin(1, layer, layer, a.N_layer, stream);
  int rc = -1;
#define G200_SYM(NS_, FAST_, D32_, NW_) \
  rc = launch_warp_kernel(sym_kernel<NS_, FAST_, D32_, NW_>, a, a.serial ? 1u : a.N_layer, smem, stream, "sym_kernel")
  if (f.fast) {
    switch (f.d32) {
      case 1: G200_SYM(2, true, 1, 2); break;
      case 2: G200_SYM(2, true, 2, 2); break;
      case 3: G200_SYM(2, true, 3, 2); break;
      case 4: G200_SYM(2, true, 4, 2); break;
    }
  }
  else {
    switch (NS) {
      case 2: G200_SYM(2, false, 1, 1); break;
      case 3: G200_SYM(3, false, 1, 1); break;
      default: rc = set_error(GGNN_B200_ERR_UNSUPPORTED, "no sym kernel variant");
    }
  }
#undef G200_SYM
  rec_end(a.counters, stream);
  return rc;
}

extern "C" int ggnn_b200_sym_buffer_merge(const ggnn_b200_graph_config* cfg, uint32_t layer, const int32_t* d_sym_buffer,
                                          const uint32_t* d_sym_atomic, void* d_graph_blob, ggnn_b200_stream_t stream_)
{
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = check_cfg(cfg)) return rc;
  if (layer >= GGNN_B200_L || !d_sym_buffer || !d_sym_atomic || !d_graph_blob) return set_error(GGNN_B200_ERR_INVALID, "bad argument");
  const GraphPtrs g = graph_ptrs(*cfg, d_graph_blob);
  const uint32_t N = cfg->Ns[layer];
  sym_buffer_merge_kernel<<<(N + 127) / 128, 128, 0, stream>>>(
      N, cfg->KBuild, d_sym_buffer, d_sym_atomic, g.graph + static_cast<size_t>(cfg->Ns_offsets[layer]) * cfg->KBuild);
  return set_cuda_error(cudaGetLastError(), "sym_buffer_merge_kernel launch");
}

// scratch layout for build_graph
namespace {
struct Scratch {
  float* nn1;
  int32_t* graph_buffer;
  float* rng;
  int32_t* sym_buffer;
  uint32_t* sym_atomic;
  void* stats;
  size_t total;
};
Scratch scratch_layout(const ggnn_b200_graph_config& c, void* basep)
{
  char* b = static_cast<char*>(basep);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = b + off;
    off += (bytes + 255) / 256 * 256;
    return p;
  };
  Scratch s;
  const size_t N = c.N;
  s.nn1 = reinterpret_cast<float*>(take(N * 4));
  s.graph_buffer = reinterpret_cast<int32_t*>(take(N * c.KBuild * 4));
  s.rng = reinterpret_cast<float*>(take(N * 4));
  s.sym_buffer = reinterpret_cast<int32_t*>(take(N * (c.KBuild / 2) * 4));
  s.sym_atomic = reinterpret_cast<uint32_t*>(take(N * 4));
  s.stats = take(8192);
  s.total = off;
  return s;
}
}  // namespace

extern "C" size_t ggnn_b200_build_scratch_bytes(const ggnn_b200_graph_config* cfg)
{
  return scratch_layout(*cfg, nullptr).total;
}

extern "C" int ggnn_b200_build_graph(const ggnn_b200_graph_config* cfg, const float* d_base, int32_t measure,
                                     float tau_build, uint32_t refinement_iterations, const float* d_rng,
                                     void* d_graph_blob, void* d_scratch, size_t scratch_bytes, ggnn_b200_stream_t stream_)
{
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = check_cfg(cfg)) return rc;
  if (!d_base || !d_graph_blob || !d_scratch) return set_error(GGNN_B200_ERR_INVALID, "null pointer");
  if (scratch_bytes < ggnn_b200_build_scratch_bytes(cfg)) return set_error(GGNN_B200_ERR_INVALID, "scratch too small");
  const Scratch s = scratch_layout(*cfg, d_scratch);
  const GraphPtrs g = graph_ptrs(*cfg, d_graph_blob);
  constexpr uint32_t L = GGNN_B200_L;

  curandGenerator_t gen = nullptr;
  if (!d_rng) {  // graph_construction.cu:96-102
    if (curandCreateGenerator(&gen, CURAND_RNG_PSEUDO_DEFAULT) != CURAND_STATUS_SUCCESS ||
        curandSetPseudoRandomGeneratorSeed(gen, 1234ULL) != CURAND_STATUS_SUCCESS ||
        curandSetStream(gen, stream) != CURAND_STATUS_SUCCESS) {
      if (gen) curandDestroyGenerator(gen);
      return set_error(GGNN_B200_ERR_INVALID, "cuRAND generator setup failed");
    }
  }
  const float* rng_cursor = d_rng;
  int rc = 0;
  auto sym_pass = [&](uint32_t layer) {
    if ((rc = ggnn_b200_sym(cfg, d_base, measure, tau_build, layer, d_graph_blob, s.sym_buffer, s.sym_atomic, stream))) return;
    rc = ggnn_b200_sym_buffer_merge(cfg, layer, s.sym_buffer, s.sym_atomic, d_graph_blob, stream);
  };
  auto merge_step = [&](uint32_t top, uint32_t btm) {
    if (top == btm) rc = ggnn_b200_top(cfg, d_base, measure, btm, d_graph_blob, s.nn1, stream);
    else rc = ggnn_b200_merge(cfg, d_base, measure, tau_build, top, btm, d_graph_blob, s.graph_buffer, s.nn1, stream);
    if (!rc && !btm) rc = ggnn_b200_nn1_stats(s.nn1, cfg->N, g.nn1_stats, s.stats, stream);
  };
  // build: graph_construction.cu:128-140
  for (uint32_t top = 0; top < L && !rc; ++top) {
    for (uint32_t btm = top; btm != 0xffffffffu && !rc; --btm) {
      merge_step(top, btm);
      if (rc) break;
      if (top < L - 1 && top == btm) {
        const float* r = rng_cursor;
        if (gen) {
          if (curandGenerateUniform(gen, s.rng, cfg->Ns[top]) != CURAND_STATUS_SUCCESS) {
            rc = set_error(GGNN_B200_ERR_INVALID, "curandGenerateUniform failed");
            break;
          }
          r = s.rng;
        }
        else rng_cursor += cfg->Ns[top];
        if ((rc = ggnn_b200_select(cfg, top, s.nn1, r, d_graph_blob, stream))) break;
      }
      sym_pass(btm);
    }
  }
  // refine: :141-147
  for (uint32_t it = 0; it < refinement_iterations && !rc; ++it)
    rc = ggnn_b200_refine_graph(cfg, d_base, measure, tau_build, d_graph_blob, d_scratch, scratch_bytes, stream_);
  if (gen) {
    cudaStreamSynchronize(stream);  // the generator must outlive its queued work
    curandDestroyGenerator(gen);
  }
  return rc;
}

extern "C" int ggnn_b200_refine_graph(const ggnn_b200_graph_config* cfg, const float* d_base, int32_t measure, float tau_build,
                                      void* d_graph_blob, void* d_scratch, size_t scratch_bytes, ggnn_b200_stream_t stream)
{
  if (int rc = check_cfg(cfg)) return rc;
  if (!d_base || !d_graph_blob || !d_scratch) return set_error(GGNN_B200_ERR_INVALID, "null pointer");
  if (scratch_bytes < ggnn_b200_build_scratch_bytes(cfg)) return set_error(GGNN_B200_ERR_INVALID, "scratch too small");
  const Scratch s = scratch_layout(*cfg, d_scratch);
  const GraphPtrs g = graph_ptrs(*cfg, d_graph_blob);
  constexpr uint32_t L = GGNN_B200_L;
  for (uint32_t layer = L - 2; layer != 0xffffffffu; --layer) {
    if (int rc = ggnn_b200_merge(cfg, d_base, measure, tau_build, L - 1, layer, d_graph_blob, s.graph_buffer, s.nn1, stream)) return rc;
    if (!layer)
      if (int rc = ggnn_b200_nn1_stats(s.nn1, cfg->N, g.nn1_stats, s.stats, stream)) return rc;
    if (int rc = ggnn_b200_sym(cfg, d_base, measure, tau_build, layer, d_graph_blob, s.sym_buffer, s.sym_atomic, stream)) return rc;
    if (int rc = ggnn_b200_sym_buffer_merge(cfg, layer, s.sym_buffer, s.sym_atomic, d_graph_blob, stream)) return rc;
  }
  return 0;
}

struct ggnn_b200_rng {
  curandGenerator_t gen;
};

extern "C" int ggnn_b200_rng_create(ggnn_b200_rng** out, uint64_t seed)
{
  if (!out) return set_error(GGNN_B200_ERR_INVALID, "null pointer");
  curandGenerator_t gen = nullptr;
  if (curandCreateGenerator(&gen, CURAND_RNG_PSEUDO_DEFAULT) != CURAND_STATUS_SUCCESS ||
      curandSetPseudoRandomGeneratorSeed(gen, seed) != CURAND_STATUS_SUCCESS) {
    if (gen) curandDestroyGenerator(gen);
    return set_error(GGNN_B200_ERR_INVALID, "cuRAND generator setup failed");
  }
  *out = new ggnn_b200_rng{gen};
  return 0;
}

extern "C" int ggnn_b200_rng_fill_build(ggnn_b200_rng* rng, const ggnn_b200_graph_config* cfg, float* d_rng, ggnn_b200_stream_t stream)
{
  if (!rng || !cfg || !d_rng) return set_error(GGNN_B200_ERR_INVALID, "null pointer");
  if (curandSetStream(rng->gen, static_cast<cudaStream_t>(stream)) != CURAND_STATUS_SUCCESS)
    return set_error(GGNN_B200_ERR_INVALID, "curandSetStream failed");
  size_t off = 0;
  for (uint32_t layer = 0; layer + 1 < GGNN_B200_L; ++layer) {  // one call per select(), like the reference
    if (curandGenerateUniform(rng->gen, d_rng + off, cfg->Ns[layer]) != CURAND_STATUS_SUCCESS)
      return set_error(GGNN_B200_ERR_INVALID, "curandGenerateUniform failed");
    off += cfg->Ns[layer];
  }
  return 0;
}

extern "C" void ggnn_b200_rng_destroy(ggnn_b200_rng* rng)
{
  if (!rng) return;
  cudaDeviceSynchronize();  // the generator must outlive its queued work
  curandDestroyGenerator(rng->gen);
  delete rng;
}
