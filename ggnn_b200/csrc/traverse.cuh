// traverse.cuh -- the per-warp "fetch" step shared by the query, merge and sym kernels:
// filter candidate ids against the cache, stage the surviving base rows into shared memory with
// one bulk async copy per row (all in flight together), compute all distances in the reference's
// summation order, then apply the pushes strictly in the reference's candidate order
// (include/ggnn/cuda_utils/simple_knn_cache.cuh:241-289).
#pragma once
#include "common.cuh"

#ifndef G200_PACKED_DIST
#define G200_PACKED_DIST 1  // Euclidean distances of row pairs in packed fp32 (FFMA2 / FADD2)
#endif

namespace g200 {

// 1.0f and -0.0f as values the compiler cannot see (dist8_fast, packed cosine path)
static __constant__ float g200_fma_consts[2] = {1.0f, -0.0f};

// per-warp shared-memory working set
struct WarpSmem {
  float* stage;      // [stage_rows * D], 16-byte aligned
  float* s_q;        // [D] query copy (generic distance path / sym half-way point), may be null
  int* s_sorted;     // [32*NS] mirror of the sorted keys for the filter (register lists); >= 32 ints of scratch
  uint64_t* bar;     // mbarriers for the bulk copies: bar[0..3], one per 8-row stage group
  uint32_t parity;   // phase bit of bar[i] in bit i
  uint32_t stage_rows;  // multiple of 8
  uint32_t stage_mode;  // 0: one cp.async.bulk (TMA engine) per row; 1: 16-byte cp.async per lane (LDGSTS);
                        // 2: no staging -- rows are not 16-byte aligned (D % 4 != 0), distances read global memory
                        // 3: TMA tile::gather4 -- four rows per instruction through a tensor map of the base
  const void* tmap;     // CUtensorMap of the base ([N, D] fp32, box = one row) for stage_mode 3 (G4 kernels only)
  int pad_row;          // row index used to fill a gather4 up to four rows (out of bounds -> zero fill, no traffic)
};

// Move rows [b0, b0+nb) of the candidate list (lane b0+r holds the base row index m of local row r) into
// the stage buffer and wait for them.  All rows are in flight together.
__device__ __forceinline__ void stage_rows_g2s(WarpSmem& ws, const float* __restrict__ base, uint32_t D, int m, int b0, int nb)
{
  const int lane = lane_id();
  const uint32_t row_bytes = D * 4u;
  if (ws.stage_mode == 3 && ws.tmap) {  // TMA tile::gather4: lane b0 + 4j fetches rows b0+4j .. b0+4j+3 with one instruction
    const int r = lane - b0;
    const int mp = (r >= 0 && r < nb) ? m : ws.pad_row;
    const int m1 = __shfl_down_sync(FULL, mp, 1);
    const int m2 = __shfl_down_sync(FULL, mp, 2);
    const int m3 = __shfl_down_sync(FULL, mp, 3);
    if (lane == 0) mbar_expect_tx(ws.bar, row_bytes * static_cast<uint32_t>((nb + 3) & ~3));
    if (r >= 0 && r < nb && (r & 3) == 0) tma_gather4(ws.stage + static_cast<size_t>(r) * D, ws.tmap, mp, m1, m2, m3, ws.bar);
    mbar_wait(ws.bar, ws.parity & 1u);
    ws.parity ^= 1u;
  }
  else if (ws.stage_mode == 0 || ws.stage_mode == 3) {
    if (lane == 0) mbar_expect_tx(ws.bar, row_bytes * nb);
    __syncwarp();
    const int r = lane - b0;
    if (r >= 0 && r < nb) bulk_g2s(ws.stage + static_cast<size_t>(r) * D, base + static_cast<size_t>(m) * D, row_bytes, ws.bar);
    mbar_wait(ws.bar, ws.parity & 1u);
    ws.parity ^= 1u;
  }
  else {
    const uint32_t chunks = row_bytes >> 4;
    for (int r = 0; r < nb; ++r) {
      const int mr = __shfl_sync(FULL, m, b0 + r);
      const char* src = reinterpret_cast<const char*>(base + static_cast<size_t>(mr) * D);
      char* dst = reinterpret_cast<char*>(ws.stage + static_cast<size_t>(r) * D);
      for (uint32_t c = lane; c < chunks; c += 32) cp_async16(dst + c * 16, src + c * 16);
    }
    cp_async_wait_all();
    __syncwarp();
  }
}

struct Stats {
  uint32_t pops, dists;
};

// FAST distance of up to 8 staged rows for D == 32*D32 and a reference block of VB == 32*NW threads
// (4 dims per thread): virtual thread (w, lane) owns the 32-float chunks c = w, w+NW, w+2NW, ...
// (dims 32*c + lane), accumulated in that order; each virtual warp w is tree-reduced, the warp
// aggregates are added sequentially (cub::BlockReduce).  Lane l gets row (l & 7).
template <int D32, int NW, bool IL = false>
__device__ __forceinline__ float dist8_fast(const float* __restrict__ rows, int nrows, int measure,
                                            const float (&q)[D32], float q_norm)
{
  const int lane = lane_id();
  constexpr int D = 32 * D32;
  if constexpr (IL && D32 == 4 && NW == 1) {
    // interleaved rows (ggnn_b200_interleave_rows), Euclidean: this lane's dims lane, lane+32, lane+64, lane+96 of a row are
    // ONE 16-byte load; the FMA chain runs over them in the same order, rows (2p, 2p+1) together in packed fp32
    const float4* r4 = reinterpret_cast<const float4*>(rows) + lane;
    const uint64_t q0 = pack2(q[0], q[0]), q1 = pack2(q[1], q[1]), q2 = pack2(q[2], q[2]), q3 = pack2(q[3], q[3]);
    float v[8];
#pragma unroll
    for (int pr = 0; pr < 4; ++pr) {
      const float4 a = r4[(2 * pr) * (D / 4)], b = r4[(2 * pr + 1) * (D / 4)];
      uint64_t d = sub2(pack2(a.x, b.x), q0);
      uint64_t acc = fma2(d, d, 0ull);
      d = sub2(pack2(a.y, b.y), q1);
      acc = fma2(d, d, acc);
      d = sub2(pack2(a.z, b.z), q2);
      acc = fma2(d, d, acc);
      d = sub2(pack2(a.w, b.w), q3);
      acc = fma2(d, d, acc);
      unpack2(acc, v[2 * pr], v[2 * pr + 1]);
    }
    return warp_tree_sum8(v);
  }
  if (measure == 0 && NW == 1 && G200_PACKED_DIST) {
    // one reference warp, Euclidean: rows (2p, 2p+1) advance together in packed fp32 (per element the same
    // sub.rn + fma.rn chain over dims lane, lane+32, ... as the scalar code).  Always the whole group: rows >= nrows
    // hold stale shared memory, their results are never used.
    uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
    for (int c = 0; c < D32; ++c) {
      const uint64_t q2 = pack2(q[c], q[c]);
#pragma unroll
      for (int pr = 0; pr < 4; ++pr) {
        const float* rp = rows + (2 * pr) * D + 32 * c + lane;
        const uint64_t diff = sub2(pack2(rp[0], rp[D]), q2);
        acc[pr] = fma2(diff, diff, acc[pr]);
      }
    }
    float v[8];
#pragma unroll
    for (int pr = 0; pr < 4; ++pr) unpack2(acc[pr], v[2 * pr], v[2 * pr + 1]);
    return warp_tree_sum8(v);
  }
  // (Cosine stays scalar: the reference's products and sums are separate roundings, and ptxas contracts
  // mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even with explicit .rn -- measured, see DESIGN.md -- which changes results.)
  if (measure == 0) {
    float v[NW][8];
    if (nrows >= 8) {  // full group: straight-line code, no per-row branches
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float* rp = rows + i * D + lane;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          float acc = 0.f;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int c = it * NW + w;
            if (c < D32) {
              const float diff = rp[32 * c] - q[c];
              acc = fmaf(diff, diff, acc);
            }
          }
          v[w][i] = acc;
        }
      }
    }
    else
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int w = 0; w < NW; ++w) v[w][i] = 0.f;
      if (i < nrows) {
        const float* rp = rows + i * D + lane;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          float acc = 0.f;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int c = it * NW + w;
            if (c < D32) {
              const float diff = rp[32 * c] - q[c];
              acc = fmaf(diff, diff, acc);
            }
          }
          v[w][i] = acc;
        }
      }
    }
    float tot = warp_tree_sum8(v[0]);
#pragma unroll
    for (int w = 1; w < NW; ++w) tot = tot + warp_tree_sum8(v[w]);
    return tot;
  }
  if (NW == 1 && G200_PACKED_DIST) {
    // one reference warp, cosine: the reference rounds every product and every sum separately (FMUL + FADD).  ptxas
    // contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even with explicit .rn (and with -fmad=false), so the two
    // roundings are spelled as two FMAs that cannot be fused: a*b = fma(a, b, -0.0) and p + s = fma(p, 1.0, s), each
    // correctly rounded exactly like the scalar instruction.  Rows (2p, 2p+1) advance together.
    // (the two constants come from constant memory: as literals ptxas simplifies the FMAs back to mul / add -- and fuses them)
    const float c_one = g200_fma_consts[0], c_neg0 = g200_fma_consts[1];
    const uint64_t neg0 = pack2(c_neg0, c_neg0), one = pack2(c_one, c_one);
    uint64_t ad[4] = {0ull, 0ull, 0ull, 0ull}, an[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
    for (int c = 0; c < D32; ++c) {
      const uint64_t q2 = pack2(q[c], q[c]);
#pragma unroll
      for (int pr = 0; pr < 4; ++pr) {
        const float* rp = rows + (2 * pr) * D + 32 * c + lane;
        const uint64_t b2 = pack2(rp[0], rp[D]);
        ad[pr] = fma2(fma2(b2, q2, neg0), one, ad[pr]);
        an[pr] = fma2(fma2(b2, b2, neg0), one, an[pr]);
      }
    }
    float vd[8], vn[8];
#pragma unroll
    for (int pr = 0; pr < 4; ++pr) {
      unpack2(ad[pr], vd[2 * pr], vd[2 * pr + 1]);
      unpack2(an[pr], vn[2 * pr], vn[2 * pr + 1]);
    }
    return cosine_finish(warp_tree_sum8(vd), warp_tree_sum8(vn), q_norm);
  }
  float dot_t = 0.f, nrm_t = 0.f;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    float vd[8], vn[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float dot = 0.f, nrm = 0.f;
      if (i < nrows) {
        const float* rp = rows + i * D + lane;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int c = it * NW + w;
          if (c < D32) {
            const float b = rp[32 * c];
            dot = __fadd_rn(dot, __fmul_rn(b, q[c]));  // not fused in the reference (see oracle header)
            nrm = __fadd_rn(nrm, __fmul_rn(b, b));
          }
        }
      }
      vd[i] = dot;
      vn[i] = nrm;
    }
    const float sd = warp_tree_sum8(vd), sn = warp_tree_sum8(vn);
    dot_t = (w == 0) ? sd : dot_t + sd;
    nrm_t = (w == 0) ? sn : nrm_t + sn;
  }
  return cosine_finish(dot_t, nrm_t, q_norm);
}

// Query-side state needed for distances.  FAST: q[c] = query[32*c + lane].
template <bool FAST, int D32, int NW>
struct QueryVec {
  float q[FAST ? D32 : 1];
  float q_norm;
  DistCfg cfg;
  const float* s_q;

  __device__ __forceinline__ void load(const DistCfg& c, const float* __restrict__ g_q, float* smem_q)
  {
    cfg = c;
    const int lane = lane_id();
    q_norm = 0.f;
    if constexpr (FAST) {
      s_q = nullptr;
#pragma unroll
      for (int ch = 0; ch < D32; ++ch) q[ch] = g_q[lane + 32 * ch];
      if (c.measure != 0) {  // distance.cuh:104-117
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          float a = 0.f;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int ch = it * NW + w;
            if (ch < D32) a = fmaf(q[ch], q[ch], a);
          }
          a = warp_tree_sum(a);
          tot = (w == 0) ? a : tot + a;
        }
        q_norm = tot;
      }
    }
    else {
      __syncwarp();
      for (uint32_t d = lane; d < c.D; d += 32) smem_q[d] = g_q[d];
      __syncwarp();
      s_q = smem_q;
      if (c.measure != 0) q_norm = query_norm_generic(c, smem_q);
    }
  }
};

// Stage the rows of candidates [b0, b0+nb) (lane b0+r holds base row index m of local row r) and
// return this lane's distance (valid for lanes in [b0, b0+nb)).  b0 % 8 == 0.
template <bool FAST, int D32, int NW>
__device__ __forceinline__ float stage_and_dist(WarpSmem& ws, const QueryVec<FAST, D32, NW>& qv,
                                                const float* __restrict__ base, int m, int b0, int nb)
{
  const int lane = lane_id();
  const uint32_t D = qv.cfg.D;
  const int r = lane - b0;
  if constexpr (!FAST) {
    if (ws.stage_mode == 2) {
      float mine_g = G200_INF;
      for (int i = 0; i < nb; ++i) {
        const int mi = __shfl_sync(FULL, m, b0 + i);
        float a, b;
        dist_partials_generic(qv.cfg, base + static_cast<size_t>(mi) * D, qv.s_q, a, b);
        const float d = qv.cfg.measure == 0 ? a : cosine_finish(a, b, qv.q_norm);
        if (r == i) mine_g = d;
      }
      return mine_g;
    }
  }
  stage_rows_g2s(ws, base, D, m, b0, nb);

  float mine = G200_INF;
  if constexpr (FAST) {
    for (int g = 0; g * 8 < nb; ++g) {
      const float dg = dist8_fast<D32, NW>(ws.stage + g * 8 * D, nb - g * 8, qv.cfg.measure, qv.q, qv.q_norm);
      if ((r >> 3) == g) mine = dg;
    }
  }
  else {
    for (int i = 0; i < nb; ++i) {
      float a, b;
      dist_partials_generic(qv.cfg, ws.stage + static_cast<size_t>(i) * D, qv.s_q, a, b);
      const float d = qv.cfg.measure == 0 ? a : cosine_finish(a, b, qv.q_norm);
      if (r == i) mine = d;
    }
  }
  __syncwarp();  // all reads of the stage are done before the next batch overwrites it
  return mine;
}

// One fetch of up to 32 candidate ids (ck per lane, EMPTY_KEY = none).
//   FILTER  : drop ids present in best list / prioQ / visited set (simple_knn_cache.cuh:246-261)
//   xi      : criteria() = dist[BEST-1] + xi, re-read after every push (:284)
//   pf_graph/pf_stride: if non-null, the adjacency row of every candidate that passes the criteria at
//             first sight (a likely future anchor) is prefetched into L2 -- the pop -> adjacency load is the
//             serial latency chain of the traversal and DRAM bandwidth is plentiful
// Speculative load of the next anchor's adjacency row (SpecRow): before the (serial, ~2k-cycle) push loop the next
// pop is already predictable -- it is the better of the current prioQ head and the best new candidate -- so its
// adjacency row is requested then and arrives while the pushes run; the caller checks the prediction after pop().
struct SpecRow {
  int key;  // predicted anchor (EMPTY_KEY = none), uniform
  int row;  // lane l: adjacency entry l of `key`
};

// Second half of a fetch, shared by all row formats: lane r < cnt holds candidate key_r with distance `mine`
// (adjacency order).  Speculative load of the next anchor's adjacency row, then the pushes in candidate order.
template <class LT, bool FILTER>
__device__ __forceinline__ void finish_fetch(LT& L, int key_r, float mine, int cnt, float xi, const int* __restrict__ pf_graph,
                                             uint32_t pf_stride, SpecRow* spec)
{
  const int lane = lane_id();
  // duplicates WITHIN this fetch (a graph row may name a point twice): only those can already be in the lists when
  // their turn comes -- every other candidate passed the filter and nothing but its fellow candidates is pushed before it
  unsigned dupmask = FULL;
  if constexpr (FILTER && G200_OPT_DUPSKIP) {
    const unsigned same = __match_any_sync(FULL, lane < cnt ? key_r : (-2 - lane));
    dupmask = __ballot_sync(FULL, (same & ((1u << lane) - 1u)) != 0u);
  }
  float best_last = L.dist_at(L.BEST - 1);  // criteria() = best_last + xi (simple_knn_cache.cuh:284)

  if (pf_graph) {
    const float crit0 = best_last + xi;
    const bool pass = lane < cnt && mine < crit0;
    if (spec && pf_stride <= 32) {
      // predicted next anchor = min(current prioQ head, best passing candidate); distances are >= 0, so their bit
      // patterns order like unsigned integers
      const unsigned mbits = pass ? __float_as_uint(mine) : 0x7f800000u;
      const unsigned best_bits = __reduce_min_sync(FULL, mbits);
      const float head_d = L.dist_at(L.head);
      int pred = L.key_at(L.head);
      if (__uint_as_float(best_bits) < head_d) {
        const unsigned who = __ballot_sync(FULL, pass && mbits == best_bits);
        pred = __shfl_sync(FULL, key_r, __ffs(who) - 1);
      }
      spec->key = pred;
      if (pred != EMPTY_KEY)
        spec->row = (static_cast<uint32_t>(lane) < pf_stride) ? __ldg(pf_graph + static_cast<size_t>(pred) * pf_stride + lane)
                                                            : EMPTY_KEY;
    }
    else if (pass) {
      const int* row = pf_graph + static_cast<size_t>(key_r) * pf_stride;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(row + pf_stride - 1));
    }
  }

  // pushes in candidate order; criteria() is re-read after each push (:284) -- it only changes when the push changed the
  // last entry of the best list, i.e. when d < best_last (it never grows, so the pending mask only loses bits)
  unsigned pm = __ballot_sync(FULL, mine < best_last + xi) & (cnt >= 32 ? FULL : ((1u << cnt) - 1u));
  while (pm) {
    const int c0 = __ffs(pm) - 1;
    const int k = __shfl_sync(FULL, key_r, c0);
    const float d = __shfl_sync(FULL, mine, c0);
    L.push(k, d, (dupmask >> c0) & 1u);
    pm &= pm - 1u;
    if (!G200_OPT_LAZYCRIT || d < best_last) {
      best_last = L.dist_at(L.BEST - 1);
      pm &= __ballot_sync(FULL, mine < best_last + xi);
    }
  }
}

template <class LT, bool FAST, int D32, int NW, bool FILTER, bool G4 = false, bool IL = false>
__device__ __forceinline__ void fetch(LT& L, const VisitedSet& V, WarpSmem& ws,
                                      const QueryVec<FAST, D32, NW>& qv, const float* __restrict__ base,
                                      const int* __restrict__ translation, int ck, float xi, Stats& st,
                                      const int* __restrict__ pf_graph = nullptr, uint32_t pf_stride = 0,
                                      SpecRow* spec = nullptr)
{
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

  // compact: lane r <- r-th surviving candidate (adjacency order preserved), through shared memory
  // (the sorted-key mirror is free again after the filter; __fns would be a ~50-instruction software loop)
  __syncwarp();
  if (valid) ws.s_sorted[__popc(mask & ((1u << lane) - 1u))] = ck;
  __syncwarp();
  const int key_r = ws.s_sorted[lane];
  int m = 0;
  if (lane < cnt) m = translation ? translation[key_r] : key_r;

  float mine = G200_INF;
  if constexpr (G4) {
    // TMA tile::gather4 staging, two 8-row buffers (stage_rows == 16): lane l (l % 4 == 0, l < cnt) fetches candidates
    // l..l+3 with ONE instruction into rows (l & 15) of the stage; group g = candidates 8g..8g+7 lives in buffer g & 1
    // and completes on bar[g & 1].  (An mbarrier phase cannot complete before its single arrival, so the order of
    // expect_tx and the copies does not matter.)
    constexpr int D = 32 * D32;
    constexpr uint32_t QUAD_BYTES = 4u * D * 4u;
    const int ngroups = (cnt + 7) >> 3;
    const int nquads = (cnt + 3) >> 2;
    const int mp = lane < cnt ? m : ws.pad_row;
    const int m1 = __shfl_down_sync(FULL, mp, 1);
    const int m2 = __shfl_down_sync(FULL, mp, 2);
    const int m3 = __shfl_down_sync(FULL, mp, 3);
    const bool issuer = ((lane & 3) == 0) && lane < cnt;
    // (the compaction scratch may overlay the stage buffer: every lane has read its key before the copies may land)
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t bar_s = smem_u32(ws.bar);  // bar[0], bar[1] at bar_s, bar_s + 8
    const uint32_t dst_s = smem_u32(ws.stage) + static_cast<uint32_t>(lane & 15) * (D * 4u);
    const uint32_t my_bar_s = bar_s + 8u * ((lane >> 3) & 1);
    if (lane == 0) {
      mbar_expect_tx_s(bar_s, static_cast<uint32_t>(min(2, nquads)) * QUAD_BYTES);
      if (nquads > 2) mbar_expect_tx_s(bar_s + 8u, static_cast<uint32_t>(min(2, nquads - 2)) * QUAD_BYTES);
    }
    if (issuer && lane < 16) tma_gather4_s(dst_s, ws.tmap, m, m1, m2, m3, my_bar_s);
    for (int g = 0; g < ngroups; ++g) {
      const int buf = g & 1;
      mbar_wait_s(bar_s + 8u * buf, (ws.parity >> buf) & 1u);
      ws.parity ^= 1u << buf;
      const float dg = dist8_fast<D32, NW, IL>(ws.stage + buf * 8 * D, min(8, cnt - 8 * g), qv.cfg.measure, qv.q, qv.q_norm);
      if ((lane >> 3) == g) mine = dg;
      __syncwarp();  // the group's rows have been read: its buffer may be refilled
      if (g + 2 < ngroups) {
        if (lane == 0) mbar_expect_tx_s(bar_s + 8u * buf, static_cast<uint32_t>(min(2, nquads - 2 * (g + 2))) * QUAD_BYTES);
        if (issuer && (lane >> 3) == g + 2) tma_gather4_s(dst_s, ws.tmap, m, m1, m2, m3, my_bar_s);
      }
    }
  }
  else if (FAST && ws.stage_mode == 0 && ws.stage_rows >= 16) {
    // software pipeline over 8-row groups: group g lives in buffer g % NBUF with its own mbarrier, so the copies
    // of the next groups are in flight while the distances of the current one are computed
    constexpr int D = 32 * D32;
    const int nbuf = static_cast<int>(ws.stage_rows >> 3);
    const int ngroups = (cnt + 7) >> 3;
    auto issue = [&](int g) {
      const int buf = g % nbuf;
      const int nr = min(8, cnt - 8 * g);
      const int r = lane - 8 * g;
      if (lane == 0) mbar_expect_tx(&ws.bar[buf], static_cast<uint32_t>(nr) * D * 4u);
      if (r >= 0 && r < nr)
        bulk_g2s(ws.stage + static_cast<size_t>(buf * 8 + r) * D, base + static_cast<size_t>(m) * D, D * 4u, &ws.bar[buf]);
    };
    for (int g = 0; g < min(nbuf, ngroups); ++g) issue(g);
    for (int g = 0; g < ngroups; ++g) {
      const int buf = g % nbuf;
      mbar_wait(&ws.bar[buf], (ws.parity >> buf) & 1u);
      ws.parity ^= 1u << buf;
      if constexpr (FAST) {
        const float dg = dist8_fast<D32, NW>(ws.stage + static_cast<size_t>(buf) * 8 * D, min(8, cnt - 8 * g), qv.cfg.measure, qv.q, qv.q_norm);
        if ((lane >> 3) == g) mine = dg;
      }
      __syncwarp();  // the group's rows have been read: its buffer may be refilled
      if (g + nbuf < ngroups) issue(g + nbuf);
    }
  }
  else {
    for (int b0 = 0; b0 < cnt; b0 += ws.stage_rows) {
      const int nb = min(static_cast<int>(ws.stage_rows), cnt - b0);
      const float d = stage_and_dist<FAST, D32, NW>(ws, qv, base, m, b0, nb);
      if (lane >= b0 && lane < b0 + nb) mine = d;
    }
  }

  finish_fetch<LT, FILTER>(L, key_r, mine, cnt, xi, pf_graph, pf_stride, spec);
}

}  // namespace g200
