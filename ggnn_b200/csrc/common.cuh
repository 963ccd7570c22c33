// common.cuh -- device building blocks shared by the sm_100a GGNN kernels.
//
// Design: one WARP per query / graph point (the reference uses one thread block;
// for every BASELINE config that block is a single warp anyway), several warps per CTA.
//   * best list + priority-queue ring  -> REGISTERS (slot p lives in lane p%32, register p/32),
//     updated with warp shuffles; semantics follow SimpleKNNCache::push/pop slot for slot
//     (include/ggnn/cuda_utils/simple_knn_cache.cuh:126-239), including the ring-wrap quirk.
//   * visited ring                      -> exact open-addressing hash set in shared memory (same
//     membership answers as scanning the reference's ring, simple_knn_cache.cuh:246-261).
//   * neighbour vectors of the popped anchor -> staged into shared memory with one
//     cp.async.bulk (TMA engine, 128-bit vectorised, mbarrier completion) per surviving row, all
//     rows in flight at once; distances are then computed from shared memory in the reference's
//     exact fp32 summation order (include/ggnn/cuda_utils/distance.cuh:119-163: per-thread FFMA
//     chain over dims t, t+B, t+2B, ...; cub::BlockReduce = shfl-down tree 1,2,4,8,16 inside
//     each warp, warp aggregates added sequentially).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// experiment switches of the push path (all on by default; none changes results -- see tools/ab_query.py)
#ifndef G200_OPT_FLAGS
#define G200_OPT_FLAGS 1     // per-lane slot predicates of push() hoisted into a flag word
#endif
#ifndef G200_OPT_DUPSKIP
#define G200_OPT_DUPSKIP 0   // duplicate check only for within-fetch duplicates (found with MATCH.ANY): measured SLOWER
#endif                       // (0.566 vs 0.544 ms per batch) -- the match sits on the serial chain of the traversal
#ifndef G200_OPT_LAZYCRIT
#define G200_OPT_LAZYCRIT 1  // criteria() re-read only after a push that changed the best list
#endif

namespace g200 {

constexpr int EMPTY_KEY = -1;
constexpr int TOMB_KEY = -2;
constexpr unsigned FULL = 0xffffffffu;
#define G200_INF __int_as_float(0x7f800000)

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ------------------------------------------------------------------------------------------------
// mbarrier + bulk async copy (TMA engine, SASS: UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  uint32_t ok = 0;
  const uint32_t a = smem_u32(bar);
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
// one row: global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// four rows of a 2-D tensor (row-major base vectors, box = one whole row) by row index in ONE instruction
// (TMA tile::gather4, sm_100): a quarter of the issue overhead of one bulk copy per row
__device__ __forceinline__ void tma_gather4(void* smem_dst, const void* tmap, int r0, int r1, int r2, int r3, uint64_t* bar)
{
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
      : "memory");
}

// the same on 32-bit shared-memory addresses computed once per warp (no generic -> shared conversion per call)
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar_s, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar_s, uint32_t parity)
{
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar_s), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_gather4_s(uint32_t dst_s, const void* tmap, int r0, int r1, int r2, int r3, uint32_t bar_s)
{
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
          dst_s),
      "l"(tmap), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar_s)
      : "memory");
}

// 16-byte asynchronous copy global -> shared (SASS: LDGSTS.E.BYPASS.128), one warp instruction moves a
// whole 512-byte row; used as the alternative row-staging mode (see traverse.cuh)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// packed fp32 arithmetic (Blackwell FFMA2 / FADD2): two independent IEEE fma.rn / add.rn per instruction, each
// element rounded exactly like the scalar instruction -- used to evaluate two base rows at once
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t pack2(float lo, float hi)
{
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi)
{
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b)
{
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ------------------------------------------------------------------------------------------------
// reference-order reductions
// ------------------------------------------------------------------------------------------------
// cub::WarpReduce (shfl.down tree, offsets 1,2,4,8,16): lane 0's value is the balanced adjacent
// tree over the 32 lane values.  The xor butterfly builds the same tree in every lane (IEEE add is
// commutative), so every lane ends with lane 0's reference value.
__device__ __forceinline__ float warp_tree_sum(float v)
{
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) v = v + __shfl_xor_sync(FULL, v, o);
  return v;
}

// Reduce 8 independent rows at once: v[i] = this lane's partial of row i.  Returns, in lane l, the
// reference-order warp sum of row (l & 7).  9 shuffles instead of 40.
__device__ __forceinline__ float warp_tree_sum8(const float (&v)[8])
{
  const int lane = lane_id();
  const bool b0 = lane & 1, b1 = lane & 2, b2 = lane & 4;
  float w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float keep = b0 ? v[2 * j + 1] : v[2 * j];
    const float send = b0 ? v[2 * j] : v[2 * j + 1];
    w[j] = keep + __shfl_xor_sync(FULL, send, 1);
  }
  float x[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float keep = b1 ? w[2 * j + 1] : w[2 * j];
    const float send = b1 ? w[2 * j] : w[2 * j + 1];
    x[j] = keep + __shfl_xor_sync(FULL, send, 2);
  }
  const float keep = b2 ? x[1] : x[0];
  const float send = b2 ? x[0] : x[1];
  float y = keep + __shfl_xor_sync(FULL, send, 4);
  y = y + __shfl_xor_sync(FULL, y, 8);
  y = y + __shfl_xor_sync(FULL, y, 16);
  return y;
}

// ------------------------------------------------------------------------------------------------
// Distance in the arithmetic order of a VB-thread reference block with ITEMS dims per thread.
//   FAST : VB == 32, D == 32*NI (NI <= 4): the query lives in NI registers per lane.
//   else : generic loops, query read from shared memory (s_q, D floats).
// `row` may point to shared or global memory.
// ------------------------------------------------------------------------------------------------
struct DistCfg {
  uint32_t D;
  uint32_t VB;     // reference BLOCK_DIM_X
  uint32_t items;  // reference DIST_ITEMS_PER_THREAD
  int measure;     // 0 Euclidean, 1 Cosine
};

// generic: one row, any VB (multiple of 32), result valid in all lanes
__device__ __forceinline__ void dist_partials_generic(const DistCfg& c, const float* __restrict__ row,
                                                      const float* __restrict__ s_q, float& out_a, float& out_b)
{
  // Euclidean: out_a = sum (b-q)^2 ; Cosine: out_a = dot, out_b = |b|^2 (unfused, see oracle header)
  const int lane = lane_id();
  float tot_a = 0.f, tot_b = 0.f;
  for (uint32_t w = 0; w < c.VB / 32; ++w) {
    float a = 0.f, b = 0.f;
    for (uint32_t it = 0; it < c.items; ++it) {
      const uint32_t d = it * c.VB + 32 * w + lane;
      if (d < c.D) {
        const float bv = row[d], qv = s_q[d];
        if (c.measure == 0) {
          const float diff = bv - qv;
          a = fmaf(diff, diff, a);
        }
        else {
          a = __fadd_rn(a, __fmul_rn(bv, qv));
          b = __fadd_rn(b, __fmul_rn(bv, bv));
        }
      }
    }
    a = warp_tree_sum(a);
    if (c.measure != 0) b = warp_tree_sum(b);
    tot_a = (w == 0) ? a : tot_a + a;
    tot_b = (w == 0) ? b : tot_b + b;
  }
  out_a = tot_a;
  out_b = tot_b;
}

__device__ __forceinline__ float cosine_finish(float dot, float norm_b, float norm_q)
{
  // distance.cuh:153-158
  const float norm_sqr = __fmul_rn(norm_q, norm_b);
  return (norm_sqr > 0.0f) ? fabsf(1.0f - __fdiv_rn(dot, __fsqrt_rn(norm_sqr))) : 1.0f;
}

// |q|^2 in reference order (distance.cuh:104-117: FFMA chain + block reduce)
__device__ __forceinline__ float query_norm_generic(const DistCfg& c, const float* __restrict__ s_q)
{
  const int lane = lane_id();
  float tot = 0.f;
  for (uint32_t w = 0; w < c.VB / 32; ++w) {
    float a = 0.f;
    for (uint32_t it = 0; it < c.items; ++it) {
      const uint32_t d = it * c.VB + 32 * w + lane;
      const float qv = d < c.D ? s_q[d] : 0.f;
      a = fmaf(qv, qv, a);
    }
    a = warp_tree_sum(a);
    tot = (w == 0) ? a : tot + a;
  }
  return tot;
}

// ------------------------------------------------------------------------------------------------
// Visited set: exact membership of the reference's visited ring, as an open-addressing hash set
// (linear probing) in shared memory; a mirror of the ring itself is only kept when it can wrap
// (ring capacity < number of pops), to know which key the reference forgets.
// ------------------------------------------------------------------------------------------------
struct VisitedSet {
  int* tab;          // [hsize] shared
  uint32_t hmask;    // hsize - 1
  uint32_t hshift;   // 32 - log2(hsize)
  int* ring;         // [vcap] shared or nullptr
  uint32_t vcap;     // CACHE - SORTED
  uint32_t vpos;     // next ring slot (uniform)

  __device__ __forceinline__ uint32_t slot(int key) const
  {
    return (static_cast<uint32_t>(key) * 0x9E3779B1u) >> hshift;
  }
  __device__ __forceinline__ void clear()
  {
    const int lane = lane_id();
    for (uint32_t i = lane; i <= hmask; i += 32) tab[i] = EMPTY_KEY;
    if (ring)
      for (uint32_t i = lane; i < vcap; i += 32) ring[i] = EMPTY_KEY;
    vpos = 0;
    __syncwarp();
  }
  // all lanes call with the same key; lane 0 mutates
  __device__ __forceinline__ void insert(int key)
  {
    if (lane_id() == 0) {
      if (ring) {
        const int old = ring[vpos];
        if (old != EMPTY_KEY) {  // the reference overwrites (forgets) this ring entry
          uint32_t h = slot(old);
          while (tab[h] != old) h = (h + 1) & hmask;
          tab[h] = TOMB_KEY;
        }
        ring[vpos] = key;
      }
      uint32_t h = slot(key);
      while (tab[h] != EMPTY_KEY) h = (h + 1) & hmask;
      tab[h] = key;
    }
    vpos = (vpos + 1 >= vcap) ? 0 : vpos + 1;
    __syncwarp();
  }
  // per-lane lookup (divergent)
  __device__ __forceinline__ bool contains(int key) const
  {
    uint32_t h = slot(key);
    while (true) {
      const int e = tab[h];
      if (e == key) return true;
      if (e == EMPTY_KEY) return false;
      h = (h + 1) & hmask;
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Best list + priority-queue ring in registers.  SORTED == 32*NS.  Physical slot p = 32*j + lane.
// ------------------------------------------------------------------------------------------------
template <int NS>
struct WarpLists {
  int key[NS];
  float dist[NS];
  uint32_t head;  // physical prioQ head (uniform)
  uint32_t BEST;  // uniform
  // per-lane slot predicates of push(), which only change when the head moves (pop / transform): for register j,
  // bit 3j = slot p may take over its predecessor's entry (p >= 1, p != BEST, p != head), bit 3j+1 = slot p has a
  // predecessor to compare with (p != 0, p != head), bit 3j+2 = p == BEST
  uint32_t lane_flags;
  uint32_t static_flags;  // the same bits as if no slot of this lane were the head (they depend on BEST only)
  static constexpr uint32_t SORTED = 32u * NS;

  // the head slot neither receives nor compares with its predecessor: clear bits 3j and 3j+1 of the register holding it
  __device__ __forceinline__ void update_flags()
  {
    const uint32_t lane = lane_id();
    uint32_t f = static_flags;
#pragma unroll
    for (int j = 0; j < NS; ++j)
      if (32u * j + lane == head) f &= ~(3u << (3 * j));
    lane_flags = f;
  }

  __device__ __forceinline__ void init(uint32_t best)
  {
    BEST = best;
    head = best;
    uint32_t f = 0;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      key[j] = EMPTY_KEY;
      dist[j] = G200_INF;
      const uint32_t p = 32u * j + lane_id();
      f |= ((p >= 1) && (p != best)) ? (1u << (3 * j)) : 0u;
      f |= (p != 0) ? (2u << (3 * j)) : 0u;
      f |= (p == best) ? (4u << (3 * j)) : 0u;
    }
    static_flags = f;
    update_flags();
  }
  // common interface with SmemLists (the lists live in registers: no backing store needed)
  __device__ __forceinline__ void init(uint32_t best, void*, uint32_t) { init(best); }

  // simple_knn_cache.cuh:335-352 write_best: slots [0, K) -> ids (+ id_off) and distances
  __device__ __forceinline__ void write_results(int* __restrict__ ids, float* __restrict__ dists, uint32_t K, int id_off) const
  {
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const uint32_t k = 32u * j + lane_id();
      if (k < K) {
        ids[k] = key[j] + id_off;
        if (dists) dists[k] = dist[j];
      }
    }
  }

  __device__ __forceinline__ float dist_at(uint32_t p) const
  {
    float v = dist[0];
#pragma unroll
    for (int j = 1; j < NS; ++j) v = (p >> 5) == j ? dist[j] : v;
    return __shfl_sync(FULL, v, p & 31);
  }
  __device__ __forceinline__ int key_at(uint32_t p) const
  {
    int v = key[0];
#pragma unroll
    for (int j = 1; j < NS; ++j) v = (p >> 5) == j ? key[j] : v;
    return __shfl_sync(FULL, v, p & 31);
  }

  // simple_knn_cache.cuh:126-213 -- all slots updated at once (see DESIGN.md for the equivalence
  // with the reference's block-by-block loop)
  // check_dup (uniform): false when the caller knows that k is not in the lists (a candidate that passed the fetch
  // filter and is no duplicate within its fetch: nothing but other candidates of that fetch has been pushed since)
  __device__ __forceinline__ void push(int k, float d, bool check_dup = true)
  {
    const int lane = lane_id();
    if (check_dup) {
      bool dup = false;
#pragma unroll
      for (int j = 0; j < NS; ++j) dup |= (key[j] == k);
      if (__any_sync(FULL, dup)) return;  // :132-146
    }

    int nk[NS];
    float asd[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      int pk = __shfl_up_sync(FULL, key[j], 1);
      float pd = __shfl_up_sync(FULL, dist[j], 1);
      if (j > 0) {
        const int ck = __shfl_sync(FULL, key[j - 1], 31);
        const float cd = __shfl_sync(FULL, dist[j - 1], 31);
        if (lane == 0) {
          pk = ck;
          pd = cd;
        }
      }
      // slot p receives old[p-1] iff p-1 is active and non-empty and p is neither the start of the
      // best list / prioQ region nor the ring head (:166-172; idx_next==BEST swallows the ring wrap)
#if G200_OPT_FLAGS
      const bool recv = ((lane_flags >> (3 * j)) & 1u) && (pd >= d) && (pk != EMPTY_KEY);
#else
      const uint32_t p = 32u * j + lane;
      const bool recv = (p >= 1) && (p != BEST) && (p != head) && (pd >= d) && (pk != EMPTY_KEY);
#endif
      asd[j] = recv ? pd : dist[j];
      nk[j] = recv ? pk : key[j];
    }
    const float last_asd = __shfl_sync(FULL, asd[NS - 1], 31);
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      float pa = __shfl_up_sync(FULL, asd[j], 1);
      if (j > 0) {
        const float ca = __shfl_sync(FULL, asd[j - 1], 31);
        if (lane == 0) pa = ca;
      }
#if G200_OPT_FLAGS
      if ((lane_flags >> (3 * j)) & 4u) pa = last_asd;  // p == BEST: idx_prev = SORTED-1 (:177)
      const bool active = dist[j] >= d;
      const bool has_prev = (lane_flags >> (3 * j)) & 2u;
#else
      const uint32_t p = 32u * j + lane;
      if (p == BEST) pa = last_asd;
      const bool active = dist[j] >= d;
      const bool has_prev = (p != 0) && (p != head);
#endif
      const bool ins = active && (!has_prev || pa < d);  // :176-182
      key[j] = ins ? k : nk[j];
      dist[j] = ins ? d : asd[j];
    }
  }

  // simple_knn_cache.cuh:215-239 (the visited-ring update is done by the caller)
  __device__ __forceinline__ int pop(float criteria)
  {
    const int k = key_at(head);
    const float dd = dist_at(head);
    if (k == EMPTY_KEY || dd >= criteria) return EMPTY_KEY;
    const int lane = lane_id();
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      if (32u * j + lane == head) {
        key[j] = EMPTY_KEY;
        dist[j] = G200_INF;
      }
    }
    head = (head + 1 >= SORTED) ? BEST : head + 1;
    update_flags();
    return k;
  }

  // mirror the sorted keys to shared memory (for the fetch filter)
  __device__ __forceinline__ void store_keys(int* s_keys) const
  {
#pragma unroll
    for (int j = 0; j < NS; ++j) s_keys[32 * j + lane_id()] = key[j];
  }
  // is `k` in the mirrored sorted part? (per lane, k may differ between lanes)
  __device__ __forceinline__ static bool in_sorted(const int* s_keys, int k)
  {
    bool hit = false;
    const int4* s4 = reinterpret_cast<const int4*>(s_keys);
#pragma unroll
    for (int i = 0; i < NS * 8; ++i) {
      const int4 v = s4[i];
      hit |= (v.x == k) | (v.y == k) | (v.z == k) | (v.w == k);
    }
    return hit;
  }
};

// ------------------------------------------------------------------------------------------------
// The same lists in SHARED memory, for SORTED sizes that do not fit the register file (KQuery up to 6000,
// src/ggnn/query/query_kernels.cu:63-69).  Same slot-for-slot semantics as WarpLists (and therefore as
// SimpleKNNCache::push/pop), processed 32 slots at a time with the previous chunk's boundary values carried
// in registers, so every slot is read (old value) before it is overwritten.
// ------------------------------------------------------------------------------------------------
struct SmemLists {
  int* key;        // [SORTED] shared, 16-byte aligned
  float* dist;     // [SORTED] shared
  uint32_t head;   // physical prioQ head (uniform)
  uint32_t BEST;   // uniform
  uint32_t SORTED; // multiple of 32

  __device__ __forceinline__ void init(uint32_t best, void* smem, uint32_t sorted)
  {
    key = reinterpret_cast<int*>(smem);
    dist = reinterpret_cast<float*>(key + sorted);
    BEST = best;
    head = best;
    SORTED = sorted;
    __syncwarp();
    for (uint32_t p = lane_id(); p < sorted; p += 32) {
      key[p] = EMPTY_KEY;
      dist[p] = G200_INF;
    }
    __syncwarp();
  }
  __device__ __forceinline__ float dist_at(uint32_t p) const { return dist[p]; }
  __device__ __forceinline__ int key_at(uint32_t p) const { return key[p]; }

  __device__ __forceinline__ void push(int k, float d, bool check_dup = true)
  {
    const int lane = lane_id();
    if (check_dup) {
      bool dup = false;
      for (uint32_t p = lane; p < SORTED; p += 32) dup |= (key[p] == k);
      if (__any_sync(FULL, dup)) return;  // :132-146
    }

    // asd[SORTED-1] is what slot BEST compares against (idx_prev = SORTED-1, :177)
    const float od_l1 = dist[SORTED - 1], od_l2 = dist[SORTED - 2];
    const int ok_l2 = key[SORTED - 2];
    const bool recv_last = (SORTED - 1 != BEST) && (SORTED - 1 != head) && (od_l2 >= d) && (ok_l2 != EMPTY_KEY);
    const float last_asd = recv_last ? od_l2 : od_l1;
    __syncwarp();

    int carry_k = EMPTY_KEY;
    float carry_d = 0.f, carry_asd = 0.f;
    for (uint32_t b = 0; b < SORTED; b += 32) {
      const uint32_t p = b + lane;
      const int ok = key[p];
      const float od = dist[p];
      int pk = __shfl_up_sync(FULL, ok, 1);
      float pd = __shfl_up_sync(FULL, od, 1);
      if (lane == 0) {
        pk = carry_k;
        pd = carry_d;
      }
      const bool recv = (p >= 1) && (p != BEST) && (p != head) && (pd >= d) && (pk != EMPTY_KEY);
      const float asd = recv ? pd : od;
      const int nk = recv ? pk : ok;
      float pa = __shfl_up_sync(FULL, asd, 1);
      if (lane == 0) pa = carry_asd;
      if (p == BEST) pa = last_asd;
      const bool active = od >= d;
      const bool has_prev = (p != 0) && (p != head);
      const bool ins = active && (!has_prev || pa < d);  // :176-182
      carry_k = __shfl_sync(FULL, ok, 31);
      carry_d = __shfl_sync(FULL, od, 31);
      carry_asd = __shfl_sync(FULL, asd, 31);
      if (ins || recv) {  // every lane writes only its own slot
        key[p] = ins ? k : nk;
        dist[p] = ins ? d : asd;
      }
    }
    __syncwarp();
  }

  __device__ __forceinline__ int pop(float criteria)
  {
    const int k = key[head];
    const float dd = dist[head];
    if (k == EMPTY_KEY || dd >= criteria) return EMPTY_KEY;
    __syncwarp();
    if (lane_id() == 0) {
      key[head] = EMPTY_KEY;
      dist[head] = G200_INF;
    }
    __syncwarp();
    head = (head + 1 >= SORTED) ? BEST : head + 1;
    return k;
  }

  // the keys already live in shared memory: nothing to mirror
  __device__ __forceinline__ void store_keys(int*) const {}
  __device__ __forceinline__ bool in_sorted(const int*, int k) const
  {
    bool hit = false;
    const int4* s4 = reinterpret_cast<const int4*>(key);
    for (uint32_t i = 0; i < SORTED / 4; ++i) {
      const int4 v = s4[i];
      hit |= (v.x == k) | (v.y == k) | (v.z == k) | (v.w == k);
    }
    return hit;
  }

  __device__ __forceinline__ void write_results(int* __restrict__ ids, float* __restrict__ dists, uint32_t K, int id_off) const
  {
    for (uint32_t k = lane_id(); k < K; k += 32) {
      ids[k] = key[k] + id_off;
      if (dists) dists[k] = dist[k];
    }
  }
};

}  // namespace g200
