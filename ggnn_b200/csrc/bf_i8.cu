// bf_i8.cu -- brute-force kNN of uint8 vectors (the reference's BaseT = uint8_t instantiation, include/ggnn/base/lib.h:26-28)
// as an EXACT integer contraction on the int8 tensor cores (tcgen05.mma.kind::i8, u8 x u8 -> s32 in TMEM).
//
// The reference computes ||b - q||^2 on static_cast<float>(value) (include/ggnn/cuda_utils/distance.cuh:104-139); for
// D <= 128 every partial sum is an integer below 2^24, so the fp32 result is the exact integer.  Hence
//   ||b - q||^2 = ||b||^2 - 2 q.b + ||q||^2        (all terms exact in int32)
// needs no error margin and no hi / lo split: ONE u8 MMA pass per tile instead of the three TF32 passes of bf_tc.cu, on
// rows a quarter the size (and the base is never widened to fp32).  Three stages, same structure as bf_tc.cu:
//   1. pack   : base and queries -> tile-major, pre-swizzled 16 KB records (128 rows x 128 bytes, rows shorter than 128
//               bytes zero padded): the bytes are exactly the SWIZZLE_128B shared-memory image of a K-major UMMA operand
//               tile (16-byte chunk c of row r stored at chunk c ^ (r & 7)); integer row norms.
//   2. gemm   : one CTA per (128-query tile, base split).  The query tile (A) is loaded once, the base tiles (B) stream
//               through an mbarrier ring, one linear 16 KB bulk copy per tile; one thread issues D/32 UTCIMMA per tile
//               into a double-buffered 128 x 128 s32 accumulator in TMEM; eight epilogue warps (one query row and one
//               column half per thread) compare  ||b||^2 - 2 acc  with the row's running K-th best (max-heap in shared
//               memory, bound shared between the splits of a query) and append every row that can still be among the K
//               best -- ties included -- to the query's candidate list.
//   3. rerank : one warp per query recomputes its candidates' integer distances from the uint8 rows (VABSDIFF4 + IDP4A +
//               REDUX) and selects the K smallest (distance, index) pairs: identical ids and distances to
//               src/ggnn/query/bf_query_layer.cu:39-65 on a uint8 base.
// A query whose candidate list overflows is re-done by an exact scan, so the result never depends on the capacity.
// Covered: Euclidean, D in {32, 64, 96, 128}, KQuery <= 128; everything else returns GGNN_B200_ERR_UNSUPPORTED (the
// caller widens the rows and uses ggnn_b200_bf_query: identical results).
#include "umma.cuh"
#include "bf_tc_host.h"
#include "host_util.h"
#include "../../include/ggnn_b200.h"

#include <algorithm>
#include <cstdio>
#include <vector>

namespace g200 {

constexpr int I8_BM = 128;                // queries per CTA (UMMA M)
constexpr int I8_BN = 128;                // base rows per tile (UMMA N)
constexpr uint32_t I8_ROW_BYTES = 128;    // one 128-byte swizzle row per vector (D <= 128, zero padded)
constexpr uint32_t I8_TILE_BYTES = I8_BN * I8_ROW_BYTES;  // 16 KB
constexpr int I8_THREADS = 384;           // warps 0 producer, 1 MMA issuer, 2 TMEM allocator, 3 merger, 4..11 epilogue
constexpr int I8_INF = 0x7f7f7f7f;        // "no bound yet" (what cudaMemset(0x7f) writes); every real score is < 2^25
constexpr int I8_KP_MAX = 128;
constexpr int I8_BNORM_INTS = 8 * 2 * 64;  // base norms of a tile, staged per epilogue warp (its 64 columns, two tiles)

// ---- stage 1: pack ------------------------------------------------------------------------------
// one thread per 16-byte chunk, 8 chunks per output row; rows >= n_rows (up to n_rows_pad) and chunks past D are zero
__global__ void __launch_bounds__(256) i8_pack_kernel(const uint8_t* __restrict__ x, uint32_t n_rows, uint32_t n_rows_pad,
                                                      uint32_t D, uint8_t* __restrict__ tiled, int32_t* __restrict__ norms)
{
  const uint32_t gid = blockIdx.x * 256u + threadIdx.x;
  const uint32_t row = gid >> 3, c = gid & 7u;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (row < n_rows && c * 16u < D) v = __ldg(reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * D + c * 16u));
  unsigned n = __dp4a(v.x, v.x, 0u);
  n = __dp4a(v.y, v.y, n);
  n = __dp4a(v.z, v.z, n);
  n = __dp4a(v.w, v.w, n);
  n += __shfl_xor_sync(FULL, n, 1);
  n += __shfl_xor_sync(FULL, n, 2);
  n += __shfl_xor_sync(FULL, n, 4);
  if (row < n_rows_pad) {
    const uint32_t t = row >> 7, r = row & 127u;
    *reinterpret_cast<uint4*>(tiled + static_cast<size_t>(t) * I8_TILE_BYTES + r * I8_ROW_BYTES + ((c ^ (r & 7u)) << 4)) = v;
    if (c == 0 && row < n_rows) norms[row] = static_cast<int32_t>(n);
  }
}

// ---- stage 2: the contraction -------------------------------------------------------------------
struct I8GemmArgs {
  uint32_t N_base, N_query, K, cap;
  uint32_t rows_per_split;  // multiple of I8_BN
  uint32_t ksteps;          // D / 32: UMMA_K = 32 bytes
  uint32_t rotate;          // CTA x starts its walk over the split's tiles at tile (x * rotate) % n_tiles (0: all at tile 0)
  unsigned int* pub;        // [N_query][2 * splits][K] published best lists (merger warp), or nullptr
  const int32_t* bnorm;     // [N_base]
  const int32_t* qnorm;     // [N_query]
  const uint8_t* q_tiled;   // packed query tiles
  const uint8_t* b_tiled;   // packed base tiles
  int32_t* cand;            // [N_query, cap]
  uint32_t* cnt;            // [N_query]
  unsigned int* tau_g;      // [N_query] best known upper bound of each query's K-th best distance, shared by all splits
};

// instruction descriptor of tcgen05.mma.kind::i8: D s32 (2 at bit 4), A / B unsigned 8 bit (0 at bits 7 and 10), both
// K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t I8_IDESC = (2u << 4) | (0u << 7) | (0u << 10) | ((I8_BN >> 3) << 17) | ((I8_BM >> 4) << 24);

// KP = capacity of the per-row best lists (K <= KP); NSTAGE = depth of the B ring
template <int KP, int NSTAGE>
__global__ void __launch_bounds__(I8_THREADS, 1) i8_gemm_kernel(const I8GemmArgs a)
{
  extern __shared__ unsigned char smem_unaligned[];
  // carve-up (operand tiles 1024-byte aligned: required by the 128-byte swizzle)
  unsigned char* smem = smem_unaligned + ((1024u - (smem_u32(smem_unaligned) & 1023u)) & 1023u);
  unsigned char* sA = smem;                                                  // 16 KB query tile
  unsigned char* sB = smem + I8_TILE_BYTES;                                  // [NSTAGE] 16 KB base tiles
  int* s_kbest = reinterpret_cast<int*>(sB + NSTAGE * I8_TILE_BYTES);        // [2 column halves][128][KP]
  int* s_bnorm = s_kbest + 2 * I8_BM * KP;                                   // [8 epilogue warps][2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bnorm + I8_BNORM_INTS);
  uint64_t* full = bars;                    // [NSTAGE]
  uint64_t* empty = bars + NSTAGE;          // [NSTAGE]
  uint64_t* a_full = bars + 2 * NSTAGE;     // [1]
  uint64_t* t_full = a_full + 1;            // [2]
  uint64_t* t_empty = t_full + 2;           // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(t_empty + 2);
  volatile uint32_t* s_done = s_tmem + 1;   // epilogue warps that have finished

  const int warp = threadIdx.x >> 5;
  const int lane = lane_id();
  const uint32_t q0 = blockIdx.x * I8_BM;
  const uint32_t n_begin = blockIdx.y * a.rows_per_split;
  const uint32_t n_end = min(a.N_base, n_begin + a.rows_per_split);
  const uint32_t n_tiles = (n_end > n_begin) ? (n_end - n_begin + I8_BN - 1) / I8_BN : 0;
  // The CTAs of a split all stream the same base tiles.  Started together they would ask for the same 16 KB record at
  // the same time, over and over, and queue up on the few L2 slices that hold it; so every CTA walks the split's tiles
  // in the same cyclic order but from its own starting tile.
  const uint32_t rot = n_tiles > 1 ? (blockIdx.x * a.rotate) % n_tiles : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    *s_done = 0;
    mbar_init(a_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 8);
    }
    mbar_fence_init();
  }
  constexpr uint32_t TMEM_COLS = 256u;  // 2 accumulators x 128 columns of s32
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===== producer: the query tile once, then one linear 16 KB bulk copy (TMA engine) per base tile =====
    // (a split past the end of the base has no tiles: nothing may be left in flight when the CTA exits)
    if (lane == 0 && n_tiles > 0) {
      mbar_expect_tx(a_full, I8_TILE_BYTES);
      bulk_g2s(sA, a.q_tiled + static_cast<size_t>(blockIdx.x) * I8_TILE_BYTES, I8_TILE_BYTES, a_full);
      uint32_t stage = 0, phase = 0;
      uint32_t ti = rot;
      for (uint32_t t = 0; t < n_tiles; ++t) {
        const size_t tile = (n_begin / I8_BN) + ti;
        if (++ti == n_tiles) ti = 0;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], I8_TILE_BYTES);
        bulk_g2s(sB + stage * I8_TILE_BYTES, a.b_tiled + tile * I8_TILE_BYTES, I8_TILE_BYTES, &full[stage]);
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  }
  else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0 && n_tiles > 0) {
      mbar_wait(a_full, 0);
      tc_fence_after();
      const uint64_t da = umma_desc_sw128(sA);
      uint32_t stage = 0, phase = 0;
      for (uint32_t t = 0; t < n_tiles; ++t) {
        const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
        mbar_wait(&t_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + acc * I8_BN;
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t db = umma_desc_sw128(sB + stage * I8_TILE_BYTES);
        for (uint32_t k = 0; k < a.ksteps; ++k) {  // +2 descriptor units (32 bytes) per UMMA_K inside the swizzle row
          const uint64_t ko = static_cast<uint64_t>(2 * k);
          umma_i8(tmem_c, da + ko, db + ko, I8_IDESC, k != 0);
        }
        umma_commit(&empty[stage]);  // frees this B stage once the MMAs above have read it
        umma_commit(&t_full[acc]);   // accumulator complete
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  }
  else if (warp == 3) {
    // ===== merger: tightens the shared bound tau_g of this CTA's 128 queries while the tiles stream =====
    // (same reasoning as bf_tc.cu: the K-th smallest entry of the union of all published lists -- distinct rows --
    // bounds the K-th best of everything seen so far; lists only decrease entry-wise, so a torn snapshot stays valid)
    if (a.pub != nullptr) {
      const uint32_t n_lists = gridDim.y * 2, K = a.K;
      const uint32_t M = min(n_lists * K, 512u);  // at most 16 values per lane (a prefix of the lists is still valid)
      bool run = true;
      while (run) {
#pragma unroll 1
        for (uint32_t r = 0; r < I8_BM; ++r) {
          const uint32_t q = q0 + r;
          if (*s_done >= 8u || q >= a.N_query) {
            run = *s_done < 8u;
            break;
          }
          const unsigned int* src = a.pub + static_cast<size_t>(q) * n_lists * K;
          uint32_t v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = lane + 32 * j < M ? __ldcg(src + lane + 32 * j) : static_cast<uint32_t>(I8_INF);
          uint32_t x = 0;  // smallest x with count(v <= x) >= K, from the top bit down; the low 4 bits stay at 1
#pragma unroll 1
          for (int b = 30; b >= 4; --b) {
            const uint32_t y = x | ((1u << b) - 1u);
            int c = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) c += v[j] <= y;
            c = __reduce_add_sync(FULL, c);
            if (static_cast<uint32_t>(c) < K) x |= 1u << b;
          }
          x |= 15u;
          if (lane == 0 && x < __ldcg(&a.tau_g[q])) atomicMin(&a.tau_g[q], x);
          __nanosleep(500);
        }
      }
    }
  }
  else if (warp >= 4) {
    // ===== epilogue: one query row and one half of the tile's columns per thread =====
    const int ew = warp & 3;                  // TMEM lane quarter == warp % 4
    const int ch = (warp - 4) >> 2;           // column half
    const uint32_t r = ew * 32 + lane;        // row in the tile
    const uint32_t q = q0 + r;
    const bool live = q < a.N_query;
    const uint32_t K = a.K;
    int* kb = s_kbest + (ch * I8_BM + r) * KP;
    for (uint32_t i = 0; i < KP; ++i) kb[i] = I8_INF;
    const int qn = live ? a.qnorm[q] : 0;
    // tau: upper bound of the K-th best distance (= |b|^2 - 2 q.b + |q|^2, exact): the K-th best of the rows this thread
    // has seen, tightened by what the other lists of the same query have published in tau_g
    int tau = I8_INF;
    uint32_t c_pos = 0, c_left = 0;  // this thread's current chunk of candidate slots
    // the per-tile global loads (base norms, shared bound) are issued one tile ahead: their latency stays off the tile
    // loop's critical path (a bound that is one tile old is still a bound)
    uint32_t ti = rot;  // position of the current tile in the split (same cyclic order as the producer's)
    // The norms of the tile's base rows are staged PER WARP (its 64 columns: lane l loads columns l and 32 + l), so the
    // eight epilogue warps never wait for each other: a warp that has candidates to file falls behind by up to the two
    // accumulator buffers and catches up again, instead of stalling the other seven at a CTA-wide barrier every tile.
    int* s_bn_w = s_bnorm + (warp - 4) * 128;
    auto load_bn = [&](uint32_t row) { return row < n_end ? a.bnorm[row] : 0x7fffffff; };
    const uint32_t c0 = 64u * ch + lane;  // this lane's first column
    int bn_next0 = n_tiles > 0 ? load_bn(n_begin + ti * I8_BN + c0) : 0;
    int bn_next1 = n_tiles > 0 ? load_bn(n_begin + ti * I8_BN + c0 + 32u) : 0;
    int tg_next = (live && n_tiles > 0) ? static_cast<int>(__ldcg(&a.tau_g[q])) : I8_INF;
    for (uint32_t t = 0; t < n_tiles; ++t) {
      const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
      const uint32_t n0 = n_begin + ti * I8_BN;
      if (++ti == n_tiles) ti = 0;
      // (rows past the end: zero rows of the packed operand, rejected again in pass 2)
      int* bnw = s_bn_w + acc * 64;
      bnw[lane] = bn_next0;
      bnw[32 + lane] = bn_next1;
      tau = min(tau, tg_next);
      if (t + 1 < n_tiles) {
        const uint32_t n1 = n_begin + ti * I8_BN;  // first row of the next tile
        bn_next0 = load_bn(n1 + c0);
        bn_next1 = load_bn(n1 + c0 + 32u);
        if (live) tg_next = static_cast<int>(__ldcg(&a.tau_g[q]));
      }
      __syncwarp();
      mbar_wait(&t_full[acc], acc_phase);
      tc_fence_after();
      const int4* bn4 = reinterpret_cast<const int4*>(bnw);  // [16 groups of 4 columns]
      bool improved = false;
      // pass 1, branch-free: which groups of 4 columns hold a row with |b|^2 - 2 acc <= tau - |q|^2 ?
      const int thr = tau - qn;
      uint32_t m = 0;
      {
        int v[64];
        tmem_ld64i(tmem_base + acc * I8_BN + 64u * ch + ((ew * 32u) << 16), v);
#pragma unroll
        for (int j4 = 0; j4 < 16; ++j4) {
          const int4 bn = bn4[j4];
          const int s0 = bn.x - 2 * v[4 * j4 + 0], s1 = bn.y - 2 * v[4 * j4 + 1];
          const int s2 = bn.z - 2 * v[4 * j4 + 2], s3 = bn.w - 2 * v[4 * j4 + 3];
          m |= (min(min(s0, s1), min(s2, s3)) <= thr ? 1u : 0u) << j4;
        }
      }
      if (!live) m = 0;  // rows past the last query
      // pass 2, rare once tau is tight: the warp walks the union of the lanes' group masks; the 4 columns of a group
      // are re-read from tensor memory (warp-wide, 4 registers)
      for (uint32_t um = __reduce_or_sync(FULL, m); um; um &= um - 1) {
        const int g = __ffs(um) - 1;                   // group within this warp's column half
        const int col = 2 * ch * 32 + 4 * g;           // first of its 4 columns in the tile
        int w[4];
        tmem_ld4i(tmem_base + acc * I8_BN + col + ((ew * 32u) << 16), w);
        if (!((m >> g) & 1u)) continue;
        const int4 bn = bn4[g];
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
          const int acc_u = u == 0 ? w[0] : (u == 1 ? w[1] : (u == 2 ? w[2] : w[3]));
          const int bn_u = u == 0 ? bn.x : (u == 1 ? bn.y : (u == 2 ? bn.z : bn.w));
          if (n0 + col + u >= n_end) continue;
          const int s = bn_u - 2 * acc_u + qn;  // exact squared distance
          if (s <= tau) {                       // ties with the K-th best stay candidates: the final order is (dist, id)
            // candidate slots are handed out in chunks of 8 per (query, list): one returning atomic per chunk
            if (c_left == 0) {
              c_pos = atomicAdd(&a.cnt[q], 8u);
              c_left = 8u;
            }
            if (c_pos < a.cap) a.cand[static_cast<size_t>(q) * a.cap + c_pos] = static_cast<int32_t>(n0 + col + u);
            ++c_pos;
            --c_left;
            if (s < tau) {
              // the row's K best distances so far form a MAX-HEAP in kb[0..K): replace its root by s and sift down
              uint32_t i = 0;
              while (true) {
                uint32_t c = 2 * i + 1;
                if (c >= K) break;
                if (c + 1 < K && kb[c + 1] > kb[c]) ++c;
                if (kb[c] <= s) break;
                kb[i] = kb[c];
                i = c;
              }
              kb[i] = s;
              if (kb[0] < tau) {
                tau = kb[0];
                improved = true;
              }
            }
          }
        }
      }
      if (improved) {
        atomicMin(&a.tau_g[q], static_cast<unsigned int>(tau));
        if (a.pub != nullptr) {
          unsigned int* dst = a.pub + (static_cast<size_t>(q) * (gridDim.y * 2) + blockIdx.y * 2 + ch) * K;
          for (uint32_t i = 0; i < K; ++i) dst[i] = static_cast<unsigned int>(kb[i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
    }
    if (lane == 0) atomicAdd(const_cast<uint32_t*>(s_done), 1u);
    // unused slots of the last chunk hold no candidate
    for (; c_left > 0; --c_left, ++c_pos)
      if (c_pos < a.cap) a.cand[static_cast<size_t>(q) * a.cap + c_pos] = -1;
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- stage 3: exact re-rank on the uint8 rows ------------------------------------------------------
struct I8RerankArgs {
  ggnn_b200_bf_query_params p;  // d_base / d_query: uint8 rows of D bytes
  uint32_t N_query, cap;
  const int32_t* cand;
  const uint32_t* cnt;
};

__device__ __forceinline__ void i8_write_results(const ggnn_b200_bf_query_params& p, uint32_t n, uint32_t K, int lane, int j,
                                                 int id, float dist)
{
  const uint32_t k = 32u * j + lane;
  if (k < K) {
    p.d_query_results[static_cast<size_t>(n) * K + k] = id == 0x7fffffff ? EMPTY_KEY : id;
    if (p.d_query_results_dists) p.d_query_results_dists[static_cast<size_t>(n) * K + k] = dist;
  }
}

template <int NSK>
__global__ void __launch_bounds__(128) i8_rerank_kernel(const I8RerankArgs a)
{
  const int lane = lane_id();
  const uint32_t n = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (n >= a.N_query) return;
  const uint32_t cnt = a.cnt[n];
  if (cnt > a.cap) return;  // overflow: handled by the exact scan (i8_fallback_kernel)
  const ggnn_b200_bf_query_params& p = a.p;
  const uint32_t W = p.D / 4;  // 32-bit words per row (<= 32): lane t owns word t
  const uint8_t* base = reinterpret_cast<const uint8_t*>(p.d_base);
  const uint32_t qw = static_cast<uint32_t>(lane) < W
                          ? reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(p.d_query) + static_cast<size_t>(n) * p.D)[lane]
                          : 0u;
  LexKBest<NSK> best;
  best.init();
  const uint32_t K = p.KQuery;
  for (uint32_t c0 = 0; c0 < cnt; c0 += 32) {
    const int nb = min(32u, cnt - c0);
    const int cid = (lane < nb) ? a.cand[static_cast<size_t>(n) * a.cap + c0 + lane] : -1;  // -1: unused slot of a chunk
    const int id = max(cid, 0);
    unsigned mine = 0u;
    for (int i = 0; i < nb; i += 4) {  // four row reads in flight
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rid = __shfl_sync(FULL, id, (i + j) & 31);
        w[j] = static_cast<uint32_t>(lane) < W ? __ldg(reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(rid) * p.D) + lane) : qw;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t s = __vabsdiffu4(w[j], qw);
        const unsigned tot = __reduce_add_sync(FULL, __dp4a(s, s, 0u));
        if (lane == i + j) mine = tot;
      }
    }
    const float d = static_cast<float>(mine);  // exact: < 2^24
    unsigned rem = __ballot_sync(FULL, cid >= 0);
    while (true) {
      float wd;
      int wi;
      best.worst(K, wd, wi);
      const unsigned pm = __ballot_sync(FULL, LexKBest<NSK>::less(d, id, wd, wi)) & rem;
      if (!pm) break;
      const int r = __ffs(pm) - 1;
      best.add(__shfl_sync(FULL, d, r), __shfl_sync(FULL, id, r));
      rem &= ~((2u << r) - 1u);
    }
  }
#pragma unroll
  for (int j = 0; j < NSK; ++j) i8_write_results(p, n, K, lane, j, best.id[j], best.dist[j]);
}

// exact scan for queries whose candidate list overflowed: lane l takes row i0 + l
template <int NSK>
__global__ void __launch_bounds__(128) i8_fallback_kernel(const I8RerankArgs a)
{
  const int lane = lane_id();
  const uint32_t n = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (n >= a.N_query) return;
  if (a.cnt[n] <= a.cap) return;
  const ggnn_b200_bf_query_params& p = a.p;
  const uint32_t W = p.D / 4;
  const uint8_t* base = reinterpret_cast<const uint8_t*>(p.d_base);
  const uint32_t qw = static_cast<uint32_t>(lane) < W
                          ? reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(p.d_query) + static_cast<size_t>(n) * p.D)[lane]
                          : 0u;
  LexKBest<NSK> best;
  best.init();
  const uint32_t K = p.KQuery;
  const uint32_t N = static_cast<uint32_t>(p.N_base);
  for (uint32_t i0 = 0; i0 < N; i0 += 32) {
    const uint32_t row = i0 + lane;
    const bool valid = row < N;
    const uint32_t* rp = reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(valid ? row : 0u) * p.D);
    unsigned acc = 0u;
    for (uint32_t w = 0; w < W; ++w) {
      const uint32_t qv = __shfl_sync(FULL, qw, w);
      const uint32_t s = __vabsdiffu4(__ldg(rp + w), qv);
      acc = __dp4a(s, s, acc);
    }
    const float d = static_cast<float>(acc);
    const int id = static_cast<int>(row);
    unsigned rem = __ballot_sync(FULL, valid);
    while (true) {
      float wd;
      int wi;
      best.worst(K, wd, wi);
      const unsigned pm = __ballot_sync(FULL, LexKBest<NSK>::less(d, id, wd, wi)) & rem;
      if (!pm) break;
      const int r = __ffs(pm) - 1;
      best.add(__shfl_sync(FULL, d, r), __shfl_sync(FULL, id, r));
      rem &= ~((2u << r) - 1u);
    }
  }
#pragma unroll
  for (int j = 0; j < NSK; ++j) i8_write_results(p, n, K, lane, j, best.id[j], best.dist[j]);
}

// ---- diagnostics: ONE 128 x 128 x (32 * ksteps) product of two packed tiles --------------------------
__global__ void __launch_bounds__(128, 1) i8_mma_probe_kernel(const uint8_t* __restrict__ a_tile, const uint8_t* __restrict__ b_tile,
                                                              uint32_t ksteps, int32_t* __restrict__ out)
{
  extern __shared__ unsigned char smem_unaligned[];
  unsigned char* smem = smem_unaligned + ((1024u - (smem_u32(smem_unaligned) & 1023u)) & 1023u);
  unsigned char* sA = smem;
  unsigned char* sB = smem + I8_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + I8_TILE_BYTES);  // [0] operands landed, [1] accumulator complete
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5;
  const int lane = lane_id();
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], 2 * I8_TILE_BYTES);
    bulk_g2s(sA, a_tile, I8_TILE_BYTES, &bars[0]);
    bulk_g2s(sB, b_tile, I8_TILE_BYTES, &bars[0]);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint64_t da = umma_desc_sw128(sA), db = umma_desc_sw128(sB);
    for (uint32_t k = 0; k < ksteps; ++k) umma_i8(tmem_base, da + 2 * k, db + 2 * k, I8_IDESC, k != 0);
    umma_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  for (int c = 0; c < 4; ++c) {
    int v[32];
    tmem_ld32i(tmem_base + c * 32 + ((warp * 32u) << 16), v);
#pragma unroll
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * I8_BN + c * 32 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
}

// ---- host ---------------------------------------------------------------------------------------
struct I8Workspace {
  uint8_t *b_tiled, *q_tiled;
  int32_t *bnorm, *qnorm;
  uint32_t* cnt;
  unsigned int* tau_g;
  unsigned int* pub;
  int32_t* cand;
  size_t total;
};
static I8Workspace i8_layout(void* basep, uint32_t N, uint32_t Nq, uint32_t cap, uint32_t K)
{
  char* b = static_cast<char*>(basep);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = b + off;
    off += (bytes + 1023) / 1024 * 1024;
    return p;
  };
  I8Workspace w;
  const size_t N_pad = (static_cast<size_t>(N) + I8_BN - 1) / I8_BN * I8_BN;
  const size_t Nq_pad = (static_cast<size_t>(Nq) + I8_BM - 1) / I8_BM * I8_BM;
  w.b_tiled = reinterpret_cast<uint8_t*>(take(N_pad * I8_ROW_BYTES));
  w.q_tiled = reinterpret_cast<uint8_t*>(take(Nq_pad * I8_ROW_BYTES));
  w.bnorm = reinterpret_cast<int32_t*>(take(static_cast<size_t>(N) * 4));
  w.qnorm = reinterpret_cast<int32_t*>(take(static_cast<size_t>(Nq) * 4));
  w.cnt = reinterpret_cast<uint32_t*>(take(static_cast<size_t>(Nq) * 4));
  w.tau_g = reinterpret_cast<unsigned int*>(take(static_cast<size_t>(Nq) * 4));
  w.pub = reinterpret_cast<unsigned int*>(take(static_cast<size_t>(Nq) * 2 * TC_MAX_SPLITS * K * 4));
  w.cand = reinterpret_cast<int32_t*>(take(static_cast<size_t>(Nq) * cap * 4));
  w.total = off;
  return w;
}

static bool i8_supported(uint32_t D, uint32_t K, int measure)
{
  return measure == GGNN_B200_EUCLIDEAN && D % 32 == 0 && D >= 32 && D <= 128 && K >= 1 && K <= I8_KP_MAX;
}

static int i8_pack(const uint8_t* x, uint32_t n_rows, uint32_t n_rows_pad, uint32_t D, uint8_t* tiled, int32_t* norms, cudaStream_t stream)
{
  const uint64_t threads = static_cast<uint64_t>(n_rows_pad) * 8;
  i8_pack_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(x, n_rows, n_rows_pad, D, tiled, norms);
  return set_cuda_error(cudaGetLastError(), "i8_pack_kernel launch");
}

template <int KP, int NSTAGE>
static int i8_run(const ggnn_b200_bf_query_params& p, uint32_t Nq, const I8Workspace& w, cudaStream_t stream)
{
  const uint32_t N = static_cast<uint32_t>(p.N_base), D = p.D;
  const uint32_t cap = tc_cap(Nq);
  cudaError_t e;
  if ((e = cudaMemsetAsync(w.cnt, 0, static_cast<size_t>(Nq) * 4, stream)) != cudaSuccess) return set_cuda_error(e, "memset cnt");
  if ((e = cudaMemsetAsync(w.tau_g, 0x7f, static_cast<size_t>(Nq) * 4, stream)) != cudaSuccess) return set_cuda_error(e, "memset tau");
  const uint32_t N_pad = (N + I8_BN - 1) / I8_BN * I8_BN;
  const uint32_t Nq_pad = (Nq + I8_BM - 1) / I8_BM * I8_BM;
  if (int rc = i8_pack(reinterpret_cast<const uint8_t*>(p.d_base), N, N_pad, D, w.b_tiled, w.bnorm, stream)) return rc;
  if (int rc = i8_pack(reinterpret_cast<const uint8_t*>(p.d_query), Nq, Nq_pad, D, w.q_tiled, w.qnorm, stream)) return rc;

  const DeviceInfo& dev = device_info();
  const uint32_t q_tiles = Nq_pad / I8_BM;
  const uint32_t n_tiles = N_pad / I8_BN;
  const uint32_t splits = tc_pick_splits(q_tiles, n_tiles, dev.num_sms, cap, p.KQuery);
  const uint32_t tiles_per_split = (n_tiles + splits - 1) / splits;

  I8GemmArgs ga{};
  ga.N_base = N;
  ga.N_query = Nq;
  ga.K = p.KQuery;
  ga.cap = cap;
  ga.rows_per_split = tiles_per_split * I8_BN;
  ga.ksteps = D / 32;
  // starting tiles spread evenly over the split (GGNN_B200_BF_I8_ROTATE=0: every CTA starts at the split's first tile)
  ga.rotate = env_u32("GGNN_B200_BF_I8_ROTATE", 1) ? std::max(1u, tiles_per_split / std::max(1u, q_tiles)) : 0u;
  ga.bnorm = w.bnorm;
  ga.qnorm = w.qnorm;
  ga.q_tiled = w.q_tiled;
  ga.b_tiled = w.b_tiled;
  ga.cand = w.cand;
  ga.cnt = w.cnt;
  ga.tau_g = w.tau_g;
  ga.pub = env_u32("GGNN_B200_BF_MERGER", 1) ? w.pub : nullptr;
  if (ga.pub && (e = cudaMemsetAsync(w.pub, 0x7f, static_cast<size_t>(Nq) * 2 * splits * p.KQuery * 4, stream)) != cudaSuccess)
    return set_cuda_error(e, "memset pub");
  const size_t smem = static_cast<size_t>(1 + NSTAGE) * I8_TILE_BYTES + 2 * I8_BM * KP * 4 + I8_BNORM_INTS * 4 + 256 + 1024;
  auto gemm = i8_gemm_kernel<KP, NSTAGE>;
  if ((e = cudaFuncSetAttribute(gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))) != cudaSuccess)
    return set_cuda_error(e, "cudaFuncSetAttribute(i8_gemm_kernel)");
  gemm<<<dim3(q_tiles, splits), I8_THREADS, smem, stream>>>(ga);
  if ((e = cudaGetLastError()) != cudaSuccess) return set_cuda_error(e, "i8_gemm_kernel launch");

  I8RerankArgs ra{};
  ra.p = p;
  ra.N_query = Nq;
  ra.cap = cap;
  ra.cand = w.cand;
  ra.cnt = w.cnt;
  i8_rerank_kernel<KP / 32><<<(Nq + 3) / 4, 128, 0, stream>>>(ra);
  i8_fallback_kernel<KP / 32><<<(Nq + 3) / 4, 128, 0, stream>>>(ra);
  if (env_u32("GGNN_B200_BF_DEBUG", 0)) {  // diagnostics only: candidate slots handed out per query
    std::vector<uint32_t> h(Nq);
    cudaStreamSynchronize(stream);
    cudaMemcpy(h.data(), w.cnt, static_cast<size_t>(Nq) * 4, cudaMemcpyDeviceToHost);
    uint64_t sum = 0;
    uint32_t mx = 0, over = 0;
    for (uint32_t c : h) {
      sum += c;
      mx = std::max(mx, c);
      over += c > cap;
    }
    fprintf(stderr, "[bf_i8] splits %u q_tiles %u cap %u: candidate slots/query mean %.1f max %u, overflowed %u\n", splits,
            q_tiles, cap, static_cast<double>(sum) / Nq, mx, over);
  }
  return set_cuda_error(cudaGetLastError(), "i8_rerank / i8_fallback launch");
}

}  // namespace g200

using namespace g200;

extern "C" size_t ggnn_b200_bf_query_u8_workspace_bytes(uint32_t D, int32_t measure, uint32_t KQuery, uint32_t N_base,
                                                        uint32_t N_query)
{
  if (!i8_supported(D, KQuery, measure) || N_base < 128 || N_query == 0) return 0;
  return i8_layout(nullptr, N_base, N_query, tc_cap(N_query), KQuery).total;
}

extern "C" int ggnn_b200_bf_query_u8(const ggnn_b200_bf_query_params* pin, uint32_t N_query, ggnn_b200_stream_t stream_)
{
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!pin) return set_error(GGNN_B200_ERR_INVALID, "null params");
  const ggnn_b200_bf_query_params& p = *pin;
  if (!p.d_base || !p.d_query || !p.d_query_results) return set_error(GGNN_B200_ERR_INVALID, "null device pointer");
  if (p.measure != GGNN_B200_EUCLIDEAN && p.measure != GGNN_B200_COSINE)
    return set_error(GGNN_B200_ERR_INVALID, "unknown distance measure");
  if (p.KQuery == 0 || p.KQuery > 6000) return set_error(GGNN_B200_ERR_INVALID, "KQuery must be in [1, 6000]");  // query_kernels.cu:204-218
  if (p.D == 0 || p.D > 4096) return set_error(GGNN_B200_ERR_INVALID, "D must be in [1, 4096]");
  if (p.N_base <= 0) return set_error(GGNN_B200_ERR_INVALID, "N_base must be positive");
  if (!i8_supported(p.D, p.KQuery, p.measure) || p.N_base < 128)
    return set_error(GGNN_B200_ERR_UNSUPPORTED,
                     "native uint8 bf_query: Euclidean, D in {32, 64, 96, 128}, KQuery <= 128, N_base >= 128 (widen to fp32 otherwise)");
  if (N_query == 0) return 0;
  const I8Workspace w = i8_layout(p.d_workspace, static_cast<uint32_t>(p.N_base), N_query, tc_cap(N_query), p.KQuery);
  if (!p.d_workspace || p.workspace_bytes < w.total) return set_error(GGNN_B200_ERR_INVALID, "bf_query_u8 workspace missing or too small");
  // shared memory: 16 KB query tile + ring of 16 KB base tiles + the per-row best lists
  return p.KQuery <= 32 ? i8_run<32, 8>(p, N_query, w, stream) : i8_run<128, 4>(p, N_query, w, stream);
}

extern "C" int ggnn_b200_debug_i8_pack(const uint8_t* d_rows, uint32_t n_rows, uint32_t n_rows_pad, uint32_t D, uint8_t* d_tiled,
                                       int32_t* d_norms, ggnn_b200_stream_t stream)
{
  if (!d_rows || !d_tiled || !d_norms || D % 32 || D < 32 || D > 128 || n_rows_pad % 128 || n_rows > n_rows_pad)
    return set_error(GGNN_B200_ERR_INVALID, "debug_i8_pack: bad arguments");
  return i8_pack(d_rows, n_rows, n_rows_pad, D, d_tiled, d_norms, static_cast<cudaStream_t>(stream));
}

extern "C" int ggnn_b200_debug_i8_mma(const uint8_t* d_a_tile, const uint8_t* d_b_tile, uint32_t ksteps, int32_t* d_out,
                                      ggnn_b200_stream_t stream)
{
  if (!d_a_tile || !d_b_tile || !d_out || ksteps == 0 || ksteps > 4) return set_error(GGNN_B200_ERR_INVALID, "debug_i8_mma: bad arguments");
  const size_t smem = 2 * I8_TILE_BYTES + 64 + 1024;
  cudaError_t e = cudaFuncSetAttribute(i8_mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(i8_mma_probe_kernel)");
  i8_mma_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(d_a_tile, d_b_tile, ksteps, d_out);
  return set_cuda_error(cudaGetLastError(), "i8_mma_probe_kernel launch");
}
