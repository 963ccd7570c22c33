// bf_tc.cu -- brute-force kNN as a dense tensor-core contraction (tcgen05 + TMEM + TMA), exact results.
//
// bf_query really is a contraction: ||b-q||^2 = ||b||^2 + (-2q).b + ||q||^2.  Three stages:
//   1. split   : one streaming pass writes hi/lo TF32 halves of base and (-2)*query (x = hi + lo, hi = x with the
//                13 low mantissa bits cleared) and the row norms.  The base operand is written tile-major and
//                pre-swizzled: one contiguous 32 KB record [hi | lo] per 128-row tile and 32-column k-block whose
//                bytes are the SWIZZLE_128B shared-memory image the UMMA descriptor expects.
//   2. gemm    : warp-specialised CTAs (one per SM), grid = query tiles x base splits.  The 128 query rows (hi, lo)
//                are written ONCE into tensor memory (tcgen05.st) and stay there as the A operand; the B records are
//                streamed by one linear cp.async.bulk per stage through a 6-deep mbarrier ring; one elected thread
//                issues tcgen05.mma.kind::tf32 (A from TMEM, B from shared memory) three times per k-step
//                (hi*hi + hi*lo + lo*hi = "3xTF32", ~2^-21 relative) into a double-buffered 128x128 fp32
//                accumulator in TMEM; eight epilogue warps (two per TMEM lane quarter, one column half each; one
//                query row per thread) read the accumulator with tcgen05.ld, add ||b||^2 and flag, branch-free, the
//                4-column groups holding a score within `margin` of the row's current K-th best; the rare flagged
//                groups are re-read and every such base row is appended to the query's candidate list while a
//                per-row sorted list of the K best approximate scores keeps the bound tight.  The bound is shared
//                between the base splits of a query (global atomicMin) and tightened by a merger warp to the K-th
//                smallest of the union of all published lists.  The margin bounds the approximation error of the
//                score, so the true top-K (in the reference's own fp32 arithmetic) is always a
//                candidate.
//   3. rerank  : one warp per query recomputes its few hundred candidates with the reference's exact fp32
//                summation order (same code as the traversal kernels) and selects the K smallest (distance, index)
//                pairs -- identical ids and distances to src/ggnn/query/bf_query_layer.cu:39-65.
// A query whose candidate list overflows is re-done by the exact SIMT scan, so the result never depends on the
// candidate capacity.
#include "traverse.cuh"
#include "umma.cuh"
#include "bf_tc_host.h"
#include "host_util.h"
#include "../../include/ggnn_b200.h"

#include <algorithm>
#include <cstdio>
#include <vector>

namespace g200 {

constexpr int TC_BM = 128;     // queries per CTA (UMMA M)
constexpr int TC_BN = 128;     // base rows per tile (UMMA N)
constexpr int TC_BK = 32;      // tf32 elements per shared-memory k-block (one 128-byte swizzle row)
constexpr int TC_STAGES = 2;   // B ring depth (each stage = hi + lo k-block = 32 KB)
constexpr int TC_KP = 128;     // max K of the tensor path (per-row best lists in shared memory: 32 or 128 slots)
constexpr int TC_THREADS = 384;  // warps 0 producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, 4..11 epilogue (2 per TMEM lane quarter)
constexpr uint32_t TC_KBLOCK_BYTES = TC_BM * TC_BK * 4;  // 16 KB

// ---- stage 1: split + norms ---------------------------------------------------------------------
// one warp per row; out rows padded with zeros up to n_rows_pad (TMA never reads past them anyway)
// tiled != 0 (base operand): output is tile-major and PRE-SWIZZLED -- for every 128-row tile t and 32-column k-block
// kb one contiguous 32 KB record [hi 16 KB | lo 16 KB] whose bytes are exactly the SWIZZLE_128B shared-memory image the
// UMMA descriptor expects (16-byte chunk c of row r stored at chunk c ^ (r & 7)), so the GEMM streams B with ONE
// linear 32 KB bulk copy per stage instead of 2 x 128 strided 128-byte row segments.  Rows past n_rows are zero.
constexpr int TC_SPLIT_ROWS = 32;  // rows per CTA of tc_split_kernel (8 warps x 4 rows)
// normalize != 0 (cosine): every row is scaled to unit length first (zero rows stay zero) and the stored norm is
// `unit_norm` (0 for the base, 1 for the queries): with -q/|q| as the A operand the score acc + 0 + 1 is 1 - cos(q, b),
// the reference's cosine distance (include/ggnn/cuda_utils/distance.cuh:140-159) up to rounding far below the margin.
__global__ void __launch_bounds__(256) tc_split_kernel(const float* __restrict__ x, uint32_t n_rows, uint32_t n_rows_out,
                                                       uint32_t D, float scale, int tiled, float* __restrict__ hi,
                                                       float* __restrict__ lo, float* __restrict__ norms,
                                                       unsigned int* __restrict__ max_norm_bits, int normalize, float unit_norm)
{
  __shared__ float s_max[8];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const uint32_t KB = D / 32;
  const uint32_t row0 = blockIdx.x * TC_SPLIT_ROWS + warp * 4;
  const bool act = static_cast<uint32_t>(lane) < D / 4;  // one 16-byte chunk of the row per lane (D <= 128)
  float4 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t row = row0 + i;
    v[i] = (act && row < n_rows) ? __ldg(reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float wmax = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t row = row0 + i;
    float acc = fmaf(v[i].w, v[i].w, fmaf(v[i].z, v[i].z, fmaf(v[i].y, v[i].y, v[i].x * v[i].x)));
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    float row_scale = scale;
    if (normalize) {
      row_scale = acc > 0.f ? scale * rsqrtf(acc) : 0.f;
      acc = unit_norm;
    }
    if (row < n_rows) {
      if (lane == 0) norms[row] = acc;
      wmax = fmaxf(wmax, acc);
    }
    if (row < n_rows_out && act) {
      // (Euclidean: power-of-two scale, exact)
      const float4 sv = make_float4(v[i].x * row_scale, v[i].y * row_scale, v[i].z * row_scale, v[i].w * row_scale);
      float4 h, l;
      h.x = __uint_as_float(__float_as_uint(sv.x) & 0xffffe000u);
      h.y = __uint_as_float(__float_as_uint(sv.y) & 0xffffe000u);
      h.z = __uint_as_float(__float_as_uint(sv.z) & 0xffffe000u);
      h.w = __uint_as_float(__float_as_uint(sv.w) & 0xffffe000u);
      l = make_float4(sv.x - h.x, sv.y - h.y, sv.z - h.z, sv.w - h.w);  // exact
      if (tiled) {
        const uint32_t t = row >> 7, r = row & 127, kb = lane >> 3, c = lane & 7;
        float* rec = hi + (static_cast<size_t>(t) * KB + kb) * (2 * 4096);  // 32 KB record: [hi 16 KB | lo 16 KB]
        const uint32_t off = r * 32 + ((c ^ (r & 7)) << 2);
        *reinterpret_cast<float4*>(rec + off) = h;
        *reinterpret_cast<float4*>(rec + 4096 + off) = l;
      }
      else {
        reinterpret_cast<float4*>(hi + static_cast<size_t>(row) * D)[lane] = h;
        reinterpret_cast<float4*>(lo + static_cast<size_t>(row) * D)[lane] = l;
      }
    }
  }
  if (max_norm_bits) {  // one atomic per CTA (non-negative floats order like their bit patterns)
    if (lane == 0) s_max[warp] = wmax;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = s_max[0];
      for (int i = 1; i < 8; ++i) m = fmaxf(m, s_max[i]);
      atomicMax(max_norm_bits, __float_as_uint(m));
    }
  }
}

// ---- stage 2: the contraction -------------------------------------------------------------------
struct TcGemmArgs {
  uint32_t N_base, N_query, K, cap;
  uint32_t rows_per_split;  // multiple of TC_BN
  unsigned int* pub;        // [N_query][2 * splits][K] published best lists (merger warp), or nullptr
  const float* bnorm;       // [N_base]
  const float* qnorm;       // [N_query]
  const float* q_hi;        // [N_query, D] hi / lo halves of -2*query (row-major)
  const float* q_lo;
  const float* b_tiled;     // base operand, tile-major pre-swizzled records (see tc_split_kernel)
  uint32_t D;
  const unsigned int* max_norm_bits;
  int32_t* cand;            // [N_query, cap]
  uint32_t* cnt;            // [N_query]
  unsigned int* tau_g;      // [N_query] bits of the best known upper bound of each query's K-th best score,
                            // shared by all base splits (atomicMin; scores are shifted to be >= 0)
};

// The 128-row query tile (hi and lo halves of -2q) lives in TENSOR MEMORY (columns 256..), written once per CTA with
// tcgen05.st; all shared memory goes to the B ring.
constexpr int TC_STAGES_TMEM = 6;
constexpr uint32_t TC_CHUNK = 8;

// KB k-blocks (D = 32*KB); KP = capacity of the per-row best lists (K <= KP); NSTAGE = depth of the B ring (the lists
// and the ring share the shared memory: KP 32 -> 6 stages, KP 128 (the API default KGT = 100) -> 3 stages)
template <int KB, int KP, int NSTAGE>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_kernel(const TcGemmArgs a)
{
  extern __shared__ unsigned char smem_unaligned[];
  // carve-up (every operand tile 1024-byte aligned: required by the 128-byte swizzle)
  unsigned char* smem = smem_unaligned + ((1024u - (smem_u32(smem_unaligned) & 1023u)) & 1023u);
  unsigned char* sB = smem;                                                       // [NSTAGE][hi 16 KB | lo 16 KB]
  float* s_kbest = reinterpret_cast<float*>(sB + NSTAGE * 2 * TC_KBLOCK_BYTES);   // [2 column halves][128][KP]
  float* s_bnorm = s_kbest + 2 * TC_BM * KP;                                      // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bnorm + 2 * TC_BN);
  uint64_t* full = bars;                    // [NSTAGE]
  uint64_t* empty = bars + NSTAGE;          // [NSTAGE]
  uint64_t* a_full = bars + 2 * NSTAGE;     // [1]
  uint64_t* t_full = a_full + 1;            // [2]
  uint64_t* t_empty = t_full + 2;           // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(t_empty + 2);
  volatile uint32_t* s_done = s_tmem + 1;  // epilogue warps that have finished

  const int warp = threadIdx.x >> 5;
  const int lane = lane_id();
  const uint32_t q0 = blockIdx.x * TC_BM;
  const uint32_t n_begin = blockIdx.y * a.rows_per_split;
  const uint32_t n_end = min(a.N_base, n_begin + a.rows_per_split);
  const uint32_t n_tiles = (n_end > n_begin) ? (n_end - n_begin + TC_BN - 1) / TC_BN : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    *s_done = 0;
    mbar_init(a_full, 8);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 8);
    }
    mbar_fence_init();
  }
  constexpr uint32_t TMEM_COLS = 512u;  // 2 accumulators x 128 columns + A hi/lo: 2 x 32*KB columns
  constexpr uint32_t COL_A_HI = 256u, COL_A_LO = 256u + 32u * KB;
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===== producer: one linear 32 KB bulk copy (TMA engine) per stage =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (uint32_t t = 0; t < n_tiles; ++t) {
        const size_t tile = (n_begin / TC_BN) + t;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], 2 * TC_KBLOCK_BYTES);
          bulk_g2s(sB + stage * 2 * TC_KBLOCK_BYTES, a.b_tiled + (tile * KB + kb) * (2 * 4096), 2 * TC_KBLOCK_BYTES, &full[stage]);
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  }
  else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      // instruction descriptor: D fp32 (bit 4), A/B tf32 (2 at bits 7 and 10), K-major both, N>>3 at 17, M>>4 at 24
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((TC_BN >> 3) << 17) | ((TC_BM >> 4) << 24);
      mbar_wait(a_full, 0);
      tc_fence_after();
      uint32_t stage = 0, phase = 0;
      for (uint32_t t = 0; t < n_tiles; ++t) {
        const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
        mbar_wait(&t_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + acc * TC_BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t db_hi = umma_desc_sw128(sB + stage * 2 * TC_KBLOCK_BYTES);
          const uint64_t db_lo = umma_desc_sw128(sB + stage * 2 * TC_KBLOCK_BYTES + TC_KBLOCK_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {  // A: +8 TMEM columns per UMMA_K; B: +2 descriptor units (32 bytes)
            const uint64_t ko = static_cast<uint64_t>(2 * k);
            const uint32_t a_hi = tmem_base + COL_A_HI + kb * TC_BK + k * 8;
            const uint32_t a_lo = tmem_base + COL_A_LO + kb * TC_BK + k * 8;
            umma_tf32_ts(tmem_c, a_hi, db_hi + ko, idesc, (kb | k) != 0);
            umma_tf32_ts(tmem_c, a_hi, db_lo + ko, idesc, 1);
            umma_tf32_ts(tmem_c, a_lo, db_hi + ko, idesc, 1);
          }
          umma_commit(&empty[stage]);  // frees this B stage once the MMAs above have read it
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&t_full[acc]);  // accumulator complete
      }
    }
  }
  else if (warp == 3) {
    // ===== merger: tightens the shared bound tau_g of this CTA's 128 queries while the tiles stream =====
    // A list's K-th entry only bounds the K-th best of ITS rows (1/n_lists of the base), so the minimum over lists is
    // about the K*n_lists-th best overall.  The K-th smallest entry of the UNION of the published lists (distinct rows)
    // is the K-th best of everything seen so far.  Lists only ever decrease entry-wise, so a snapshot torn by
    // concurrent updates undercounts and the bound stays valid.  Non-negative floats are compared as bit patterns.
    if (a.pub != nullptr) {
      const uint32_t n_lists = gridDim.y * 2, K = a.K;
      const uint32_t M = min(n_lists * K, 512u);  // at most 16 values per lane (a prefix of the lists is still valid)
      bool run = true;
      while (run) {
#pragma unroll 1
        for (uint32_t r = 0; r < TC_BM; ++r) {
          const uint32_t q = q0 + r;
          if (*s_done >= 8u || q >= a.N_query) {
            run = *s_done < 8u;
            break;
          }
          const unsigned int* src = a.pub + static_cast<size_t>(q) * n_lists * K;
          uint32_t v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = lane + 32 * j < M ? __ldcg(src + lane + 32 * j) : 0x7f7f7f7fu;
          uint32_t x = 0;  // smallest x with count(v <= x) >= K, from the top bit down; the low 12 bits stay at 1
#pragma unroll 1
          for (int b = 30; b >= 12; --b) {
            const uint32_t y = x | ((1u << b) - 1u);
            int c = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) c += v[j] <= y;
            c = __reduce_add_sync(FULL, c);
            if (static_cast<uint32_t>(c) < K) x |= 1u << b;
          }
          x |= 4095u;
          if (lane == 0 && x < __ldcg(&a.tau_g[q])) atomicMin(&a.tau_g[q], x);
          __nanosleep(500);
        }
      }
    }
  }
  else if (warp >= 4) {
    // ===== epilogue: one query row and one half of the tile's columns per thread =====
    // warps 4..7 take columns [0,64), warps 8..11 columns [64,128) of every tile; each keeps its own best list
    // (an extra base split in effect), the shared bound tau_g keeps both tight
    const int ew = warp & 3;                  // TMEM lane quarter == warp % 4
    const int ch = (warp - 4) >> 2;           // column half
    const uint32_t r = ew * 32 + lane;        // row in the tile
    const uint32_t q = q0 + r;
    const bool live = q < a.N_query;
    const uint32_t K = a.K;
    float* kb = s_kbest + (ch * TC_BM + r) * KP;
    for (uint32_t i = 0; i < KP; ++i) kb[i] = G200_INF;
    // this thread's query row (hi, lo halves of -2q) -> tensor memory lane r, one column per K element
    {
      const int half = ch;  // column-half-0 warps write the hi operand, column-half-1 warps the lo operand
      const float* src = (half ? a.q_lo : a.q_hi) + static_cast<size_t>(min(q, a.N_query - 1)) * a.D;
      for (int c = 0; c < KB; ++c) {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 f = live ? reinterpret_cast<const float4*>(src + c * 32)[j] : make_float4(0.f, 0.f, 0.f, 0.f);
          v[4 * j + 0] = __float_as_uint(f.x);
          v[4 * j + 1] = __float_as_uint(f.y);
          v[4 * j + 2] = __float_as_uint(f.z);
          v[4 * j + 3] = __float_as_uint(f.w);
        }
        tmem_st32(tmem_base + (half ? COL_A_LO : COL_A_HI) + c * 32 + ((ew * 32u) << 16), v);
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(a_full);
    // error bound of the approximate score: 2^-13 * (|q|^2 + max |b|^2)
    const float qn = live ? a.qnorm[q] : 0.f;
    const float margin = live ? ldexpf(qn + __uint_as_float(*a.max_norm_bits), -13) : 0.f;
    // tau: upper bound of the K-th best score (= |b|^2 - 2 q.b + |q|^2 >= 0 up to rounding): the K-th best of the rows
    // this CTA has seen, tightened by what the other base splits of the same query have published in tau_g
    float tau = G200_INF;
    uint32_t c_pos = 0, c_left = 0;  // this thread's current chunk of candidate slots
    for (uint32_t t = 0; t < n_tiles; ++t) {
      const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
      const uint32_t n0 = n_begin + t * TC_BN;
      // stage the tile's base norms (rows past the end never qualify)
      if (ch == 0) s_bnorm[acc * TC_BN + r] = (n0 + r < n_end) ? a.bnorm[n0 + r] : G200_INF;
      if (live) tau = fminf(tau, __uint_as_float(__ldcg(&a.tau_g[q])));
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&t_full[acc], acc_phase);
      tc_fence_after();
      const float4* bn4 = reinterpret_cast<const float4*>(s_bnorm + acc * TC_BN);
      bool improved = false;
      // pass 1, branch-free: which groups of 4 columns hold a score below the threshold?  (compare the raw
      // accumulator + |b|^2 against the shifted threshold)
      const float thr = tau + margin - qn;
      uint32_t m = 0;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * ch + cc;
        float v[32];
        tmem_ld32(tmem_base + acc * TC_BN + c * 32 + ((ew * 32u) << 16), v);
        uint32_t mc = 0;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 bn = bn4[c * 8 + j4];
          const float s0 = v[4 * j4 + 0] + bn.x, s1 = v[4 * j4 + 1] + bn.y;
          const float s2 = v[4 * j4 + 2] + bn.z, s3 = v[4 * j4 + 3] + bn.w;
          mc |= (fminf(fminf(s0, s1), fminf(s2, s3)) < thr ? 1u : 0u) << j4;
        }
        m |= mc << (8 * cc);
      }
      if (!live) m = 0;  // rows past the last query
      // pass 2, rare once tau is tight: ONE copy of the candidate / insertion code.  The warp walks the union of the
      // lanes' group masks; the 4 columns of a group are re-read from tensor memory (warp-wide, 4 registers).
      for (uint32_t um = __reduce_or_sync(FULL, m); um; um &= um - 1) {
        const int g = __ffs(um) - 1;                   // group within this warp's column half
        const int col = 2 * ch * 32 + 4 * g;           // first of its 4 columns in the tile
        float w[4];
        tmem_ld4(tmem_base + acc * TC_BN + col + ((ew * 32u) << 16), w);
        if (!((m >> g) & 1u)) continue;
        const float4 bn = bn4[col >> 2];
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
          const float acc_u = u == 0 ? w[0] : (u == 1 ? w[1] : (u == 2 ? w[2] : w[3]));
          const float bn_u = u == 0 ? bn.x : (u == 1 ? bn.y : (u == 2 ? bn.z : bn.w));
          const float s = fmaxf(acc_u + bn_u + qn, 0.f);
          if (s < tau + margin) {
            // candidate slots are handed out in chunks of TC_CHUNK per (query, list): one returning atomic per chunk
            if (c_left == 0) {
              c_pos = atomicAdd(&a.cnt[q], TC_CHUNK);
              c_left = TC_CHUNK;
            }
            if (c_pos < a.cap) a.cand[static_cast<size_t>(q) * a.cap + c_pos] = static_cast<int32_t>(n0 + col + u);
            ++c_pos;
            --c_left;
            if (s < tau) {
              // the row's K best scores so far form a MAX-HEAP in kb[0..K): replace its root (the K-th best, which
              // bounds tau) by s and sift down -- O(log K) instead of the O(K) shifts of a sorted list (K = 100: the
              // whole call went from 91 to 35 ms).  Every slot's value only ever decreases.
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
        atomicMin(&a.tau_g[q], __float_as_uint(tau));  // non-negative floats order like their bit patterns
        if (a.pub != nullptr) {
          unsigned int* dst = a.pub + (static_cast<size_t>(q) * (gridDim.y * 2) + blockIdx.y * 2 + ch) * K;
          for (uint32_t i = 0; i < K; ++i) dst[i] = __float_as_uint(kb[i]);
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

// ---- stage 3: exact re-rank ----------------------------------------------------------------------
struct TcRerankArgs {
  ggnn_b200_bf_query_params p;
  uint32_t N_query, cap, warp_smem_bytes;
  const int32_t* cand;
  const uint32_t* cnt;
};

template <int D32, int NSK>
__global__ void __launch_bounds__(128) tc_rerank_kernel(const TcRerankArgs a)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const uint32_t n = blockIdx.x * 4 + warp;
  if (n >= a.N_query) return;
  const uint32_t cnt = a.cnt[n];
  if (cnt > a.cap) return;  // overflow: handled by the exact scan (tc_fallback_kernel)
  unsigned char* wbase = smem_raw + static_cast<size_t>(warp) * a.warp_smem_bytes;
  WarpSmem ws;
  ws.stage = reinterpret_cast<float*>(wbase);
  ws.s_q = nullptr;
  ws.s_sorted = nullptr;
  ws.bar = reinterpret_cast<uint64_t*>(wbase + 32 * D32 * 32 * 4);
  ws.parity = 0;
  ws.stage_rows = 32;
  ws.stage_mode = 0;
  ws.tmap = nullptr;
  ws.pad_row = 0;
  if (lane == 0) mbar_init(ws.bar, 1);
  mbar_fence_init();
  __syncwarp();
  const ggnn_b200_bf_query_params& p = a.p;
  const DistCfg dc{p.D, 32u, 4u, p.measure};
  QueryVec<true, D32, 1> qv;
  qv.load(dc, p.d_query + static_cast<size_t>(n) * p.D, nullptr);
  LexKBest<NSK> best;
  best.init();
  const uint32_t K = p.KQuery;
  for (uint32_t c0 = 0; c0 < cnt; c0 += 32) {
    const int nb = min(32u, cnt - c0);
    const int cid = (lane < nb) ? a.cand[static_cast<size_t>(n) * a.cap + c0 + lane] : -1;  // -1: unused slot of a chunk
    const int id = max(cid, 0);
    const float d = stage_and_dist<true, D32, 1>(ws, qv, p.d_base, id, 0, nb);
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
  for (int j = 0; j < NSK; ++j) {
    const uint32_t k = 32u * j + lane;
    if (k < K) {
      p.d_query_results[static_cast<size_t>(n) * K + k] = best.id[j] == 0x7fffffff ? EMPTY_KEY : best.id[j];
      if (p.d_query_results_dists) p.d_query_results_dists[static_cast<size_t>(n) * K + k] = best.dist[j];
    }
  }
}

// exact scan for queries whose candidate list overflowed (same arithmetic as bf_query.cu's generic path)
struct KBest1 {
  int id;
  float dist;
};
template <int NSK>
__global__ void __launch_bounds__(128) tc_fallback_kernel(const TcRerankArgs a)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const uint32_t n = blockIdx.x * 4 + warp;
  if (n >= a.N_query) return;
  if (a.cnt[n] <= a.cap) return;
  const ggnn_b200_bf_query_params& p = a.p;
  float* s_q = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(warp) * p.D;
  const DistCfg dc{p.D, 32u, 4u, p.measure};
  QueryVec<false, 1, 1> qv;
  qv.load(dc, p.d_query + static_cast<size_t>(n) * p.D, s_q);
  LexKBest<NSK> best;
  best.init();
  const uint32_t K = p.KQuery;
  for (int i = 0; i < p.N_base; ++i) {
    float x, y;
    dist_partials_generic(dc, p.d_base + static_cast<size_t>(i) * p.D, s_q, x, y);
    if (p.measure != 0) x = cosine_finish(x, y, qv.q_norm);
    float wd;
    int wi;
    best.worst(K, wd, wi);
    if (LexKBest<NSK>::less(x, i, wd, wi)) best.add(x, i);
  }
#pragma unroll
  for (int j = 0; j < NSK; ++j) {
    const uint32_t k = 32u * j + lane;
    if (k < K) {
      p.d_query_results[static_cast<size_t>(n) * K + k] = best.id[j] == 0x7fffffff ? EMPTY_KEY : best.id[j];
      if (p.d_query_results_dists) p.d_query_results_dists[static_cast<size_t>(n) * K + k] = best.dist[j];
    }
  }
}

// ---- host ---------------------------------------------------------------------------------------
struct TcWorkspace {
  float *b_hi, *b_lo, *bnorm, *q_hi, *q_lo, *qnorm;
  unsigned int* max_norm;
  uint32_t* cnt;
  unsigned int* tau_g;
  unsigned int* pub;
  int32_t* cand;
  size_t total;
};
static TcWorkspace tc_layout(void* basep, uint32_t N, uint32_t Nq, uint32_t D, uint32_t cap, uint32_t K)
{
  char* b = static_cast<char*>(basep);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = b + off;
    off += (bytes + 1023) / 1024 * 1024;
    return p;
  };
  TcWorkspace w;
  const size_t N_pad = (static_cast<size_t>(N) + TC_BN - 1) / TC_BN * TC_BN;
  w.b_hi = reinterpret_cast<float*>(take(N_pad * D * 4 * 2));  // tile-major [hi | lo] records
  w.b_lo = nullptr;
  w.bnorm = reinterpret_cast<float*>(take(static_cast<size_t>(N) * 4));
  w.q_hi = reinterpret_cast<float*>(take(static_cast<size_t>(Nq) * D * 4));
  w.q_lo = reinterpret_cast<float*>(take(static_cast<size_t>(Nq) * D * 4));
  w.qnorm = reinterpret_cast<float*>(take(static_cast<size_t>(Nq) * 4));
  w.max_norm = reinterpret_cast<unsigned int*>(take(1024));
  w.cnt = reinterpret_cast<uint32_t*>(take(static_cast<size_t>(Nq) * 4));
  w.tau_g = reinterpret_cast<unsigned int*>(take(static_cast<size_t>(Nq) * 4));
  w.pub = reinterpret_cast<unsigned int*>(take(static_cast<size_t>(Nq) * 2 * TC_MAX_SPLITS * K * 4));
  w.cand = reinterpret_cast<int32_t*>(take(static_cast<size_t>(Nq) * cap * 4));
  w.total = off;
  return w;
}

// Candidate capacity per query.  Every base split of a query emits roughly K*(1 + ln(rows/K)) candidates (its own
// running K-th best starts at +inf), so the capacity bounds the number of splits; overflowing queries are re-done
// exactly, the capacity only affects speed.
constexpr uint32_t TC_CAP_MIN = 1024, TC_CAP_MAX = 8192, TC_CAND_PER_SPLIT = 192;

uint32_t tc_cap(uint32_t Nq)
{
  // keep the candidate buffer below ~256 MB
  const uint64_t by_mem = (256ull << 20) / (static_cast<uint64_t>(std::max(1u, Nq)) * 4);
  return static_cast<uint32_t>(std::max<uint64_t>(TC_CAP_MIN, std::min<uint64_t>(TC_CAP_MAX, by_mem)));
}

// number of base splits: fill the SMs (one 209 KB CTA each) in whole waves, within the candidate capacity
uint32_t tc_pick_splits(uint32_t q_tiles, uint32_t n_tiles, uint32_t num_sms, uint32_t cap, uint32_t K)
{
  const uint32_t forced = env_u32("GGNN_B200_BF_SPLITS", 0);
  const uint32_t per_split = std::max(TC_CAND_PER_SPLIT, 20u * K);  // ~2 lists x K (1 + ln(rows / K)) per split
  const uint32_t s_max = std::max(1u, std::min(std::min(n_tiles, TC_MAX_SPLITS), cap / per_split));
  if (forced) return std::max(1u, std::min(std::min(forced, TC_MAX_SPLITS), n_tiles));
  uint32_t best = 1;
  double best_score = -1.0;
  for (uint32_t s = 1; s <= s_max; ++s) {
    const uint64_t ctas = static_cast<uint64_t>(q_tiles) * s;
    const uint64_t waves = (ctas + num_sms - 1) / num_sms;
    const double eff = static_cast<double>(ctas) / static_cast<double>(waves * num_sms);
    // prefer full waves; among equals, enough CTAs to cover the machine but as few splits as possible
    const double score = eff - 0.002 * s;
    if (score > best_score) {
      best_score = score;
      best = s;
    }
  }
  return best;
}

bool tc_supported(uint32_t D, uint32_t K, int measure)
{
  return (measure == GGNN_B200_EUCLIDEAN || measure == GGNN_B200_COSINE) && D % 32 == 0 && D >= 32 && D <= 128 && K >= 1 &&
         K <= TC_KP;
}

template <int KB, int KP, int NSTAGE>
static int tc_run(const ggnn_b200_bf_query_params& p, uint32_t Nq, const TcWorkspace& w, cudaStream_t stream)
{
  const uint32_t N = static_cast<uint32_t>(p.N_base), D = p.D;
  const uint32_t cap = tc_cap(Nq);
  cudaError_t e;
  if ((e = cudaMemsetAsync(w.max_norm, 0, 4, stream)) != cudaSuccess) return set_cuda_error(e, "memset max_norm");
  if ((e = cudaMemsetAsync(w.cnt, 0, static_cast<size_t>(Nq) * 4, stream)) != cudaSuccess) return set_cuda_error(e, "memset cnt");
  if ((e = cudaMemsetAsync(w.tau_g, 0x7f, static_cast<size_t>(Nq) * 4, stream)) != cudaSuccess) return set_cuda_error(e, "memset tau");  // 0x7f7f7f7f = 3.4e38
  const uint32_t N_pad = (N + TC_BN - 1) / TC_BN * TC_BN;
  const int cosine = p.measure == GGNN_B200_COSINE;
  // Euclidean: |b|^2 + (-2q).b + |q|^2.  Cosine: unit rows, 0 + (-q/|q|).(b/|b|) + 1 = 1 - cos
  tc_split_kernel<<<(N_pad + TC_SPLIT_ROWS - 1) / TC_SPLIT_ROWS, 256, 0, stream>>>(p.d_base, N, N_pad, D, 1.0f, 1, w.b_hi, nullptr, w.bnorm,
                                                                                    w.max_norm, cosine, 0.0f);
  tc_split_kernel<<<(Nq + TC_SPLIT_ROWS - 1) / TC_SPLIT_ROWS, 256, 0, stream>>>(p.d_query, Nq, Nq, D, cosine ? -1.0f : -2.0f, 0, w.q_hi,
                                                                                 w.q_lo, w.qnorm, nullptr, cosine, 1.0f);
  if ((e = cudaGetLastError()) != cudaSuccess) return set_cuda_error(e, "tc_split_kernel launch");

  const DeviceInfo& dev = device_info();
  const uint32_t q_tiles = (Nq + TC_BM - 1) / TC_BM;
  const uint32_t n_tiles = (N + TC_BN - 1) / TC_BN;
  const uint32_t splits = tc_pick_splits(q_tiles, n_tiles, dev.num_sms, cap, p.KQuery);
  const uint32_t tiles_per_split = (n_tiles + splits - 1) / splits;

  TcGemmArgs ga{};
  ga.N_base = N;
  ga.N_query = Nq;
  ga.K = p.KQuery;
  ga.cap = cap;
  ga.rows_per_split = tiles_per_split * TC_BN;
  ga.bnorm = w.bnorm;
  ga.qnorm = w.qnorm;
  ga.max_norm_bits = w.max_norm;
  ga.cand = w.cand;
  ga.cnt = w.cnt;
  ga.tau_g = w.tau_g;
  ga.pub = env_u32("GGNN_B200_BF_MERGER", 1) ? w.pub : nullptr;
  if (ga.pub && (e = cudaMemsetAsync(w.pub, 0x7f, static_cast<size_t>(Nq) * 2 * splits * p.KQuery * 4, stream)) != cudaSuccess)
    return set_cuda_error(e, "memset pub");
  ga.q_hi = w.q_hi;
  ga.q_lo = w.q_lo;
  ga.b_tiled = w.b_hi;
  ga.D = D;
  const size_t smem = static_cast<size_t>(NSTAGE) * 2 * TC_KBLOCK_BYTES + 2 * TC_BM * KP * 4 + 2 * TC_BN * 4 + 256 + 1024;
  auto gemm = tc_gemm_kernel<KB, KP, NSTAGE>;
  if ((e = cudaFuncSetAttribute(gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))) != cudaSuccess)
    return set_cuda_error(e, "cudaFuncSetAttribute(tc_gemm_kernel)");
  gemm<<<dim3(q_tiles, splits), TC_THREADS, smem, stream>>>(ga);
  if ((e = cudaGetLastError()) != cudaSuccess) return set_cuda_error(e, "tc_gemm_kernel launch");

  TcRerankArgs ra{};
  ra.p = p;
  ra.N_query = Nq;
  ra.cap = cap;
  ra.cand = w.cand;
  ra.cnt = w.cnt;
  ra.warp_smem_bytes = align_up(32 * D * 4 + 32, 128);
  const size_t rsmem = static_cast<size_t>(ra.warp_smem_bytes) * 4;
  auto rr = tc_rerank_kernel<KB, KP / 32>;
  if ((e = cudaFuncSetAttribute(rr, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(rsmem))) != cudaSuccess)
    return set_cuda_error(e, "cudaFuncSetAttribute(tc_rerank_kernel)");
  rr<<<(Nq + 3) / 4, 128, rsmem, stream>>>(ra);
  tc_fallback_kernel<KP / 32><<<(Nq + 3) / 4, 128, 4 * D * 4, stream>>>(ra);
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
    fprintf(stderr, "[bf_tc] splits %u q_tiles %u cap %u: candidate slots/query mean %.1f max %u, overflowed %u\n", splits,
            q_tiles, cap, static_cast<double>(sum) / Nq, mx, over);
  }
  return set_cuda_error(cudaGetLastError(), "tc_rerank / tc_fallback launch");
}

int tc_bf_query(const ggnn_b200_bf_query_params& p, uint32_t Nq, void* workspace, size_t workspace_bytes, cudaStream_t stream)
{
  const TcWorkspace w = tc_layout(workspace, static_cast<uint32_t>(p.N_base), Nq, p.D, tc_cap(Nq), p.KQuery);
  if (workspace_bytes < w.total) return set_error(GGNN_B200_ERR_INVALID, "bf_query workspace too small");
  const bool small_k = p.KQuery <= 32;
  switch (p.D / 32) {
    case 1: return small_k ? tc_run<1, 32, 6>(p, Nq, w, stream) : tc_run<1, 128, 3>(p, Nq, w, stream);
    case 2: return small_k ? tc_run<2, 32, 6>(p, Nq, w, stream) : tc_run<2, 128, 3>(p, Nq, w, stream);
    case 3: return small_k ? tc_run<3, 32, 6>(p, Nq, w, stream) : tc_run<3, 128, 3>(p, Nq, w, stream);
    case 4: return small_k ? tc_run<4, 32, 6>(p, Nq, w, stream) : tc_run<4, 128, 3>(p, Nq, w, stream);
  }
  return set_error(GGNN_B200_ERR_UNSUPPORTED, "tensor-core bf_query needs D in {32, 64, 96, 128}");
}

size_t tc_workspace_bytes(uint32_t N, uint32_t Nq, uint32_t D, uint32_t K) { return tc_layout(nullptr, N, Nq, D, tc_cap(Nq), K).total; }

}  // namespace g200
