/*
 * ggnn_oracle.c -- CPU restatement of the reference GGNN hot path (see ggnn_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
 *
 * Compile with -ffp-contract=off: every fused multiply-add the reference's SASS contains is
 * written as an explicit fmaf() here, every unfused one as separate * and + (verified against
 * `cuobjdump -sass` of the reference objects built by oracle/build_ref.sh with nvcc 12.9 for
 * sm_100a: Euclidean distance.cuh:128-135 -> FFMA chain; cosine distance.cuh:141-150 -> FMUL,
 * FSEL, FADD (not fused); simple_knn_sym_cache.cuh:163-176,225-251 -> FFMA chains).
 */
#include "ggnn_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EMPTY_KEY (-1)
#define EMPTY_DIST INFINITY
#define K_BLOCK 32u

/* number of OpenMP threads for the query-parallel loops (launchers like torchrun export OMP_NUM_THREADS=1) */
void orc_set_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static uint32_t bit_ceil_u32(uint32_t v) /* def.h:42-54 */
{
  if (v <= 1) return 1;
  v--;
  v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
  return v + 1;
}
static uint32_t next_multiple32(uint32_t v) { return v % 32 == 0 ? v : 32 * (v / 32 + 1); } /* def.h:56-60 */
static size_t align8(size_t s) { return ((s + 7) / 8) * 8; }                               /* def.h:63-66 */
static uint32_t umax(uint32_t a, uint32_t b) { return a > b ? a : b; }
static uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }

/* ------------------------------------------------------------------------------------------ */
/* GraphConfig: graph_config.cpp:39-98                                                         */
/* ------------------------------------------------------------------------------------------ */
void orc_graph_config_init(orc_graph_config* c, uint32_t N, uint32_t D, uint32_t KBuild)
{
  memset(c, 0, sizeof(*c));
  c->N = N; c->D = D; c->KBuild = KBuild;
  c->KF = KBuild / 2;
  c->S = next_multiple32(c->KF + 1);
  const int L = ORC_L;
  /* graph_config.cpp:68: std::pow(float, float) -> float overload */
  const float growth = powf((float)N / (float)c->S, 1.f / (L - 1));
  const uint32_t Gf = (uint32_t)growth;
  const uint32_t Gc = Gf + 1;
  const float S0f = (float)N / powf((float)Gf, (L - 1.0f));
  const float S0c = (float)N / powf((float)Gc, (L - 1.0f));
  const int is_floor = ((uint32_t)S0c < KBuild) || (fabsf(S0f - (float)c->S) < fabsf(S0c - (float)c->S));
  c->G = is_floor ? Gf : Gc;
  c->S0 = is_floor ? (uint32_t)S0f : (uint32_t)S0c;
  c->S0_off = N - c->G * c->G * c->G * c->S0;
  c->SG = c->S / c->G;
  c->SG_off = c->S - c->SG * c->G;
  /* GraphDimensions: graph_config.cpp:39-61 */
  uint32_t B = 1;
  for (int l = L - 1; l >= 0; --l, B *= c->G) {
    c->Bs[l] = B;
    c->Ns[l] = B * c->S;
  }
  c->Ns[0] = N;
  c->Ns_offsets[0] = 0; c->STs_offsets[0] = 0;
  c->STs_offsets[1] = 0; c->Ns_offsets[1] = N;
  for (int l = 2; l < L; ++l) {
    c->Ns_offsets[l] = c->Ns_offsets[l - 1] + c->Ns[l - 1];
    c->STs_offsets[l] = c->STs_offsets[l - 1] + c->Ns[l - 1];
  }
  c->N_all = c->Ns_offsets[L - 1] + c->Ns[L - 1];
  c->ST_all = c->STs_offsets[L - 1] + c->Ns[L - 1];
}

size_t orc_graph_blob_bytes(const orc_graph_config* c) /* graph.h:38-55 */
{
  return align8((size_t)c->N_all * c->KBuild * 4) + 2 * align8((size_t)c->ST_all * 4) + align8(8);
}

void orc_graph_view_init(orc_graph_view* v, const orc_graph_config* c, void* blob) /* graph.cpp:48-84 */
{
  char* p = (char*)blob;
  v->graph = (int32_t*)p;
  p += (size_t)c->N_all * c->KBuild * 4;
  v->translation = (int32_t*)p;
  v->selection = v->translation + c->ST_all;
  v->nn1_stats = (float*)(p + (size_t)c->ST_all * 2 * 4);
}

/* ------------------------------------------------------------------------------------------ */
/* launch parameter derivation                                                                 */
/* ------------------------------------------------------------------------------------------ */
int orc_query_launch_params(uint32_t D, uint32_t KQuery, uint32_t max_iters, orc_query_launch* o)
{ /* query_kernels.cu:63-110 */
  if (KQuery > 6000) return -1;
  const uint32_t required_sorted = next_multiple32(KQuery + 1 + 16);
  uint32_t cache = umax(256, umax(required_sorted + 32, bit_ceil_u32(max_iters)));
  const uint32_t cache_block = bit_ceil_u32((cache + 15) / 16);
  const uint32_t dim_block = bit_ceil_u32((D + 3) / 4);
  const uint32_t block = umax(32, umax(cache_block, dim_block));
  if (max_iters > 8192 || D > 4096 || cache > 8192 || block > 1024) return -1;
  o->cache_size = cache;
  o->block_dim_x = block;
  o->sorted_size = umax(cache < 512 ? 64 : 32, required_sorted);
  return 0;
}
uint32_t orc_bf_block_dim(uint32_t D) { return umax(32, bit_ceil_u32((D + 3) / 4)); } /* query_kernels.cu:213-215 */
void orc_construction_config(uint32_t D, uint32_t min_block, uint32_t* block, uint32_t* items)
{ /* graph_construction.cu:154-161 */
  *items = D <= 1024 ? 4 : 8;
  *block = umax(min_block, bit_ceil_u32((D + *items - 1) / *items));
}

/* ------------------------------------------------------------------------------------------ */
/* block reduction: cub::BlockReduce (warp shfl-down tree 1,2,4,8,16; warp aggregates summed   */
/* sequentially by thread 0) -- SURVEY.md 8(a) A1                                              */
/* ------------------------------------------------------------------------------------------ */
static float block_reduce_sum(const float* p, uint32_t vblock)
{
  float total = 0.f;
  for (uint32_t w = 0; w < vblock / 32; ++w) {
    float v[32];
    memcpy(v, p + 32 * w, sizeof(v));
    for (uint32_t off = 1; off < 32; off <<= 1)
      for (uint32_t l = 0; l + off < 32; ++l) v[l] = v[l] + v[l + off];
    total = (w == 0) ? v[0] : total + v[0];
  }
  return total;
}

typedef struct {
  uint32_t D, vblock, items;
  int measure;
  const float* base;
  const float* q;    /* query vector */
  float q_norm;      /* cosine only */
} dist_ctx;

/* distance.cuh:104-117 */
static void dist_ctx_init(dist_ctx* c, const float* base, const float* q, uint32_t D, int measure,
                          uint32_t vblock, uint32_t items)
{
  c->D = D; c->vblock = vblock; c->items = items; c->measure = measure; c->base = base; c->q = q;
  c->q_norm = 0.f;
  if (measure == ORC_COSINE) {
    float p[1024];
    for (uint32_t t = 0; t < vblock; ++t) {
      float acc = 0.f;
      for (uint32_t it = 0; it < items; ++it) {
        const uint32_t d = it * vblock + t;
        const float qv = d < D ? q[d] : 0.f;
        acc = fmaf(qv, qv, acc); /* FFMA chain (SASS bf_query 0230-0270) */
      }
      p[t] = acc;
    }
    c->q_norm = block_reduce_sum(p, vblock);
  }
}

/* distance.cuh:119-163 */
static float dist_to(const dist_ctx* c, const float* b)
{
  float p[1024], pn[1024];
  const uint32_t D = c->D, VB = c->vblock;
  if (c->measure == ORC_EUCLIDEAN) {
    for (uint32_t t = 0; t < VB; ++t) {
      float acc = 0.f;
      for (uint32_t it = 0; it < c->items; ++it) {
        const uint32_t d = it * VB + t;
        const float diff = d < D ? b[d] - c->q[d] : 0.f;
        acc = fmaf(diff, diff, acc);
      }
      p[t] = acc;
    }
    return block_reduce_sum(p, VB);
  }
  for (uint32_t t = 0; t < VB; ++t) {
    float dot = 0.f, nrm = 0.f;
    for (uint32_t it = 0; it < c->items; ++it) {
      const uint32_t d = it * VB + t;
      /* not fused in the reference SASS: FMUL, FSEL, FADD */
      const float m = d < D ? b[d] * c->q[d] : 0.f;
      const float n = d < D ? b[d] * b[d] : 0.f;
      dot = dot + m;
      nrm = nrm + n;
    }
    p[t] = dot; pn[t] = nrm;
  }
  const float dot = block_reduce_sum(p, VB);
  const float nrm = block_reduce_sum(pn, VB);
  const float norm_sqr = c->q_norm * nrm;
  return (norm_sqr > 0.0f) ? fabsf(1.0f - dot / sqrtf(norm_sqr)) : 1.0f;
}

float orc_distance(const float* q, const float* b, uint32_t D, int measure, uint32_t vblock, uint32_t items)
{
  dist_ctx c;
  dist_ctx_init(&c, NULL, q, D, measure, vblock, items);
  return dist_to(&c, b);
}

/* ------------------------------------------------------------------------------------------ */
/* KBestList: k_best_list.cuh:29-109 (lock-step emulation, BLOCK threads)                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  uint32_t K, block;
  float* dists;
  int32_t* ids;
} kbest;

static void kbest_init(kbest* l, uint32_t K, uint32_t block)
{
  l->K = K; l->block = block;
  l->dists = (float*)malloc(sizeof(float) * K);
  l->ids = (int32_t*)malloc(sizeof(int32_t) * K);
  for (uint32_t k = 0; k < K; ++k) { l->dists[k] = INFINITY; l->ids[k] = EMPTY_KEY; }
}
static void kbest_free(kbest* l) { free(l->dists); free(l->ids); }

static void kbest_add_unique(kbest* l, float dist, int32_t id) /* k_best_list.cuh:77-109 */
{
  const uint32_t B = l->block, K = l->K;
  float r_dist[1024]; int32_t r_id[1024];
  for (uint32_t i = ((K - 1) / B) * B;; i -= B) {
    for (uint32_t t = 0; t < B; ++t) {            /* read current value */
      const uint32_t k = i + t;
      if (k < K) { r_dist[t] = l->dists[k]; r_id[t] = l->ids[k]; }
    }
    /* __syncthreads */
    for (uint32_t t = 0; t < B; ++t) {            /* phase A: shift */
      const uint32_t k = i + t;
      if (k < K && dist < r_dist[t] && k < K - 1) { l->dists[k + 1] = r_dist[t]; l->ids[k + 1] = r_id[t]; }
    }
    uint8_t ins[1024];
    for (uint32_t t = 0; t < B; ++t) {            /* phase B: read left neighbour */
      const uint32_t k = i + t;
      ins[t] = (k < K && dist < r_dist[t] && (!k || l->dists[k - 1] <= dist));
    }
    for (uint32_t t = 0; t < B; ++t) {            /* phase C: insert */
      const uint32_t k = i + t;
      if (ins[t]) { l->dists[k] = dist; l->ids[k] = id; }
    }
    if (!i) break;
  }
}

void orc_bf_query(const float* base, uint32_t N_base, const float* query, uint32_t N_query,
                  uint32_t D, uint32_t K, int measure, int32_t* ids, float* dists)
{ /* bf_query_layer.cu:39-65 */
  const uint32_t B = orc_bf_block_dim(D);
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t n = 0; n < (int64_t)N_query; ++n) {
    dist_ctx dc;
    dist_ctx_init(&dc, base, query + (size_t)n * D, D, measure, B, 4);
    kbest best;
    kbest_init(&best, K, B);
    for (uint32_t i = 0; i < N_base; ++i) {
      const float d = dist_to(&dc, base + (size_t)i * D);
      if (d < best.dists[K - 1]) kbest_add_unique(&best, d, (int32_t)i);
    }
    for (uint32_t k = 0; k < K; ++k) {
      ids[(size_t)n * K + k] = best.ids[k];
      dists[(size_t)n * K + k] = best.dists[k];
    }
    kbest_free(&best);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* SimpleKNNCache: simple_knn_cache.cuh:31-352 (lock-step emulation)                           */
/* ------------------------------------------------------------------------------------------ */
struct orc_cache {
  uint32_t BEST, SORTED, CACHE, B; /* B = BLOCK_DIM_X */
  int32_t* s_cache;                /* [CACHE] */
  float* s_dists;                  /* [SORTED] */
  uint32_t prioQ_head, visited_head;
  float xi;
  /* SimpleKNNSymCache extras (simple_knn_sym_cache.cuh:80-89) */
  int sym;
  float criteria_half;
  /* statistics for roofline accounting (not in the reference) */
  uint32_t n_dist, n_pop;
};

static void cache_init(orc_cache* c) /* simple_knn_cache.cuh:73-87 */
{
  for (uint32_t i = 0; i < c->CACHE; ++i) {
    c->s_cache[i] = EMPTY_KEY;
    if (i < c->SORTED) c->s_dists[i] = EMPTY_DIST;
  }
  c->prioQ_head = c->BEST;
  c->visited_head = c->SORTED;
}

orc_cache* orc_cache_create(uint32_t best, uint32_t sorted, uint32_t cache, uint32_t vblock)
{
  orc_cache* c = (orc_cache*)calloc(1, sizeof(orc_cache));
  c->BEST = best; c->SORTED = sorted; c->CACHE = cache; c->B = vblock;
  c->s_cache = (int32_t*)malloc(sizeof(int32_t) * cache);
  c->s_dists = (float*)malloc(sizeof(float) * sorted);
  cache_init(c);
  return c;
}
void orc_cache_destroy(orc_cache* c) { free(c->s_cache); free(c->s_dists); free(c); }
void orc_cache_set_xi(orc_cache* c, float xi) { c->xi = xi; }

static float cache_criteria(const orc_cache* c)
{ /* simple_knn_cache.cuh:121-124 / simple_knn_sym_cache.cuh:285-288 */
  return (c->sym ? c->s_dists[0] : c->s_dists[c->BEST - 1]) + c->xi;
}

void orc_cache_push(orc_cache* c, int32_t key, float dist)
{ /* simple_knn_cache.cuh:126-213 == simple_knn_sym_cache.cuh:290-377 */
  const uint32_t B = c->B, SORTED = c->SORTED, BEST = c->BEST;
  for (uint32_t idx = 0; idx < SORTED; ++idx)      /* :132-146 duplicate check */
    if (c->s_cache[idx] == key) return;
  const uint32_t head = c->prioQ_head;
  const uint32_t head_in = head - BEST;
  int32_t r_cache[1024]; float r_dists[1024]; uint32_t r_idx[1024]; uint8_t active[1024], ins[1024];
  memset(active, 0, B);
  uint32_t block_start = ((SORTED + B - 1) / B) * B;
  for (;;) {
    /* phase A: shift (:164-173) */
    for (uint32_t t = 0; t < B; ++t) {
      if (!active[t] || r_cache[t] == EMPTY_KEY) continue;
      const uint32_t idx = r_idx[t];
      const uint32_t idx_next = (idx + 1 == SORTED) ? BEST : idx + 1;
      const int has_next = idx_next != BEST && idx_next != head;
      if (has_next) { c->s_cache[idx_next] = r_cache[t]; c->s_dists[idx_next] = r_dists[t]; }
    }
    /* phase B: read previous (:176-178) */
    for (uint32_t t = 0; t < B; ++t) {
      ins[t] = 0;
      if (!active[t]) continue;
      const uint32_t idx = r_idx[t];
      const int has_prev = idx != 0 && idx != head;
      const uint32_t idx_prev = idx != BEST ? idx - 1 : SORTED - 1;
      ins[t] = (!has_prev || c->s_dists[idx_prev] < dist);
    }
    /* phase C: insert (:179-182) */
    for (uint32_t t = 0; t < B; ++t)
      if (ins[t]) { c->s_cache[r_idx[t]] = key; c->s_dists[r_idx[t]] = dist; }
    if (!block_start) break;
    block_start -= B;
    for (uint32_t t = 0; t < B; ++t) { /* :189-208 */
      uint32_t idx = block_start + t;
      active[t] = idx < SORTED;
      if (active[t]) {
        if (idx >= BEST)
          idx = (idx + head_in < SORTED) ? idx + head_in : idx + head_in - SORTED + BEST;
        r_cache[t] = c->s_cache[idx];
        r_dists[t] = c->s_dists[idx];
        active[t] = active[t] && (r_dists[t] >= dist);
      }
      r_idx[t] = idx;
    }
    /* __syncthreads */
  }
}

int32_t orc_cache_pop(orc_cache* c)
{ /* simple_knn_cache.cuh:215-239 */
  const uint32_t head = c->prioQ_head;
  const int32_t key = c->s_cache[head];
  const float dist = c->s_dists[head];
  if (key == EMPTY_KEY || dist >= cache_criteria(c)) return EMPTY_KEY;
  const uint32_t vh = c->visited_head;
  c->s_cache[vh] = key;
  c->visited_head = (vh + 1) >= c->CACHE ? c->SORTED : vh + 1;
  c->s_cache[head] = EMPTY_KEY;
  c->s_dists[head] = EMPTY_DIST;
  c->prioQ_head = (head + 1) >= c->SORTED ? c->BEST : head + 1;
  c->n_pop++;
  return key;
}

void orc_cache_state(const orc_cache* c, int32_t* keys, float* dists, uint32_t* ph, uint32_t* vh)
{
  memcpy(keys, c->s_cache, sizeof(int32_t) * c->CACHE);
  memcpy(dists, c->s_dists, sizeof(float) * c->SORTED);
  *ph = c->prioQ_head; *vh = c->visited_head;
}

/* filter stage of fetch: simple_knn_cache.cuh:246-261 (with the per-thread early break) and
 * simple_knn_sym_cache.cuh:408-419 (no break) */
static void cache_filter(const orc_cache* c, int32_t* s_keys, uint32_t len)
{
  for (uint32_t t = 0; t < c->B; ++t) {
    for (uint32_t i = t; i < c->CACHE; i += c->B) {
      const int32_t n = c->s_cache[i];
      if (n == EMPTY_KEY) {
        if (!c->sym && i >= c->SORTED) break;
        continue;
      }
      for (uint32_t k = 0; k < len; ++k)
        if (s_keys[k] == n) s_keys[k] = EMPTY_KEY;
    }
  }
}

/* fetch: simple_knn_cache.cuh:241-289 */
static void cache_fetch(orc_cache* c, const dist_ctx* dc, int32_t* s_keys, const int32_t* translation,
                        uint32_t len, int filter)
{
  if (filter) cache_filter(c, s_keys, len);
  for (uint32_t k = 0; k < len; ++k) { /* ballot/ffs lane order == ascending index */
    const int32_t other_n = s_keys[k];
    if (other_n == EMPTY_KEY) continue;
    const int32_t other_m = translation ? translation[other_n] : other_n;
    const float dist = dist_to(dc, dc->base + (size_t)other_m * dc->D);
    c->n_dist++;
    if (dist < cache_criteria(c)) orc_cache_push(c, other_n, dist);
  }
}

/* transform: simple_knn_cache.cuh:297-333 */
static void cache_transform(orc_cache* c, const int32_t* transform)
{
  int32_t nk[8192]; float nd[8192];
  memcpy(nk, c->s_cache, sizeof(int32_t) * c->CACHE);
  memcpy(nd, c->s_dists, sizeof(float) * c->SORTED);
  /* threads i < BEST read s_cache[i], s_dists[i] (never written by other threads) */
  for (uint32_t i = 0; i < c->CACHE; ++i) {
    if (i < c->BEST) {
      int32_t key = c->s_cache[i];
      if (key != EMPTY_KEY) key = transform[key];
      nk[i] = key;
      if (i + c->BEST < c->SORTED) { nk[i + c->BEST] = key; nd[i + c->BEST] = c->s_dists[i]; }
    } else if (i < 2 * c->BEST && i < c->SORTED) {
      /* handled by previous threads */
    } else {
      nk[i] = EMPTY_KEY;
      if (i < c->SORTED) nd[i] = EMPTY_DIST;
    }
  }
  memcpy(c->s_cache, nk, sizeof(int32_t) * c->CACHE);
  memcpy(c->s_dists, nd, sizeof(float) * c->SORTED);
  c->prioQ_head = c->BEST;
  c->visited_head = c->SORTED;
}

/* ------------------------------------------------------------------------------------------ */
/* query: query_layer.cu:39-97                                                                 */
/* ------------------------------------------------------------------------------------------ */
void orc_query(const float* base, uint32_t N_base, const float* query, uint32_t N_query, uint32_t D,
               int measure, const int32_t* graph0, uint32_t KBuild, const int32_t* start_points,
               uint32_t n_start, const float* nn1_stats, uint32_t KQuery, float tau_query,
               uint32_t max_iterations, uint32_t shards_per_gpu, uint32_t on_gpu_shard_id,
               int32_t* ids, float* dists, uint32_t* stats)
{
  orc_query_launch lp;
  if (orc_query_launch_params(D, KQuery, max_iterations, &lp)) { fprintf(stderr, "orc_query: bad params\n"); abort(); }
  /* query_layer.cu:48-50 */
  const float xi = (measure == ORC_EUCLIDEAN) ? (nn1_stats[1] * nn1_stats[1]) * tau_query * tau_query
                                              : nn1_stats[1] * tau_query;
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t n = 0; n < (int64_t)N_query; ++n) {
    dist_ctx dc;
    dist_ctx_init(&dc, base, query + (size_t)n * D, D, measure, lp.block_dim_x, 4);
    orc_cache* c = orc_cache_create(KQuery, lp.sorted_size, lp.cache_size, lp.block_dim_x);
    c->xi = xi;
    int32_t s_knn[K_BLOCK];
    /* :55 fetch_unfiltered(d_starting_points, nullptr, S) -- chunks of 32 via the ballot loop */
    {
      int32_t* sp = (int32_t*)malloc(sizeof(int32_t) * n_start);
      memcpy(sp, start_points, sizeof(int32_t) * n_start);
      cache_fetch(c, &dc, sp, NULL, n_start, 0);
      free(sp);
    }
    for (uint32_t ite = 0; ite < max_iterations; ++ite) {
      if (measure == ORC_EUCLIDEAN) c->xi = fminf(xi, c->s_dists[0] * tau_query * tau_query);
      else c->xi = fminf(xi, c->s_dists[0] * tau_query);
      const int32_t anchor = orc_cache_pop(c);
      if (anchor == EMPTY_KEY) break;
      for (uint32_t i = 0; i < KBuild; i += K_BLOCK) {
        for (uint32_t t = 0; t < K_BLOCK; ++t)
          s_knn[t] = (i + t < KBuild) ? graph0[(size_t)anchor * KBuild + i + t] : EMPTY_KEY;
        cache_fetch(c, &dc, s_knn, NULL, K_BLOCK, 1);
      }
    }
    const size_t row = (size_t)n * shards_per_gpu + on_gpu_shard_id;
    for (uint32_t k = 0; k < KQuery; ++k) { /* :81-90, simple_knn_cache.cuh:344-352 */
      ids[row * KQuery + k] = c->s_cache[k] + (int32_t)(on_gpu_shard_id * N_base);
      dists[row * KQuery + k] = c->s_dists[k];
    }
    if (stats) { stats[2 * n] = c->n_pop; stats[2 * n + 1] = c->n_dist; }
    orc_cache_destroy(c);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* construction                                                                                */
/* ------------------------------------------------------------------------------------------ */
static const int32_t* layer_translation(const orc_graph_config* c, const orc_graph_view* g, uint32_t layer)
{ /* graph.cpp:66-75: translation[0] is empty (nullptr) */
  return layer ? g->translation + c->STs_offsets[layer] : NULL;
}

void orc_top(const orc_graph_config* c, const float* base, int measure, uint32_t layer,
             orc_graph_view* g, float* nn1_dist_buffer)
{ /* top_merge_layer.cu:40-82, launch graph_construction.cu:201-238 */
  uint32_t B, items;
  orc_construction_config(c->D, 128, &B, &items);
  const uint32_t S = layer ? c->S : c->S0;
  const uint32_t S_offset = layer ? 0 : c->S0_off;
  const int32_t* tr = layer_translation(c, g, layer);
  int32_t* graph = g->graph + (size_t)c->Ns_offsets[layer] * c->KBuild;
  const uint32_t K = c->KBuild;
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t nn = 0; nn < (int64_t)c->Ns[layer]; ++nn) {
    const uint32_t n = (uint32_t)nn;
    const int32_t m = (!layer) ? (int32_t)n : tr[n];
    dist_ctx dc;
    dist_ctx_init(&dc, base, base + (size_t)m * c->D, c->D, measure, B, items);
    kbest best;
    kbest_init(&best, K, B);
    const uint32_t S_plus_offset = S_offset * (S + 1);
    const uint32_t S_actual = (!layer && n < S_plus_offset) ? S + 1 : S;
    const int32_t start = (layer || n < S_plus_offset)
                              ? (int32_t)((n / S_actual) * S_actual)
                              : (int32_t)(S_plus_offset + ((n - S_plus_offset) / S_actual) * S_actual);
    const int32_t end = start + (int32_t)S_actual;
    for (int32_t other_n = start; other_n < end; other_n++) {
      const int32_t other_m = layer ? tr[other_n] : other_n;
      if (m == other_m) continue;
      const float dist = dist_to(&dc, base + (size_t)other_m * c->D);
      kbest_add_unique(&best, dist, other_n);
    }
    for (uint32_t k = 0; k < K; ++k) graph[(size_t)n * K + k] = best.ids[k];
    float nn1 = best.dists[1];
    if (measure == ORC_EUCLIDEAN) nn1 = sqrtf(nn1);
    nn1_dist_buffer[n] = nn1;
    kbest_free(&best);
  }
}

void orc_nn1_stats(const float* buf, uint32_t N, float* nn1_stats)
{ /* graph_construction.cu:381-393: {sum/N, max}.  The reference sums with a CUB device tree in
     fp32; this uses double accumulation -> parity on the mean is a tolerance check. */
  double sum = 0.0; float mx = -INFINITY;
  for (uint32_t i = 0; i < N; ++i) { sum += buf[i]; if (buf[i] > mx) mx = buf[i]; }
  nn1_stats[0] = (float)sum / (float)N;
  nn1_stats[1] = mx;
}

typedef struct { float key; int32_t val; uint32_t pos; } sel_item;
static int sel_cmp(const void* a, const void* b)
{ /* stable descending (cub::BlockRadixSort::SortDescending is stable) */
  const sel_item* x = (const sel_item*)a; const sel_item* y = (const sel_item*)b;
  if (x->key > y->key) return -1;
  if (x->key < y->key) return 1;
  return (x->pos > y->pos) - (x->pos < y->pos);
}

void orc_select(const orc_graph_config* c, uint32_t layer, const float* nn1_dist_buffer,
                const float* rng, orc_graph_view* g)
{ /* wrs_select_layer.cu:41-102, launch graph_construction.cu:163-187 */
  const uint32_t S = layer ? c->S : c->S0;
  const uint32_t S_offset = layer ? 0 : c->S0_off;
  int32_t* d_selection = g->selection + c->STs_offsets[layer + 1];
  int32_t* d_translation = g->translation + c->STs_offsets[layer + 1];
  const int32_t* d_translation_layer = layer_translation(c, g, layer);
  for (uint32_t b = 0; b < c->Bs[layer]; ++b) {
    const uint32_t S_current = S + (b < S_offset);
    const uint32_t start = b * S + umin(b, S_offset);
    sel_item items[256];
    for (uint32_t i = 0; i < 256; ++i) {
      if (i < S_current) {
        const int32_t n = (int32_t)(start + i);
        const float e = (-1 * logf(rng[n])) / (nn1_dist_buffer[n] + FLT_EPSILON);
        items[i].key = e; items[i].val = n;
      } else { items[i].key = -1.f; items[i].val = -1; }
      items[i].pos = i;
    }
    qsort(items, 256, sizeof(sel_item), sel_cmp);
    const uint32_t upper_segment = b / c->G;
    const uint32_t nth = b - upper_segment * c->G;
    const uint32_t num_selected = c->SG + (nth < c->SG_off);
    const uint32_t dest = upper_segment * c->S + nth * c->SG + umin(nth, c->SG_off);
    for (uint32_t s = 0; s < num_selected; ++s) {
      const int32_t n = items[s].val;
      d_selection[dest + s] = n;
      d_translation[dest + s] = (!layer) ? n : d_translation_layer[n];
    }
  }
}

void orc_merge(const orc_graph_config* c, const float* base, int measure, float tau_build,
               uint32_t layer_top, uint32_t layer_btm, orc_graph_view* g, float* nn1_dist_buffer)
{ /* merge_layer.cu:39-158, launch graph_construction.cu:240-296 */
  uint32_t B, items;
  orc_construction_config(c->D, 32, &B, &items);
  const uint32_t K = c->KBuild, S = c->S;
  const uint32_t SORTED = umax(64, next_multiple32(K + 1 + 16)); /* merge_layer.cuh:64-65 */
  const uint32_t CACHE = 256, MAX_IT = 200;
  const float xi = (measure == ORC_EUCLIDEAN) ? (g->nn1_stats[0] * g->nn1_stats[0]) * tau_build * tau_build
                                              : g->nn1_stats[0] * tau_build;
  const uint32_t Nb = c->Ns[layer_btm];
  int32_t* graph_buffer = (int32_t*)malloc(sizeof(int32_t) * (size_t)Nb * K);
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t nn = 0; nn < (int64_t)Nb; ++nn) {
    const int32_t n = (int32_t)nn;
    const int32_t m = (!layer_btm) ? n : g->translation[c->STs_offsets[layer_btm] + n];
    dist_ctx dc;
    dist_ctx_init(&dc, base, base + (size_t)m * c->D, c->D, measure, B, items);
    orc_cache* cache = orc_cache_create(K + 1, SORTED, CACHE, B);
    cache->xi = xi;
    int32_t s_knn[K_BLOCK];
    { /* get_top_seg_offset merge_layer.cu:41-62 */
      uint32_t seg_btm = (uint32_t)n / S;
      if (!layer_btm) {
        const uint32_t offset_points = c->S0_off * (c->S0 + 1);
        seg_btm = ((uint32_t)n < offset_points) ? (uint32_t)n / (c->S0 + 1)
                                                 : c->S0_off + ((uint32_t)n - offset_points) / c->S0;
      }
      uint32_t powG = c->G;
      for (uint32_t i = 1; i < layer_top - layer_btm; ++i) powG *= c->G;
      const uint32_t s_offset = (seg_btm / powG) * S;
      for (uint32_t i = 0; i < S; i += K_BLOCK) {
        for (uint32_t t = 0; t < K_BLOCK; ++t)
          s_knn[t] = (i + t < S) ? (int32_t)(s_offset + i + t) : EMPTY_KEY;
        cache_fetch(cache, &dc, s_knn, g->translation + c->STs_offsets[layer_top], K_BLOCK, 0);
      }
    }
    for (uint32_t layer = layer_top - 1; layer >= layer_btm && layer != (uint32_t)-1; layer--) {
      cache_transform(cache, g->selection + c->STs_offsets[layer + 1]);
      const int32_t* tr = (!layer) ? NULL : g->translation + c->STs_offsets[layer];
      if (layer == layer_btm) {
        int32_t self = n;
        cache_fetch(cache, &dc, &self, tr, 1, 0);
      }
      for (uint32_t ite = 0; ite < MAX_IT; ++ite) {
        const int32_t anchor = orc_cache_pop(cache);
        if (anchor == EMPTY_KEY) break;
        for (uint32_t j = 0; j < K; j += K_BLOCK) {
          for (uint32_t t = 0; t < K_BLOCK; ++t)
            s_knn[t] = (j + t < K) ? g->graph[((size_t)c->Ns_offsets[layer] + anchor) * K + j + t] : EMPTY_KEY;
          cache_fetch(cache, &dc, s_knn, tr, K_BLOCK, 1);
        }
      }
    }
    /* :122-145 */
    int32_t s_own_idx = -1;
    for (uint32_t k = 0; k < K; ++k)
      if (cache->s_cache[k] == n) s_own_idx = (int32_t)k;
    for (uint32_t k = 0; k < K; ++k) {
      const int32_t idx = cache->s_cache[k + ((int32_t)k >= s_own_idx)];
      graph_buffer[(size_t)n * K + k] = (idx != EMPTY_KEY) ? idx : n;
    }
    if (!layer_btm) { /* :147-157 */
      uint32_t i = (uint32_t)(s_own_idx + 1);
      float dist;
      do { dist = cache->s_dists[i]; ++i; } while (dist == 0.0f && i < cache->BEST);
      if (measure == ORC_EUCLIDEAN) dist = sqrtf(dist);
      nn1_dist_buffer[n] = dist;
    }
    orc_cache_destroy(cache);
  }
  memcpy(g->graph + (size_t)c->Ns_offsets[layer_btm] * K, graph_buffer, sizeof(int32_t) * (size_t)Nb * K);
  free(graph_buffer);
}

/* ---- SimpleKNNSymCache distances: simple_knn_sym_cache.cuh:143-283 ---- */
typedef struct {
  uint32_t D, vblock, items;
  int measure;
  const float* base;
  const float* q;
  float* half;           /* [D] */
  float q_norm, half_norm;
} sym_ctx;

static void sym_distance(const sym_ctx* s, const float* o, float* d_query, float* d_half)
{ /* :214-283 */
  float pq[1024], ph[1024], pn[1024];
  const uint32_t D = s->D, VB = s->vblock;
  for (uint32_t t = 0; t < VB; ++t) {
    float aq = 0.f, ah = 0.f, an = 0.f;
    for (uint32_t it = 0; it < s->items; ++it) {
      const uint32_t d = it * VB + t;
      if (d >= D) continue;
      if (s->measure == ORC_EUCLIDEAN) {
        const float dq = s->q[d] - o[d];
        aq = fmaf(dq, dq, aq);
        const float dh = s->half[d] - o[d];
        ah = fmaf(dh, dh, ah);
      } else {
        aq = fmaf(s->q[d], o[d], aq);
        ah = fmaf(s->half[d], o[d], ah);
        an = fmaf(o[d], o[d], an);
      }
    }
    pq[t] = aq; ph[t] = ah; pn[t] = an;
  }
  float dq = block_reduce_sum(pq, VB), dh = block_reduce_sum(ph, VB);
  if (s->measure == ORC_COSINE) {
    const float norm_other = block_reduce_sum(pn, VB);
    const float qn = norm_other * s->q_norm, hn = norm_other * s->half_norm;
    dq = (qn > 0.0f) ? fabsf(1.0f - dq / sqrtf(qn)) : 1.0f;
    dh = (hn > 0.0f) ? fabsf(1.0f - dh / sqrtf(hn)) : 1.0f;
  }
  *d_query = dq; *d_half = dh;
}

void orc_sym(const orc_graph_config* c, const float* base, int measure, float tau_build,
             uint32_t layer, const orc_graph_view* g, int32_t* sym_buffer, uint32_t* sym_atomic)
{ /* sym_query_layer.cu:39-145, launch graph_construction.cu:298-352 */
  uint32_t B, items;
  orc_construction_config(c->D, 64, &B, &items);
  const uint32_t K = c->KBuild, KF = K / 2, KL = K - KF, D = c->D;
  const uint32_t CACHE = 128, MAX_PATH = 20;
  const uint32_t sorted = umax(64, next_multiple32(K / 2 + 16)); /* sym_query_layer.cuh:58-59 */
  const float xi = (measure == ORC_EUCLIDEAN) ? (g->nn1_stats[0] * g->nn1_stats[0]) * tau_build * tau_build
                                              : g->nn1_stats[0] * tau_build;
  const int32_t* tr = layer_translation(c, g, layer);
  const int32_t* graph = g->graph + (size_t)c->Ns_offsets[layer] * K;
  const float EPS = 0.1f;
  const float half_w = 0.5f - EPS;
  float* half = (float*)malloc(sizeof(float) * D);
  orc_cache* cache = orc_cache_create(KF, sorted, CACHE, B);
  cache->sym = 1;
  cache->xi = xi;
  for (uint32_t n = 0; n < c->Ns[layer]; ++n) {
    sym_ctx s;
    s.D = D; s.vblock = B; s.items = items; s.measure = measure; s.base = base; s.half = half;
    s.q = base + (size_t)(tr ? tr[n] : (int32_t)n) * D;
    s.q_norm = 0.f; s.half_norm = 0.f;
    for (uint32_t k = 0; k < KL; ++k) {
      const int32_t start_n = graph[(size_t)n * K + k];
      /* init_start_point simple_knn_sym_cache.cuh:159-201 */
      const int32_t start_m = tr ? tr[start_n] : start_n;
      const float* sv = base + (size_t)start_m * D;
      for (uint32_t d = 0; d < D; ++d) half[d] = fmaf(sv[d] - s.q[d], half_w, s.q[d]);
      if (measure == ORC_COSINE) {
        float pq[1024], ph[1024];
        for (uint32_t t = 0; t < B; ++t) {
          float aq = 0.f, ah = 0.f;
          for (uint32_t it = 0; it < items; ++it) {
            const uint32_t d = it * B + t;
            if (d >= D) continue;
            aq = fmaf(s.q[d], s.q[d], aq);
            ah = fmaf(half[d], half[d], ah);
          }
          pq[t] = aq; ph[t] = ah;
        }
        s.q_norm = block_reduce_sum(pq, B);
        s.half_norm = block_reduce_sum(ph, B);
      }
      float dq0, dh0;
      sym_distance(&s, sv, &dq0, &dh0);
      cache->criteria_half = dh0 + xi;
      for (uint32_t i = 0; i < CACHE; ++i) {
        cache->s_cache[i] = (i == 0 || i == KF) ? start_n : EMPTY_KEY;
        if (i < sorted) cache->s_dists[i] = (i == 0 || i == KF) ? dq0 : EMPTY_DIST;
      }
      cache->prioQ_head = KF;
      cache->visited_head = sorted;

      int found = 0;
      for (uint32_t ite = 0; ite < MAX_PATH && !found; ++ite) {
        const int32_t anchor = orc_cache_pop(cache);
        if (anchor == EMPTY_KEY) break;
        for (uint32_t i = 0; i < K; i += K_BLOCK) {
          int32_t s_knn[K_BLOCK];
          int connected = 0;
          for (uint32_t t = 0; t < K_BLOCK; ++t) {
            const uint32_t kk = i + t;
            if (kk < K) {
              const int32_t other = (kk < KL) ? graph[(size_t)anchor * K + kk]
                                              : sym_buffer[(size_t)anchor * KF + kk - KL];
              if (other == (int32_t)n) connected = 1;
              s_knn[t] = other;
            } else s_knn[t] = EMPTY_KEY;
          }
          if (connected) { found = 1; break; }
          /* fetch simple_knn_sym_cache.cuh:405-436 */
          cache_filter(cache, s_knn, K_BLOCK);
          for (uint32_t t = 0; t < K_BLOCK; ++t) {
            const int32_t other_n = s_knn[t];
            if (other_n == EMPTY_KEY) continue;
            const int32_t other_m = tr ? tr[other_n] : other_n;
            float dq, dh;
            sym_distance(&s, base + (size_t)other_m * D, &dq, &dh);
            if (dq < cache_criteria(cache) && dh < cache->criteria_half) orc_cache_push(cache, other_n, dq);
          }
        }
      }
      if (!found) { /* :121-141 */
        for (uint32_t i = 0; i < KF; i++) {
          const int32_t other_n = cache->s_cache[i];
          if (other_n == EMPTY_KEY) break;
          const uint32_t pos = sym_atomic[other_n]++;
          if (pos < KF) { sym_buffer[(size_t)other_n * KF + pos] = (int32_t)n; break; }
        }
      }
    }
  }
  orc_cache_destroy(cache);
  free(half);
}

void orc_sym_buffer_merge(const orc_graph_config* c, uint32_t layer, const int32_t* sym_buffer,
                          const uint32_t* sym_atomic, orc_graph_view* g)
{ /* sym_buffer_merge_layer.cu:36-99 */
  const uint32_t K = c->KBuild, KF = K / 2, KL = K - KF;
  int32_t* graph = g->graph + (size_t)c->Ns_offsets[layer] * K;
  for (uint32_t n = 0; n < c->Ns[layer]; ++n) {
    int32_t s_sym[512], s_graph[512];
    uint32_t num_links = sym_atomic[n];
    for (uint32_t kf = 0; kf < KF; ++kf) {
      s_sym[kf] = sym_buffer[(size_t)n * KF + kf];
      s_graph[kf] = graph[(size_t)n * K + KL + kf];
    }
    for (uint32_t i = 0; i < KF; i++) {
      int found = num_links >= KF;
      const int32_t r_graph = s_graph[i];
      if (!found)
        for (uint32_t kf = 0; kf < KF; ++kf)
          if (r_graph == s_sym[kf]) found = 1;
      if (!found) { s_sym[num_links] = r_graph; ++num_links; }
    }
    for (uint32_t kf = 0; kf < KF; ++kf) {
      const int32_t res = s_sym[kf];
      graph[(size_t)n * K + KL + kf] = (res >= 0) ? res : (int32_t)n;
    }
  }
}

static void orc_sym_pass(const orc_graph_config* c, const float* base, int measure, float tau_build,
                         uint32_t layer, orc_graph_view* g)
{ /* graph_construction.cu:298-352 */
  const uint32_t KF = c->KBuild / 2;
  int32_t* sb = (int32_t*)malloc(sizeof(int32_t) * (size_t)c->Ns[layer] * KF);
  uint32_t* sa = (uint32_t*)calloc(c->Ns[layer], sizeof(uint32_t));
  memset(sb, 0xff, sizeof(int32_t) * (size_t)c->Ns[layer] * KF);
  orc_sym(c, base, measure, tau_build, layer, g, sb, sa);
  orc_sym_buffer_merge(c, layer, sb, sa, g);
  free(sb); free(sa);
}

void orc_build(const orc_graph_config* c, const float* base, int measure, float tau_build,
               uint32_t refinement_iterations, const float* rng, void* blob)
{ /* graph_construction.cu:128-147 + gpu_instance.cu:550-555 */
  orc_graph_view g;
  orc_graph_view_init(&g, c, blob);
  float* nn1 = (float*)malloc(sizeof(float) * c->N);
  const float* rng_layer = rng;
  for (uint32_t layer_top = 0; layer_top < ORC_L; layer_top++) {
    for (uint32_t layer_btm = layer_top; layer_btm != (uint32_t)-1; layer_btm--) {
      if (layer_top == layer_btm) orc_top(c, base, measure, layer_btm, &g, nn1);
      else orc_merge(c, base, measure, tau_build, layer_top, layer_btm, &g, nn1);
      if (!layer_btm) orc_nn1_stats(nn1, c->N, g.nn1_stats);
      if (layer_top < ORC_L - 1 && layer_top == layer_btm) {
        orc_select(c, layer_top, nn1, rng_layer, &g);
        rng_layer += c->Ns[layer_top];
      }
      orc_sym_pass(c, base, measure, tau_build, layer_btm, &g);
    }
  }
  for (uint32_t r = 0; r < refinement_iterations; ++r) {
    for (uint32_t layer = ORC_L - 2; layer != (uint32_t)-1; layer--) {
      orc_merge(c, base, measure, tau_build, ORC_L - 1, layer, &g, nn1);
      if (!layer) orc_nn1_stats(nn1, c->N, g.nn1_stats);
      orc_sym_pass(c, base, measure, tau_build, layer, &g);
    }
  }
  free(nn1);
}

/* ------------------------------------------------------------------------------------------ */
/* result merge: result_merger.cpp:51-149.  The reference's tie order is arbitrary (heap with  */
/* >=); this oracle breaks ties by (partition, position).                                      */
/* ------------------------------------------------------------------------------------------ */
void orc_merge_results(const int32_t* ids, const float* dists, uint32_t n_parts, uint32_t N_query,
                       uint32_t K_in, uint32_t K, uint32_t spg_N_shard, int32_t* out_ids, float* out_dists)
{
  uint32_t* pos = (uint32_t*)malloc(sizeof(uint32_t) * n_parts);
  for (uint32_t n = 0; n < N_query; ++n) {
    memset(pos, 0, sizeof(uint32_t) * n_parts);
    for (uint32_t k = 0; k < K; ++k) {
      uint32_t best_p = 0; float best_d = INFINITY; int have = 0;
      for (uint32_t p = 0; p < n_parts; ++p) {
        if (pos[p] >= K_in) continue;
        const float d = dists[((size_t)p * N_query + n) * K_in + pos[p]];
        if (!have || d < best_d) { best_d = d; best_p = p; have = 1; }
      }
      const size_t src = ((size_t)best_p * N_query + n) * K_in + pos[best_p];
      out_ids[(size_t)n * K + k] = (int32_t)(best_p * spg_N_shard) + ids[src];
      out_dists[(size_t)n * K + k] = dists[src];
      pos[best_p]++;
    }
  }
  free(pos);
}

/* ------------------------------------------------------------------------------------------ */
/* Evaluator: eval.cpp:37-65, 88-242                                                           */
/* ------------------------------------------------------------------------------------------ */
static float eval_distance(const float* a, const float* b, uint32_t D, int measure)
{ /* eval.cpp:37-65 (a = base vector, b = query; note b_norm uses a -- reference quirk :52) */
  float distance = 0.0f, a_norm = 0.0f, b_norm = 0.0f;
  for (uint32_t d = 0; d < D; ++d) {
    if (measure == ORC_EUCLIDEAN) distance += (a[d] - b[d]) * (a[d] - b[d]);
    else { distance += a[d] * b[d]; a_norm += a[d] * a[d]; b_norm += a[d] * a[d]; }
  }
  if (measure == ORC_EUCLIDEAN) distance = sqrtf(distance);
  else distance = (a_norm * b_norm > 0.0f) ? fabsf(1.0f - distance / sqrtf(a_norm * b_norm)) : 1.0f;
  return distance;
}

void orc_eval(const float* base, uint32_t N_base, const float* query, uint32_t N_query, uint32_t D,
              int measure, const int32_t* gt, uint32_t K_gt, const int32_t* results, uint32_t KQuery, float* out)
{
  (void)N_base;
  const float Epsilon = 0.000001f;
  uint32_t c1 = 0, c1_dup = 0, cK = 0, cK_dup = 0, rK = 0, rK_dup = 0;
  for (uint32_t n = 0; n < N_query; ++n) {
    uint32_t endTop1 = 1, endTopK = KQuery;
    if (base && query) { /* eval.cpp:135-167 */
      const float* q = query + (size_t)n * D;
      const float d1 = eval_distance(base + (size_t)gt[(size_t)n * K_gt] * D, q, D, measure);
      uint32_t dup1 = 0, dupk = 0;
      for (uint32_t k = 1; k < K_gt; ++k) {
        const float dk = eval_distance(base + (size_t)gt[(size_t)n * K_gt + k] * D, q, D, measure);
        if (dk - d1 > Epsilon) break;
        ++dup1;
      }
      endTop1 = 1 + dup1;
      if (KQuery <= K_gt) {
        const float dK = eval_distance(base + (size_t)gt[(size_t)n * K_gt + KQuery - 1] * D, q, D, measure);
        for (uint32_t k = KQuery; k < K_gt; ++k) {
          const float dk = eval_distance(base + (size_t)gt[(size_t)n * K_gt + k] * D, q, D, measure);
          if (dk - dK > Epsilon) break;
          ++dupk;
        }
        endTopK = KQuery + dupk;
      } else endTopK = K_gt;
    }
    for (uint32_t kr = 0; kr < KQuery; kr++) { /* eval.cpp:202-226 */
      const int32_t q = results[(size_t)n * KQuery + kr];
      for (uint32_t kg = 0; kg < endTopK; kg++) {
        if (q == gt[(size_t)n * K_gt + kg]) {
          if (!kg) { if (!kr) ++c1; if (kg < KQuery) ++rK; ++rK_dup; }
          if (kg < endTop1 && !kr) ++c1_dup;
          if (kg < KQuery) ++cK;
          ++cK_dup;
        }
      }
    }
  }
  const float inv_q = 1.0f / (float)N_query;
  const float inv_r = 1.0f / (float)(N_query * KQuery);
  out[0] = (float)c1 * inv_q; out[1] = (float)c1_dup * inv_q;
  out[2] = (float)cK * inv_r; out[3] = (float)cK_dup * inv_r;
  out[4] = (float)rK * inv_q; out[5] = (float)rK_dup * inv_q;
}
