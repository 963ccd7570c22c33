/*
 * ggnn_b200.h -- C ABI of libggnn_b200.so: the B200 (sm_100a) implementation of GGNN's batched
 * query hot path and graph-construction kernels.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Every kernel entry
 * point carries exactly the fields of the reference's by-value kernel parameter aggregate it
 * replaces (cited per struct, paths relative to the reference tree), is asynchronous on the given
 * CUDA stream, allocates nothing, keeps no global mutable state and requires the right device to
 * be current -- the same contract as the reference's thin host->CUDA boundary
 * (include/ggnn/query/query_kernels.cuh:47-57, include/ggnn/construction/graph_construction.cuh:41-55).
 *
 * All device pointers must be 16-byte aligned.  Keys are int32, distances fp32, base vectors fp32.
 * Return value: 0 (cudaSuccess) or a cudaError_t / GGNN_B200_ERR_* code; the message of the last
 * error of the calling thread is available from ggnn_b200_last_error().
 * There is no CPU fallback anywhere behind this interface.
 */
#ifndef GGNN_B200_H
#define GGNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GGNN_B200_EUCLIDEAN 0 /* include/ggnn/base/def.h:27-30 */
#define GGNN_B200_COSINE 1
#define GGNN_B200_L 4 /* include/ggnn/base/graph_config.h:42-43 */

#define GGNN_B200_ERR_INVALID 10001     /* precondition the reference CHECKs (would abort there) */
#define GGNN_B200_ERR_UNSUPPORTED 10002 /* valid in the reference, not built here yet */

typedef void* ggnn_b200_stream_t; /* cudaStream_t */

const char* ggnn_b200_last_error(void);
/* "sm_100a" + build info; never NULL */
const char* ggnn_b200_version(void);

/* ------------------------------------------------------------------------------------------------
 * Graph configuration / blob layout (host only).
 * Replaces ggnn::GraphConfig (include/ggnn/base/graph_config.h:32-107, src/ggnn/base/graph_config.cpp:39-98)
 * and ggnn::Graph::PartSizes / layout (include/ggnn/base/graph.h:38-55, src/ggnn/base/graph.cpp:33-92).
 * The blob is byte-compatible with the reference's part_<id>.ggnn files.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t N, D, KBuild;
  uint32_t KF, G, S, S0, S0_off, SG, SG_off;
  uint32_t N_all, ST_all;
  uint32_t Bs[GGNN_B200_L], Ns[GGNN_B200_L], Ns_offsets[GGNN_B200_L], STs_offsets[GGNN_B200_L];
} ggnn_b200_graph_config;

int ggnn_b200_graph_config_init(ggnn_b200_graph_config* cfg, uint32_t N, uint32_t D, uint32_t KBuild);
size_t ggnn_b200_graph_blob_bytes(const ggnn_b200_graph_config* cfg);
/* byte offsets of the parts inside the blob: neighbourhoods [N_all,K] | translation [ST_all] |
 * selection [ST_all] | nn1_stats [2] */
typedef struct {
  size_t graph, translation, selection, nn1_stats, total;
} ggnn_b200_graph_offsets;
void ggnn_b200_graph_blob_offsets(const ggnn_b200_graph_config* cfg, ggnn_b200_graph_offsets* out);
/* scratch needed by ggnn_b200_build_graph (replaces GraphBuffer, src/ggnn/construction/graph_buffer.cu:38-81) */
size_t ggnn_b200_build_scratch_bytes(const ggnn_b200_graph_config* cfg);

/* derived launch shape of the reference's query kernel (src/ggnn/query/query_kernels.cu:77-110).
 * block_dim_x fixes the floating-point summation order of every distance (cub::BlockReduce over
 * block_dim_x threads, 4 dims per thread), so it is part of the results contract. */
typedef struct {
  uint32_t cache_size, sorted_size, block_dim_x;
} ggnn_b200_query_shape;
int ggnn_b200_query_shape_init(ggnn_b200_query_shape* s, uint32_t D, uint32_t KQuery, uint32_t max_iterations);

/* ------------------------------------------------------------------------------------------------
 * ANN query.  Replaces ggnn::QueryKernel (include/ggnn/query/query_layer.cuh:56-85,
 * src/ggnn/query/query_layer.cu:39-97) and its launcher QueryKernelsImpl::query
 * (src/ggnn/query/query_kernels.cu:50-186).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t D;
  int32_t measure;
  uint32_t KQuery;
  uint32_t sorted_size; /* 0 = derive like the reference */
  uint32_t cache_size;  /* 0 = derive like the reference */
  uint32_t block_dim_x; /* 0 = derive like the reference */
  float tau_query;
  uint32_t max_iterations;
  int32_t N_base;
  uint32_t KBuild;
  uint32_t num_starting_points;
  const float* d_base;              /* [N_base, D] */
  const float* d_query;             /* [N_query, D] */
  const int32_t* d_graph;           /* [N_base, KBuild] layer-0 neighbourhoods */
  const int32_t* d_starting_points; /* [num_starting_points] = translation[L-1] */
  const float* d_nn1_stats;         /* [2] = {mean, max} */
  int32_t* d_query_results;         /* [N_query, KQuery * shards_per_gpu] */
  float* d_query_results_dists;     /* same shape; may be NULL */
  uint32_t* d_stats;                /* optional [N_query, 2] = {pops, distance evaluations} (reference: d_dist_stats, dead) */
  uint32_t shards_per_gpu;          /* >= 1 */
  uint32_t on_gpu_shard_id;
  uint32_t* d_work_counter;         /* optional device uint32 used for dynamic query scheduling
                                       (zeroed by the call on `stream`); NULL = static mapping */
  /* --- fused shard-merge exchange (optional; all zero = off).  Replaces the D2H copy + CPU heap merge of
   * src/ggnn/base/result_merger.cpp:51-149 / gpu_instance.cu:714-742: the kernel's epilogue stores each query's
   * KQuery (id, dist) pairs straight into the gathered buffer of every destination GPU (peer-mapped memory: NVLink
   * stores), and the last warp of the launch bumps a flag word in every destination.  A destination buffer holds
   * ids [n_slots][scatter_rows][KQuery] int32 and, scatter_dists_offset bytes further, dists of the same shape;
   * this launch writes list `scatter_slot` with shard-local ids (no id offset).  d_query_results may be NULL then. */
  uint32_t n_scatter;               /* number of destination buffers */
  uint32_t scatter_slot;
  uint32_t scatter_rows;            /* >= N_query */
  size_t scatter_dists_offset;
  void* const* d_scatter_dst;       /* device array [n_scatter] of destination buffer base pointers */
  uint32_t* const* d_scatter_flags; /* device array [n_scatter] of flag words (one per destination), or NULL */
  uint32_t* d_scatter_done;         /* local device word, zero before the first use (reset by every launch) */
  /* --- native uint8 vectors (the reference's BaseT = uint8_t instantiation, include/ggnn/base/lib.h:26-28): with
   * base_type = GGNN_B200_BASE_U8, d_base and d_query point to uint8 rows of D bytes (a quarter of the gather traffic).
   * The reference computes every distance on static_cast<float>(value) (distance.cuh:104-148); for D <= 256 all
   * partial sums are integers below 2^24, i.e. exact in fp32, so integer arithmetic (dp4a) gives bit-identical
   * distances.  Shapes the native kernel does not cover (it is built for D in {32, 64, 96, 128, 256} and needs KQuery <= 47) return
   * GGNN_B200_ERR_UNSUPPORTED (widen with
   * ggnn_b200_widen_u8 and use the fp32 path: identical results). */
  uint32_t base_type;
  /* --- optional interleaved copy of an fp32 base (ggnn_b200_interleave_rows; D = 128, Euclidean, register-resident
   * lists): row n holds element 32c + t at position 4t + c, so that the four elements a lane accumulates in the
   * reference's order (dims t, t+32, t+64, t+96; distance.cuh:104-139) are ONE 16-byte shared-memory load.  Same results;
   * costs a second copy of the base.  NULL = rows are read in their natural layout. */
  const float* d_base_interleaved;
} ggnn_b200_query_params;
#define GGNN_B200_BASE_F32 0
#define GGNN_B200_BASE_U8 1

int ggnn_b200_query(const ggnn_b200_query_params* p, uint32_t N_query, ggnn_b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Brute-force query.  Replaces ggnn::BruteForceQueryKernel (include/ggnn/query/bf_query_layer.cuh:54-64,
 * src/ggnn/query/bf_query_layer.cu:39-65) and QueryKernelsImpl::bruteForceQuery (query_kernels.cu:188-264).
 * Results are the K smallest (distance, index) pairs in the reference's own fp32 arithmetic.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t D;
  int32_t measure;
  uint32_t KQuery;
  int32_t N_base;
  const float* d_base;          /* [N_base, D] */
  const float* d_query;         /* [N_query, D] */
  int32_t* d_query_results;     /* [N_query, KQuery] */
  float* d_query_results_dists; /* [N_query, KQuery]; may be NULL */
  void* d_workspace;            /* optional scratch of >= ggnn_b200_bf_query_workspace_bytes(): enables the
                                   tensor-core path (tcgen05 3xTF32 contraction + exact re-rank, identical
                                   results); NULL = exact SIMT scan */
  size_t workspace_bytes;
} ggnn_b200_bf_query_params;

/* 0 if the tensor-core path does not apply to this shape (Euclidean, D in {32,64,96,128}, KQuery <= 32) */
size_t ggnn_b200_bf_query_workspace_bytes(uint32_t D, int32_t measure, uint32_t KQuery, uint32_t N_base, uint32_t N_query);

int ggnn_b200_bf_query(const ggnn_b200_bf_query_params* p, uint32_t N_query, ggnn_b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Graph construction kernels (one entry point per reference kernel) and the full schedule.
 * d_graph_blob is a device blob laid out by ggnn_b200_graph_blob_offsets().
 * ---------------------------------------------------------------------------------------------- */
/* TopMergeKernel: include/ggnn/construction/top_merge_layer.cuh:48-61, src/.../top_merge_layer.cu:40-82 */
int ggnn_b200_top(const ggnn_b200_graph_config* cfg, const float* d_base, int32_t measure, uint32_t layer,
                  void* d_graph_blob, float* d_nn1_dist_buffer, ggnn_b200_stream_t stream);
/* cub::DeviceReduce Sum/Max + divide: src/ggnn/construction/graph_construction.cu:381-393, 79-83.
 * d_scratch: >= 8 KiB */
int ggnn_b200_nn1_stats(const float* d_nn1_dist_buffer, uint32_t N, float* d_nn1_stats, void* d_scratch,
                        ggnn_b200_stream_t stream);
/* WRSSelectionKernel: include/.../wrs_select_layer.cuh:51-66, src/.../wrs_select_layer.cu:41-102.
 * d_rng: Ns[layer] uniforms in (0,1] */
int ggnn_b200_select(const ggnn_b200_graph_config* cfg, uint32_t layer, const float* d_nn1_dist_buffer,
                     const float* d_rng, void* d_graph_blob, ggnn_b200_stream_t stream);
/* MergeKernel + publish copy: include/.../merge_layer.cuh:61-88, src/.../merge_layer.cu:39-158,
 * graph_construction.cu:240-296.  d_graph_buffer: [Ns[layer_btm], KBuild] scratch */
int ggnn_b200_merge(const ggnn_b200_graph_config* cfg, const float* d_base, int32_t measure, float tau_build,
                    uint32_t layer_top, uint32_t layer_btm, void* d_graph_blob, int32_t* d_graph_buffer,
                    float* d_nn1_dist_buffer, ggnn_b200_stream_t stream);
/* SymQueryKernel (+ the two memsets before it): include/.../sym_query_layer.cuh:54-70,
 * src/.../sym_query_layer.cu:39-145, graph_construction.cu:298-352 */
int ggnn_b200_sym(const ggnn_b200_graph_config* cfg, const float* d_base, int32_t measure, float tau_build,
                  uint32_t layer, void* d_graph_blob, int32_t* d_sym_buffer, uint32_t* d_sym_atomic,
                  ggnn_b200_stream_t stream);
/* SymBufferMergeKernel: include/.../sym_buffer_merge_layer.cuh:45-53, src/.../sym_buffer_merge_layer.cu:36-99 */
int ggnn_b200_sym_buffer_merge(const ggnn_b200_graph_config* cfg, uint32_t layer, const int32_t* d_sym_buffer,
                               const uint32_t* d_sym_atomic, void* d_graph_blob, ggnn_b200_stream_t stream);
/* GraphConstruction::build + refine x refinement_iterations (graph_construction.cu:128-147,
 * gpu_instance.cu:550-555).  d_rng: NULL = draw with cuRAND XORWOW seed 1234 like the reference
 * (graph_construction.cu:96-102,168-169); else Ns[0]+Ns[1]+Ns[2] uniforms consumed layer by layer. */
int ggnn_b200_build_graph(const ggnn_b200_graph_config* cfg, const float* d_base, int32_t measure, float tau_build,
                          uint32_t refinement_iterations, const float* d_rng, void* d_graph_blob, void* d_scratch,
                          size_t scratch_bytes, ggnn_b200_stream_t stream);
/* one refinement pass (GraphConstructionImpl::refine, graph_construction.cu:141-147): merge(L-1 -> layer) + sym for
 * layer = L-2 .. 0 on a graph built before; same scratch as ggnn_b200_build_graph */
int ggnn_b200_refine_graph(const ggnn_b200_graph_config* cfg, const float* d_base, int32_t measure, float tau_build,
                           void* d_graph_blob, void* d_scratch, size_t scratch_bytes, ggnn_b200_stream_t stream);

/* Diagnostics (roofline accounting of the construction kernels): between _begin and _end every
 * ggnn_b200_merge / ggnn_b200_sym launch of the calling thread records its traversal counters and its time on the stream.
 * Algorithmic bytes of a pass = 4*D*(points + dists) + 4*KBuild*pops (merge; sym evaluates two distances per row read).
 * _begin allocates a small device buffer, _end synchronises the device and frees it. */
#define GGNN_B200_MAX_BUILD_PASSES 64
typedef struct {
  uint32_t kernel;    /* 0 = merge, 1 = sym */
  uint32_t layer_top, layer_btm, points;
  uint64_t pops, dists;
  float ms;
} ggnn_b200_build_pass_stats;
int ggnn_b200_build_stats_begin(void);
int ggnn_b200_build_stats_end(ggnn_b200_build_pass_stats* out, uint32_t max_passes, uint32_t* n_passes);

/* The reference keeps ONE cuRAND generator (XORWOW, seed 1234) per GPU whose sequence continues over all shards built on
 * that GPU (graph_construction.cu:96-102,127).  A caller that wants the reference's selection for every shard owns such a
 * generator and draws the uniforms of a build itself: ggnn_b200_rng_fill_build() makes the reference's three
 * curandGenerateUniform calls (Ns[0], Ns[1], Ns[2] values, graph_construction.cu:168-169) into d_rng, which is then
 * passed to ggnn_b200_build_graph.  (d_rng = NULL there = a fresh generator per call: right for the first shard only.) */
typedef struct ggnn_b200_rng ggnn_b200_rng;
int ggnn_b200_rng_create(ggnn_b200_rng** out, uint64_t seed);
int ggnn_b200_rng_fill_build(ggnn_b200_rng* rng, const ggnn_b200_graph_config* cfg, float* d_rng, ggnn_b200_stream_t stream);
void ggnn_b200_rng_destroy(ggnn_b200_rng* rng);

/* ------------------------------------------------------------------------------------------------
 * Shard merge.  Replaces the per-GPU cub::DeviceSegmentedRadixSort (src/ggnn/base/gpu_instance.cu:745-790)
 * and the CPU heap merge ResultMerger::merge (src/ggnn/base/result_merger.cpp:51-149) with one kernel:
 * n_lists sorted runs of K_in (id, dist) per query -> the K smallest.  List l of query n starts at
 * d_ids + (l * list_stride + n * query_stride); ids get id_offset_per_list * l added
 * (result_merger.cpp:115-116).  Ties: lower list index first (the reference's order is arbitrary).
 * ---------------------------------------------------------------------------------------------- */
int ggnn_b200_merge_topk(const int32_t* d_ids, const float* d_dists, uint32_t n_lists, size_t list_stride,
                         size_t query_stride, uint32_t K_in, uint32_t N_query, uint32_t K,
                         int64_t id_offset_per_list, int32_t* d_out_ids, float* d_out_dists,
                         ggnn_b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Peer-memory plumbing of the fused shard-merge exchange (one process per GPU: CUDA IPC; one process with several
 * GPUs: peer access).  The reference has no device-side exchange (per-GPU D2H + CPU merge, gpu_instance.cu:714-742).
 * ggnn_b200_ipc_alloc: cudaMalloc + zero fill + cudaIpcGetMemHandle (handle = 64 bytes, to be sent to the peers);
 * ggnn_b200_ipc_open / _close: map / unmap a peer's allocation in this process; ggnn_b200_ipc_free: cudaFree.
 * ggnn_b200_peer_enable: cudaDeviceEnablePeerAccess(peer) for the current device (already enabled = success).
 * ggnn_b200_wait_flag: stream-ordered wait until *d_flag >= expected (one polling thread, no SM time to speak of);
 * after timeout_ms (0 = 10 s) it gives up, stores 1 to *d_timed_out (may be NULL) and lets the stream continue.
 * ---------------------------------------------------------------------------------------------- */
int ggnn_b200_ipc_alloc(size_t bytes, void** d_ptr, unsigned char* handle64);
int ggnn_b200_ipc_open(const unsigned char* handle64, void** d_ptr);
int ggnn_b200_ipc_close(void* d_ptr);
int ggnn_b200_ipc_free(void* d_ptr);
int ggnn_b200_peer_enable(int peer_device);
int ggnn_b200_wait_flag(const uint32_t* d_flag, uint32_t expected, uint32_t timeout_ms, uint32_t* d_timed_out,
                        ggnn_b200_stream_t stream);

/* d_dst[n][(D/32) * t + c] = d_src[n][32 * c + t] for t < 32, c < D/32 (D a multiple of 32, D <= 128): the layout of
 * ggnn_b200_query_params.d_base_interleaved */
int ggnn_b200_interleave_rows(const float* d_src, float* d_dst, uint32_t N, uint32_t D, ggnn_b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * uint8 base / query vectors (the reference's BaseT = uint8_t instantiation, include/ggnn/base/lib.h:26-28).
 * Every distance of the reference is computed on static_cast<ValueT>(value) (include/ggnn/cuda_utils/
 * distance.cuh:104-148), so widening the vectors to fp32 once on the device and running the fp32 kernels gives
 * bit-identical results.  d_dst[i] = (float)d_src[i] for i < count.
 * ---------------------------------------------------------------------------------------------- */
int ggnn_b200_widen_u8(const uint8_t* d_src, float* d_dst, size_t count, ggnn_b200_stream_t stream);

/* Brute-force query of uint8 vectors WITHOUT widening them: the same contract as ggnn_b200_bf_query (it replaces the
 * BaseT = uint8_t instantiation of src/ggnn/query/bf_query_layer.cu:39-65), but d_base / d_query point to uint8 rows of D
 * bytes and d_workspace (>= ggnn_b200_bf_query_u8_workspace_bytes()) is required.  The distances of uint8 vectors are
 * integers below 2^24 for D <= 128, so  |b|^2 - 2 q.b + |q|^2  is evaluated EXACTLY on the int8 tensor cores
 * (tcgen05.mma.kind::i8, u8 x u8 -> s32) and the K smallest (distance, index) pairs are bit-identical to the reference's
 * fp32 arithmetic on the widened values.  Covered: Euclidean, D in {32, 64, 96, 128}, KQuery <= 128, N_base >= 128;
 * _workspace_bytes returns 0 and the call GGNN_B200_ERR_UNSUPPORTED otherwise (widen the rows with ggnn_b200_widen_u8 and
 * call ggnn_b200_bf_query: identical results). */
size_t ggnn_b200_bf_query_u8_workspace_bytes(uint32_t D, int32_t measure, uint32_t KQuery, uint32_t N_base, uint32_t N_query);
int ggnn_b200_bf_query_u8(const ggnn_b200_bf_query_params* p, uint32_t N_query, ggnn_b200_stream_t stream);

/* Diagnostics of the path above (tools/bf_i8_check.py): _pack writes the tile-major, 128-byte-swizzled operand image of
 * n_rows_pad (multiple of 128) rows -- 16 KB per 128-row tile, rows >= n_rows zero -- and the integer row norms;
 * _mma multiplies ONE packed 128-row tile by another with `ksteps` (= D / 32) UMMA instructions and returns the
 * 128 x 128 int32 products d_out[i][j] = sum_k a[i][k] * b[j][k]. */
int ggnn_b200_debug_i8_pack(const uint8_t* d_rows, uint32_t n_rows, uint32_t n_rows_pad, uint32_t D, uint8_t* d_tiled,
                            int32_t* d_norms, ggnn_b200_stream_t stream);
int ggnn_b200_debug_i8_mma(const uint8_t* d_a_tile, const uint8_t* d_b_tile, uint32_t ksteps, int32_t* d_out,
                           ggnn_b200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
