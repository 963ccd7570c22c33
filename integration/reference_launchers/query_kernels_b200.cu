// query_kernels_b200.cu -- INTEGRATION.md section 1, compiled: a replacement for the reference's
// src/ggnn/query/query_kernels.cu (the thin host -> CUDA boundary of the query path).  It defines the same
// QueryKernelsImpl class behind the reference's own QueryKernels interface (include/ggnn/query/query_kernels.cuh:33-66),
// but its two launchers call libggnn_b200.so through the C ABI (include/ggnn_b200.h) instead of launching the reference's
// kernels.  Everything above it -- ggnn::GGNN, GPUInstance (sharding, swapping, result sort), datasets, file IO -- is the
// UNMODIFIED reference, compiled from /root/reference by oracle/build_ref.sh into oracle/_ref/libggnn_ref_hybrid.so.
// tests/test_gpu_parity.py::test_reference_with_our_launchers_is_bit_identical runs the reference's driver against both
// libraries.  This file is written for this repo; it contains no reference code beyond the interface it implements.
#include <ggnn/query/query_kernels.cuh>

#include <ggnn/base/def.h>
#include <ggnn/base/graph_config.h>
#include <ggnn/base/lib.h>
#include <ggnn/base/dataset.cuh>
#include <ggnn/base/gpu_instance.cuh>

#include <glog/logging.h>

#include <ggnn_b200.h>

#include <cstdint>
#include <type_traits>

namespace ggnn {

template <typename KeyT, typename ValueT, typename BaseT>
class QueryKernelsImpl : public QueryKernels<KeyT, ValueT, BaseT> {
 public:
  using GPUInstance = ggnn::GPUInstance<KeyT, ValueT, BaseT>;
  using Results = ggnn::Results<KeyT, ValueT>;

  QueryKernelsImpl(const DistanceMeasure measure) : measure{measure} {}

 private:
  const DistanceMeasure measure;

  static const float* as_float(const BaseT* p)
  {
    if constexpr (std::is_same_v<BaseT, float>) return p;
    else {
      LOG(FATAL) << "this integration example wires up float base vectors only";
      return nullptr;
    }
  }

  // replaces src/ggnn/query/query_kernels.cu:50-186
  void query(const GPUInstance& gpu_instance, const uint32_t shard_id, const Dataset<BaseT>& query, const uint32_t KQuery,
             const uint32_t max_iters, const float tau_query, Results& results) override
  {
    gpu_instance.gpu_ctx.activate();
    const uint32_t on_gpu_shard_id = shard_id - gpu_instance.shard_config.num_shards * gpu_instance.shard_config.device_index;
    const auto& gpu_buffer = gpu_instance.getGPUGraphBuffer(on_gpu_shard_id);
    const auto& base = gpu_instance.getGPUBaseBuffer(on_gpu_shard_id).base;
    const auto& graph = gpu_buffer.graph;
    ggnn_b200_query_params p{};  // cache / sorted / block sizes = 0: derived like query_kernels.cu:77-110
    p.D = query.D;
    p.measure = static_cast<int>(measure);
    p.KQuery = KQuery;
    p.tau_query = tau_query;
    p.max_iterations = max_iters;
    p.N_base = static_cast<int32_t>(base.N);
    p.KBuild = gpu_instance.graph_config.KBuild;
    p.num_starting_points = gpu_instance.graph_config.S;
    p.d_base = as_float(base.data());
    p.d_query = as_float(query.data());
    p.d_graph = graph.graph[0].data();
    p.d_starting_points = graph.translation[GraphConfig::L - 1].data();
    p.d_nn1_stats = graph.nn1_stats.data();
    p.d_query_results = results.ids.data();
    p.d_query_results_dists = results.dists.data();
    p.shards_per_gpu = gpu_instance.shard_config.num_shards;
    p.on_gpu_shard_id = on_gpu_shard_id;
    CHECK_EQ(ggnn_b200_query(&p, static_cast<uint32_t>(query.N), gpu_buffer.stream.get()), 0) << ggnn_b200_last_error();
  }

  // replaces src/ggnn/query/query_kernels.cu:188-264 (exact SIMT scan: the reference hands over no scratch buffer)
  void bruteForceQuery(const Dataset<BaseT>& base, const Dataset<BaseT>& query, const uint32_t KQuery, Results& results,
                       cudaStream_t stream) override
  {
    ggnn_b200_bf_query_params p{};
    p.D = query.D;
    p.measure = static_cast<int>(measure);
    p.KQuery = KQuery;
    p.N_base = static_cast<int32_t>(base.N);
    p.d_base = as_float(base.data());
    p.d_query = as_float(query.data());
    p.d_query_results = results.ids.data();
    p.d_query_results_dists = results.dists.data();
    CHECK_EQ(ggnn_b200_bf_query(&p, static_cast<uint32_t>(query.N), stream), 0) << ggnn_b200_last_error();
  }
};

template <typename KeyT, typename ValueT, typename BaseT>
QueryKernels<KeyT, ValueT, BaseT>::QueryKernels(const DistanceMeasure measure)
{
  pimpl.reset(new QueryKernelsImpl<KeyT, ValueT, BaseT>{measure});
}

GGNN_EVAL(GGNN_KEYS, GGNN_VALUES, GGNN_BASES, GGNN_INSTANTIATE_CLASS, QueryKernels);
GGNN_EVAL(GGNN_KEYS, GGNN_VALUES, GGNN_BASES, GGNN_INSTANTIATE_CLASS, QueryKernelsImpl);
};  // namespace ggnn
