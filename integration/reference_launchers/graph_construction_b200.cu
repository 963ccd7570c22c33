// graph_construction_b200.cu -- INTEGRATION.md section 1, compiled: a replacement for the reference's
// src/ggnn/construction/graph_construction.cu.  Same GraphConstructionImpl class behind the reference's own
// GraphConstruction interface (include/ggnn/construction/graph_construction.cuh:34-59); build() / refine() call
// libggnn_b200.so (ggnn_b200_build_graph / ggnn_b200_refine_graph) on the reference's own graph blob -- the layouts are
// byte-identical (include/ggnn/base/graph.h:38-72).  Like the reference it keeps ONE cuRAND generator per GPU instance
// whose sequence continues over the shards (graph_construction.cu:96-102,127).  Written for this repo.
#include <ggnn/construction/graph_construction.cuh>

#include <ggnn/base/def.h>
#include <ggnn/base/graph.h>
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
class GraphConstructionImpl : public GraphConstruction<KeyT, ValueT, BaseT> {
 public:
  using Graph = ggnn::Graph<KeyT, ValueT>;
  using GPUInstance = ggnn::GPUInstance<KeyT, ValueT, BaseT>;

  GraphConstructionImpl(GPUInstance& gpu_instance, const float tau_build, const DistanceMeasure measure)
      : gpu_instance{gpu_instance}, tau_build{tau_build}, measure{measure}
  {
    const GraphConfig& gc = gpu_instance.graph_config;
    CHECK_EQ(ggnn_b200_graph_config_init(&cfg, gc.N, gc.D, gc.KBuild), 0) << ggnn_b200_last_error();
    scratch_bytes = ggnn_b200_build_scratch_bytes(&cfg);
    gpu_instance.gpu_ctx.activate();
    scratch = Dataset<std::byte>::emptyOnGPU(scratch_bytes, 1, gpu_instance.gpu_ctx.gpu_id);
    uniforms = Dataset<float>::emptyOnGPU(static_cast<uint64_t>(cfg.Ns[0]) + cfg.Ns[1] + cfg.Ns[2], 1, gpu_instance.gpu_ctx.gpu_id);
    CHECK_EQ(ggnn_b200_rng_create(&rng, 1234ULL), 0) << ggnn_b200_last_error();
  }
  ~GraphConstructionImpl() override { ggnn_b200_rng_destroy(rng); }

 private:
  GPUInstance& gpu_instance;
  float tau_build{};
  DistanceMeasure measure;
  ggnn_b200_graph_config cfg{};
  size_t scratch_bytes{0};
  Dataset<std::byte> scratch{};
  Dataset<float> uniforms{};
  ggnn_b200_rng* rng{nullptr};

  static const float* as_float(const BaseT* p)
  {
    if constexpr (std::is_same_v<BaseT, float>) return p;
    else {
      LOG(FATAL) << "this integration example wires up float base vectors only";
      return nullptr;
    }
  }

  // replaces graph_construction.cu:128-140 (and everything it launches)
  void build(Graph& graph, const Dataset<BaseT>& base, const cudaStream_t stream) override
  {
    CHECK_EQ(ggnn_b200_rng_fill_build(rng, &cfg, uniforms.data(), stream), 0) << ggnn_b200_last_error();
    CHECK_EQ(ggnn_b200_build_graph(&cfg, as_float(base.data()), static_cast<int>(measure), tau_build, 0, uniforms.data(),
                                   graph.memory.data(), scratch.data(), scratch_bytes, stream), 0)
        << ggnn_b200_last_error();
  }
  // replaces graph_construction.cu:141-147
  void refine(Graph& graph, const Dataset<BaseT>& base, const cudaStream_t stream) override
  {
    CHECK_EQ(ggnn_b200_refine_graph(&cfg, as_float(base.data()), static_cast<int>(measure), tau_build, graph.memory.data(),
                                    scratch.data(), scratch_bytes, stream), 0)
        << ggnn_b200_last_error();
  }
};

template <typename KeyT, typename ValueT, typename BaseT>
GraphConstruction<KeyT, ValueT, BaseT>::GraphConstruction(GPUInstance& gpu_instance, const float tau_build,
                                                          const DistanceMeasure measure)
{
  pimpl.reset(new GraphConstructionImpl<KeyT, ValueT, BaseT>{gpu_instance, tau_build, measure});
}

GGNN_EVAL(GGNN_KEYS, GGNN_VALUES, GGNN_BASES, GGNN_INSTANTIATE_CLASS, GraphConstruction);
GGNN_EVAL(GGNN_KEYS, GGNN_VALUES, GGNN_BASES, GGNN_INSTANTIATE_CLASS, GraphConstructionImpl);

};  // namespace ggnn
