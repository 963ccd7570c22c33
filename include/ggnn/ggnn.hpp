// ggnn.hpp -- header-only C++20 host API over the C ABI (ggnn_b200.h), mirroring the reference's public class
// ggnn::GGNN<KeyT, ValueT> (include/ggnn/base/ggnn.cuh:41-182) and its value types Dataset<T> / Results
// (include/ggnn/base/dataset.cuh:93-166), so that programs written against the reference -- e.g.
// examples/cpp-and-cuda/ggnn_main.cpp -- compile against this header and link libggnn_b200.so + libcudart.
//
// Same method names, argument meaning, defaults and error behaviour (std::runtime_error / std::out_of_range for
// API misuse).  All computation happens in the sm_100a kernels behind the C ABI; there is no CPU fallback.
// Differences (documented in DESIGN.md): datasets are fp32 / int32 only (uint8 base vectors are a "next" row),
// shards stay resident in HBM (no swap to host / disk), multi-GPU results are merged on GPU 0 by a kernel
// instead of the reference's CPU heap merge.
#pragma once

#include <ggnn_b200.h>

#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <limits>
#include <memory>
#include <span>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace ggnn {

enum class DistanceMeasure : int { Euclidean = 0, Cosine = 1 };  // include/ggnn/base/def.h:27-30

enum class DataLocation : uint16_t { UNKNOWN, GPU, MANAGED, CPU_PINNED, CPU_MALLOC, FOREIGN_GPU, FOREIGN_CPU };  // data.cuh:36-44

namespace detail {
inline void cuda_check(cudaError_t e, const char* what)
{
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
inline void abi_check(int rc)
{
  if (rc == 0) return;
  const std::string msg = ggnn_b200_last_error();
  if (rc == GGNN_B200_ERR_INVALID) throw std::out_of_range(msg);
  throw std::runtime_error(msg);
}
struct DeviceGuard {
  int prev{0};
  explicit DeviceGuard(int dev)
  {
    cudaGetDevice(&prev);
    cuda_check(cudaSetDevice(dev), "cudaSetDevice");
  }
  ~DeviceGuard() { cudaSetDevice(prev); }
};
}  // namespace detail

/// 2-D row-major buffer, owning or referencing, on host or device (dataset.cuh:93-160)
template <typename T>
struct Dataset {
  uint64_t N{0};
  uint32_t D{0};
  DataLocation location{DataLocation::UNKNOWN};
  int32_t gpu_id{-1};

  Dataset() = default;
  Dataset(const Dataset&) = delete;
  Dataset& operator=(const Dataset&) = delete;
  Dataset(Dataset&& o) noexcept { *this = std::move(o); }
  Dataset& operator=(Dataset&& o) noexcept
  {
    if (this != &o) {
      release();
      N = o.N; D = o.D; location = o.location; gpu_id = o.gpu_id; mem = o.mem;
      o.mem = nullptr; o.N = 0; o.location = DataLocation::UNKNOWN;
    }
    return *this;
  }
  ~Dataset() { release(); }

  T* data() { return mem; }
  const T* data() const { return mem; }
  size_t numel() const { return static_cast<size_t>(N) * D; }
  size_t size() const { return numel(); }
  size_t size_bytes() const { return numel() * sizeof(T); }
  bool isCPUAccessible() const
  {
    return location == DataLocation::CPU_MALLOC || location == DataLocation::CPU_PINNED ||
           location == DataLocation::FOREIGN_CPU || location == DataLocation::MANAGED;
  }
  bool isGPUAccessible() const
  {
    return location == DataLocation::GPU || location == DataLocation::FOREIGN_GPU || location == DataLocation::MANAGED;
  }
  T& operator[](size_t i) { return mem[i]; }
  const T& operator[](size_t i) const { return mem[i]; }
  T& at(size_t i)
  {
    if (i >= numel()) throw std::out_of_range("Index " + std::to_string(i) + " is out of bounds (size " + std::to_string(numel()) + ").");
    return mem[i];
  }
  const T& at(size_t i) const { return const_cast<Dataset*>(this)->at(i); }
  operator T*() { return mem; }
  operator const T*() const { return mem; }

  static Dataset empty(uint64_t N, uint32_t D, bool pin_memory = false)
  {
    Dataset d;
    d.N = N; d.D = D;
    if (pin_memory) {
      detail::cuda_check(cudaMallocHost(reinterpret_cast<void**>(&d.mem), d.size_bytes()), "cudaMallocHost");
      d.location = DataLocation::CPU_PINNED;
    }
    else {
      d.mem = static_cast<T*>(std::malloc(std::max<size_t>(1, d.size_bytes())));
      if (!d.mem) throw std::bad_alloc();
      d.location = DataLocation::CPU_MALLOC;
    }
    return d;
  }
  static Dataset emptyOnGPU(uint64_t N, uint32_t D, int32_t gpu_id)
  {
    Dataset d;
    d.N = N; d.D = D; d.gpu_id = gpu_id;
    detail::DeviceGuard g(gpu_id);
    detail::cuda_check(cudaMalloc(reinterpret_cast<void**>(&d.mem), std::max<size_t>(16, d.size_bytes())), "cudaMalloc");
    d.location = DataLocation::GPU;
    return d;
  }
  static Dataset copy(const std::span<const T>& data, uint32_t D, bool pin_memory = false)
  {
    if (D == 0 || data.size() % D) throw std::invalid_argument("data size is not a multiple of D");
    Dataset d = empty(data.size() / D, D, pin_memory);
    std::memcpy(d.mem, data.data(), d.size_bytes());
    return d;
  }
  static Dataset referenceCPUData(T* data, uint64_t N, uint32_t D)
  {
    Dataset d;
    d.N = N; d.D = D; d.mem = data; d.location = DataLocation::FOREIGN_CPU;
    return d;
  }
  static Dataset referenceGPUData(T* data, uint64_t N, uint32_t D, int32_t gpu_id)
  {
    Dataset d;
    d.N = N; d.D = D; d.mem = data; d.gpu_id = gpu_id; d.location = DataLocation::FOREIGN_GPU;
    return d;
  }
  /// fvecs / ivecs: [int32 D][D values] per row (src/ggnn/base/dataset.cu:118-233)
  static Dataset load(const std::filesystem::path& file, uint32_t from = 0,
                      uint32_t num = std::numeric_limits<uint32_t>::max(), bool pin_memory = false)
  {
    std::ifstream f(file, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + file.string());
    int32_t dim = 0;
    f.read(reinterpret_cast<char*>(&dim), 4);
    const size_t rec = 4 + static_cast<size_t>(dim) * sizeof(T);
    const size_t total = std::filesystem::file_size(file) / rec;
    const size_t lo = std::min<size_t>(from, total), hi = std::min<size_t>(total, static_cast<size_t>(from) + num);
    Dataset d = empty(hi - lo, static_cast<uint32_t>(dim), pin_memory);
    for (size_t r = lo; r < hi; ++r) {
      f.seekg(static_cast<std::streamoff>(r * rec + 4));
      f.read(reinterpret_cast<char*>(d.mem + (r - lo) * dim), static_cast<std::streamsize>(dim * sizeof(T)));
    }
    return d;
  }
  void store(const std::filesystem::path& file) const
  {
    if (!isCPUAccessible()) throw std::runtime_error("store() needs CPU-accessible data");
    std::ofstream f(file, std::ios::binary);
    const int32_t dim = static_cast<int32_t>(D);
    for (uint64_t r = 0; r < N; ++r) {
      f.write(reinterpret_cast<const char*>(&dim), 4);
      f.write(reinterpret_cast<const char*>(mem + r * D), static_cast<std::streamsize>(D * sizeof(T)));
    }
  }
  void copyTo(Dataset& other, cudaStream_t stream = nullptr) const
  {
    if (other.numel() != numel()) throw std::invalid_argument("copyTo: size mismatch");
    detail::cuda_check(cudaMemcpyAsync(other.mem, mem, size_bytes(), cudaMemcpyDefault, stream), "cudaMemcpyAsync");
    if (!isGPUAccessible() || !other.isGPUAccessible()) detail::cuda_check(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  }
  Dataset clone(cudaStream_t stream = nullptr) const
  {
    Dataset d = isGPUAccessible() && !isCPUAccessible() ? emptyOnGPU(N, D, gpu_id) : empty(N, D, location == DataLocation::CPU_PINNED);
    copyTo(d, stream);
    return d;
  }

 private:
  T* mem{nullptr};
  void release()
  {
    if (!mem) return;
    switch (location) {
      case DataLocation::GPU:
      case DataLocation::MANAGED: cudaFree(mem); break;
      case DataLocation::CPU_PINNED: cudaFreeHost(mem); break;
      case DataLocation::CPU_MALLOC: std::free(mem); break;
      default: break;  // FOREIGN_*: never freed (data.cu:147-149)
    }
    mem = nullptr;
  }
};

using GenericDataset = Dataset<float>;  // only float base / query vectors are built here

template <typename KeyT, typename ValueT>
struct Results {
  Dataset<KeyT> ids{};
  Dataset<ValueT> dists{};
};

/// device-side view of one shard's graph blob (include/ggnn/base/graph.h:36-72)
template <typename KeyT, typename ValueT>
struct Graph {
  ggnn_b200_graph_config config{};
  ggnn_b200_graph_offsets offsets{};
  Dataset<uint8_t> memory{};
  const KeyT* graph() const { return reinterpret_cast<const KeyT*>(memory.data() + offsets.graph); }
  const KeyT* translation() const { return reinterpret_cast<const KeyT*>(memory.data() + offsets.translation); }
  const KeyT* selection() const { return reinterpret_cast<const KeyT*>(memory.data() + offsets.selection); }
  const ValueT* nn1_stats() const { return reinterpret_cast<const ValueT*>(memory.data() + offsets.nn1_stats); }
};

template <typename KeyT = int32_t, typename ValueT = float>
class GGNN {
  static_assert(std::is_same_v<KeyT, int32_t> && std::is_same_v<ValueT, float>, "GGNN<int32_t, float> is the instantiated type (lib.h:23-28)");

 public:
  using Results = ggnn::Results<KeyT, ValueT>;
  using Graph = ggnn::Graph<KeyT, ValueT>;
  static constexpr uint32_t MIN_D = 1, MAX_D = 4096, MIN_KBUILD = 2, MAX_KBUILD = 512;

  GGNN() = default;
  ~GGNN()
  {
    for (auto& sh : shards) {
      cudaSetDevice(sh.gpu);
      if (sh.stream) cudaStreamDestroy(sh.stream);
      if (sh.work_counter) cudaFree(sh.work_counter);
    }
  }
  GGNN(const GGNN&) = delete;
  GGNN& operator=(const GGNN&) = delete;
  GGNN(GGNN&&) noexcept = default;
  GGNN& operator=(GGNN&&) noexcept = default;

  void setWorkingDirectory(const std::filesystem::path& dir) { graph_dir = dir; }
  void setCPUMemoryLimit(size_t) {}      // shards stay resident in HBM
  void setReservedGPUMemory(size_t) {}
  void setGPUs(const std::span<const int>& ids)
  {
    if (!shards.empty()) throw std::runtime_error("GPUs cannot be changed after the graph has been set up.");
    if (ids.empty()) throw std::out_of_range("at least one GPU is required");
    gpu_ids.assign(ids.begin(), ids.end());
  }
  void setGPUs(const std::vector<int>& ids) { setGPUs(std::span<const int>{ids.data(), ids.size()}); }
  void setShardSize(uint32_t n)
  {
    if (!shards.empty()) throw std::runtime_error("The shard size cannot be changed after the graph has been set up.");
    N_shard = n;
  }
  void setReturnResultsOnGPU(bool flag = true) { return_results_on_gpu = flag; }

  void setBase(GenericDataset&& b)
  {
    owned_base = std::move(b);
    setBaseReference(owned_base);
  }
  void setBaseReference(const GenericDataset& b)
  {
    if (!shards.empty()) throw std::runtime_error("The base cannot be changed after the graph has been set up.");
    if (b.D < MIN_D || b.D > MAX_D) throw std::out_of_range("unsupported dimension");
    base = &b;
  }
  void setBaseReference(GenericDataset&&) = delete;

  void build(uint32_t KBuild, float tau_build, uint32_t refinement_iterations = 2, DistanceMeasure measure = DistanceMeasure::Euclidean)
  {
    prepare(KBuild);
    for (auto& sh : shards) {
      detail::DeviceGuard g(sh.gpu);
      const size_t scratch_bytes = ggnn_b200_build_scratch_bytes(&cfg);
      void* scratch = nullptr;
      detail::cuda_check(cudaMalloc(&scratch, scratch_bytes), "cudaMalloc(build scratch)");
      detail::cuda_check(cudaMemsetAsync(sh.graph.memory.data(), 0, sh.graph.memory.size_bytes(), sh.stream), "cudaMemsetAsync");
      const int rc = ggnn_b200_build_graph(&cfg, sh.base.data(), static_cast<int>(measure), tau_build, refinement_iterations,
                                           nullptr, sh.graph.memory.data(), scratch, scratch_bytes, sh.stream);
      cudaStreamSynchronize(sh.stream);
      cudaFree(scratch);
      detail::abi_check(rc);
    }
  }

  void store()
  {
    if (shards.empty()) throw std::runtime_error("There is no graph to store.");
    std::filesystem::create_directories(graph_dir);
    for (auto& sh : shards) {  // gpu_instance.cu:86-115: part_<global_shard_id>.ggnn = raw blob
      std::vector<uint8_t> h(sh.graph.memory.size_bytes());
      detail::DeviceGuard g(sh.gpu);
      detail::cuda_check(cudaMemcpy(h.data(), sh.graph.memory.data(), h.size(), cudaMemcpyDeviceToHost), "cudaMemcpy");
      std::ofstream f(graph_dir / ("part_" + std::to_string(sh.global_id) + ".ggnn"), std::ios::binary);
      f.write(reinterpret_cast<const char*>(h.data()), static_cast<std::streamsize>(h.size()));
    }
  }

  void load(uint32_t KBuild)
  {
    prepare(KBuild);
    for (auto& sh : shards) {
      const auto path = graph_dir / ("part_" + std::to_string(sh.global_id) + ".ggnn");
      if (!std::filesystem::exists(path) || std::filesystem::file_size(path) != sh.graph.memory.size_bytes())
        throw std::runtime_error(path.string() + ": missing or unexpected file size");
      std::vector<uint8_t> h(sh.graph.memory.size_bytes());
      std::ifstream f(path, std::ios::binary);
      f.read(reinterpret_cast<char*>(h.data()), static_cast<std::streamsize>(h.size()));
      detail::DeviceGuard g(sh.gpu);
      detail::cuda_check(cudaMemcpy(sh.graph.memory.data(), h.data(), h.size(), cudaMemcpyHostToDevice), "cudaMemcpy");
    }
  }

  [[nodiscard]] Results query(const GenericDataset& query, uint32_t KQuery, float tau_query, uint32_t max_iterations = 400,
                              DistanceMeasure measure = DistanceMeasure::Euclidean)
  {
    if (shards.empty()) throw std::runtime_error("There is no graph to query.");
    if (query.D != base->D) throw std::out_of_range("query dimension does not match the base");
    const uint32_t n_gpus = static_cast<uint32_t>(gpu_ids.size());
    if (return_results_on_gpu && n_gpus > 1)
      throw std::runtime_error("Returning query results on GPU is only possible when using a single GPU.");
    const uint32_t Nq = static_cast<uint32_t>(query.N);
    std::vector<Dataset<KeyT>> ids(n_gpus);
    std::vector<Dataset<ValueT>> dists(n_gpus);
    std::vector<Dataset<float>> q_dev(n_gpus);
    std::vector<Dataset<KeyT>> merged_i(n_gpus);
    std::vector<Dataset<ValueT>> merged_d(n_gpus);
    // launch everything asynchronously on every GPU first
    for (uint32_t gi = 0; gi < n_gpus; ++gi) {
      const int gpu = gpu_ids[gi];
      detail::DeviceGuard g(gpu);
      cudaStream_t stream = shards[gi * spg].stream;
      const float* dq = query.data();
      if (!(query.isGPUAccessible() && query.gpu_id == gpu)) {
        q_dev[gi] = Dataset<float>::emptyOnGPU(query.N, query.D, gpu);
        detail::cuda_check(cudaMemcpyAsync(q_dev[gi].data(), query.data(), query.size_bytes(), cudaMemcpyDefault, stream), "cudaMemcpyAsync(query)");
        dq = q_dev[gi].data();
      }
      ids[gi] = Dataset<KeyT>::emptyOnGPU(Nq, KQuery * spg, gpu);
      dists[gi] = Dataset<ValueT>::emptyOnGPU(Nq, KQuery * spg, gpu);
      for (uint32_t s = 0; s < spg; ++s) {
        Shard& sh = shards[gi * spg + s];
        ggnn_b200_query_params p{};
        p.D = cfg.D; p.measure = static_cast<int>(measure); p.KQuery = KQuery;
        p.tau_query = tau_query; p.max_iterations = max_iterations;
        p.N_base = static_cast<int32_t>(cfg.N); p.KBuild = cfg.KBuild; p.num_starting_points = cfg.S;
        p.d_base = sh.base.data(); p.d_query = dq;
        p.d_graph = sh.graph.graph();
        p.d_starting_points = sh.graph.translation() + cfg.STs_offsets[GGNN_B200_L - 1];
        p.d_nn1_stats = sh.graph.nn1_stats();
        p.d_query_results = ids[gi].data(); p.d_query_results_dists = dists[gi].data();
        p.shards_per_gpu = spg; p.on_gpu_shard_id = s;
        p.d_work_counter = sh.work_counter;
        detail::abi_check(ggnn_b200_query(&p, Nq, stream));
      }
      if (spg > 1) {  // replaces gpu_instance.cu:745-790
        merged_i[gi] = Dataset<KeyT>::emptyOnGPU(Nq, KQuery, gpu);
        merged_d[gi] = Dataset<ValueT>::emptyOnGPU(Nq, KQuery, gpu);
        detail::abi_check(ggnn_b200_merge_topk(ids[gi].data(), dists[gi].data(), spg, KQuery, static_cast<size_t>(KQuery) * spg,
                                               KQuery, Nq, KQuery, 0, merged_i[gi].data(), merged_d[gi].data(), stream));
        ids[gi] = std::move(merged_i[gi]);
        dists[gi] = std::move(merged_d[gi]);
      }
    }
    Results out;
    if (n_gpus == 1) {
      detail::DeviceGuard g(gpu_ids[0]);
      detail::cuda_check(cudaStreamSynchronize(shards[0].stream), "cudaStreamSynchronize");
      out.ids = std::move(ids[0]);
      out.dists = std::move(dists[0]);
    }
    else {  // replaces ResultMerger::merge (result_merger.cpp:51-149): peer copies + one merge kernel on GPU 0
      const int g0 = gpu_ids[0];
      Dataset<KeyT> all_i = Dataset<KeyT>::emptyOnGPU(static_cast<uint64_t>(n_gpus) * Nq, KQuery, g0);
      Dataset<ValueT> all_d = Dataset<ValueT>::emptyOnGPU(static_cast<uint64_t>(n_gpus) * Nq, KQuery, g0);
      for (uint32_t gi = 0; gi < n_gpus; ++gi) {
        detail::DeviceGuard g(gpu_ids[gi]);
        cudaStream_t stream = shards[gi * spg].stream;
        const size_t n = static_cast<size_t>(Nq) * KQuery;
        detail::cuda_check(cudaMemcpyPeerAsync(all_i.data() + gi * n, g0, ids[gi].data(), gpu_ids[gi], n * sizeof(KeyT), stream), "cudaMemcpyPeerAsync");
        detail::cuda_check(cudaMemcpyPeerAsync(all_d.data() + gi * n, g0, dists[gi].data(), gpu_ids[gi], n * sizeof(ValueT), stream), "cudaMemcpyPeerAsync");
        detail::cuda_check(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
      }
      detail::DeviceGuard g(g0);
      out.ids = Dataset<KeyT>::emptyOnGPU(Nq, KQuery, g0);
      out.dists = Dataset<ValueT>::emptyOnGPU(Nq, KQuery, g0);
      detail::abi_check(ggnn_b200_merge_topk(all_i.data(), all_d.data(), n_gpus, static_cast<size_t>(Nq) * KQuery, KQuery, KQuery, Nq,
                                             KQuery, static_cast<int64_t>(spg) * cfg.N, out.ids.data(), out.dists.data(), shards[0].stream));
      detail::cuda_check(cudaStreamSynchronize(shards[0].stream), "cudaStreamSynchronize");
    }
    return return_results_on_gpu ? std::move(out) : to_host(std::move(out));
  }

  [[nodiscard]] Results bfQuery(const GenericDataset& query, uint32_t KGT = 100, DistanceMeasure measure = DistanceMeasure::Euclidean)
  {
    if (!base) throw std::runtime_error("The base needs to be set before running a brute-force query.");
    if (gpu_ids.size() > 1) throw std::runtime_error("bfQuery supports only a single GPU.");  // ggnn.cu:338-339
    const int gpu = gpu_ids[0];
    detail::DeviceGuard g(gpu);
    Dataset<float> b_dev, q_dev;
    const float* db = base->data();
    if (!shards.empty() && shards.size() == 1) db = shards[0].base.data();
    else if (!(base->isGPUAccessible() && base->gpu_id == gpu)) {
      b_dev = Dataset<float>::emptyOnGPU(base->N, base->D, gpu);
      detail::cuda_check(cudaMemcpy(b_dev.data(), base->data(), base->size_bytes(), cudaMemcpyDefault), "cudaMemcpy(base)");
      db = b_dev.data();
    }
    const float* dq = query.data();
    if (!(query.isGPUAccessible() && query.gpu_id == gpu)) {
      q_dev = Dataset<float>::emptyOnGPU(query.N, query.D, gpu);
      detail::cuda_check(cudaMemcpy(q_dev.data(), query.data(), query.size_bytes(), cudaMemcpyDefault), "cudaMemcpy(query)");
      dq = q_dev.data();
    }
    Results out;
    out.ids = Dataset<KeyT>::emptyOnGPU(query.N, KGT, gpu);
    out.dists = Dataset<ValueT>::emptyOnGPU(query.N, KGT, gpu);
    ggnn_b200_bf_query_params p{};
    p.D = base->D; p.measure = static_cast<int>(measure); p.KQuery = KGT; p.N_base = static_cast<int32_t>(base->N);
    p.d_base = db; p.d_query = dq; p.d_query_results = out.ids.data(); p.d_query_results_dists = out.dists.data();
    p.workspace_bytes = ggnn_b200_bf_query_workspace_bytes(p.D, p.measure, KGT, static_cast<uint32_t>(base->N), static_cast<uint32_t>(query.N));
    if (p.workspace_bytes) detail::cuda_check(cudaMalloc(&p.d_workspace, p.workspace_bytes), "cudaMalloc(bf workspace)");
    const int rc = ggnn_b200_bf_query(&p, static_cast<uint32_t>(query.N), nullptr);
    cudaDeviceSynchronize();
    if (p.d_workspace) cudaFree(p.d_workspace);
    detail::abi_check(rc);
    return return_results_on_gpu ? std::move(out) : to_host(std::move(out));
  }

  [[nodiscard]] const Graph& getGraph(uint32_t global_shard_id = 0)
  {
    if (global_shard_id >= shards.size()) throw std::out_of_range("no such shard");
    return shards[global_shard_id].graph;
  }

 private:
  struct Shard {
    int gpu{0};
    uint32_t global_id{0};
    Dataset<float> base;
    Graph graph;
    cudaStream_t stream{nullptr};
    uint32_t* work_counter{nullptr};
  };

  static Results to_host(Results r)
  {
    Results h;
    h.ids = Dataset<KeyT>::empty(r.ids.N, r.ids.D, true);
    h.dists = Dataset<ValueT>::empty(r.dists.N, r.dists.D, true);
    detail::cuda_check(cudaMemcpy(h.ids.data(), r.ids.data(), r.ids.size_bytes(), cudaMemcpyDeviceToHost), "cudaMemcpy(ids)");
    detail::cuda_check(cudaMemcpy(h.dists.data(), r.dists.data(), r.dists.size_bytes(), cudaMemcpyDeviceToHost), "cudaMemcpy(dists)");
    return h;
  }

  // src/ggnn/base/ggnn.cu:154-203
  void prepare(uint32_t KBuild)
  {
    if (!base || !base->data()) throw std::runtime_error("The base needs to be set before building a graph.");
    if (KBuild < MIN_KBUILD || KBuild > MAX_KBUILD) throw std::out_of_range("KBuild out of range");
    if (!shards.empty()) {
      if (cfg.KBuild != KBuild) throw std::runtime_error("graph already set up with a different KBuild");
      return;
    }
    const uint64_t N = base->N;
    const uint64_t n_shard = N_shard ? N_shard : N;
    if (N % n_shard) throw std::out_of_range("The base size needs to be divisible by the shard size.");
    const uint64_t num_shards = N / n_shard;
    if (num_shards % gpu_ids.size()) throw std::out_of_range("The number of shards needs to be divisible by the number of GPUs.");
    spg = static_cast<uint32_t>(num_shards / gpu_ids.size());
    detail::abi_check(ggnn_b200_graph_config_init(&cfg, static_cast<uint32_t>(n_shard), base->D, KBuild));
    ggnn_b200_graph_offsets off;
    ggnn_b200_graph_blob_offsets(&cfg, &off);
    shards.resize(num_shards);
    for (uint32_t gi = 0; gi < gpu_ids.size(); ++gi) {
      detail::DeviceGuard g(gpu_ids[gi]);
      for (uint32_t s = 0; s < spg; ++s) {
        Shard& sh = shards[gi * spg + s];
        sh.gpu = gpu_ids[gi];
        sh.global_id = gi * spg + s;
        detail::cuda_check(cudaStreamCreate(&sh.stream), "cudaStreamCreate");
        detail::cuda_check(cudaMalloc(reinterpret_cast<void**>(&sh.work_counter), 16), "cudaMalloc");
        sh.base = Dataset<float>::emptyOnGPU(n_shard, base->D, sh.gpu);
        detail::cuda_check(cudaMemcpyAsync(sh.base.data(), base->data() + static_cast<size_t>(sh.global_id) * n_shard * base->D,
                                           sh.base.size_bytes(), cudaMemcpyDefault, sh.stream), "cudaMemcpyAsync(base shard)");
        sh.graph.config = cfg;
        sh.graph.offsets = off;
        sh.graph.memory = Dataset<uint8_t>::emptyOnGPU(off.total, 1, sh.gpu);
        detail::cuda_check(cudaStreamSynchronize(sh.stream), "cudaStreamSynchronize");
      }
    }
  }

  std::filesystem::path graph_dir{"."};
  std::vector<int> gpu_ids{0};
  uint32_t N_shard{0}, spg{1};
  bool return_results_on_gpu{false};
  GenericDataset owned_base{};
  const GenericDataset* base{nullptr};
  ggnn_b200_graph_config cfg{};
  std::vector<Shard> shards;
};

}  // namespace ggnn
