// ggnn.hpp -- header-only C++20 host API over the C ABI (ggnn_b200.h), mirroring the reference's public
// surface: ggnn::GGNN<KeyT, ValueT> (include/ggnn/base/ggnn.cuh:41-182), the value types GenericDataset /
// Dataset<T> / Results (include/ggnn/base/dataset.cuh:38-166, data.cuh:26-149) and Evaluator / Evaluation
// (include/ggnn/base/eval.h:31-65), so that programs written against the reference -- its own
// examples/cpp-and-cuda/{ggnn_main.cpp, ggnn_main_gpu_data.cu, ggnn_main_multi_gpu.cpp, ggnn_benchmark.cpp} --
// compile unchanged against this header tree and link libggnn_b200.so + libcudart.
//
// Same names, argument meaning, defaults and error behaviour (std::runtime_error / std::out_of_range for API
// misuse).  All computation happens in the sm_100a kernels behind the C ABI; there is no CPU fallback.
// Differences: uint8 vectors stay uint8 on the device and are read natively by the traversal kernel where it
// has a variant (D in {32, 64, 96, 128, 256}, KQuery <= 47); the other kernels run on rows widened on the device for the duration
// of the call (bit-identical results, see ggnn_b200_widen_u8); multi-GPU results are merged on the first GPU by a kernel -- the traversal kernels store their
// lists straight into its memory (peer access) -- instead of the reference's CPU heap merge, so results may stay on the
// GPU for any number of GPUs; queryAsync() keeps several host batches in flight.  Shards that do not fit on their GPU
// are swapped GPU <-> pinned host memory <-> part_<id>.ggnn files like the reference does (setCPUMemoryLimit /
// setReservedGPUMemory, src/ggnn/base/gpu_instance.cu:135-227, 370-467).
#pragma once

#include <ggnn_b200.h>

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <limits>
#include <memory>
#include <mutex>
#include <ostream>
#include <span>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace ggnn {

enum class DistanceMeasure : int { Euclidean = 0, Cosine = 1 };  // include/ggnn/base/def.h:27-30

// include/ggnn/base/data.cuh:26-44
enum class DataType : uint16_t { UNKNOWN, BYTE, UINT8, INT32, UINT32, FLOAT };
enum class DataLocation : uint16_t { UNKNOWN, GPU, MANAGED, CPU_PINNED, CPU_MALLOC, FOREIGN_GPU, FOREIGN_CPU };

inline std::ostream& operator<<(std::ostream& os, DataType t)
{
  static constexpr const char* names[] = {"unknown", "byte", "uint8", "int32", "uint32", "float"};
  return os << names[static_cast<size_t>(t)];
}
inline std::ostream& operator<<(std::ostream& os, DataLocation l)
{
  static constexpr const char* names[] = {"unknown", "GPU", "managed", "CPU (pinned)", "CPU", "GPU (foreign)", "CPU (foreign)"};
  return os << names[static_cast<size_t>(l)];
}

namespace detail {
template <typename T> struct TypeTag;
template <> struct TypeTag<std::byte> { static constexpr DataType value = DataType::BYTE; };
template <> struct TypeTag<uint8_t> { static constexpr DataType value = DataType::UINT8; };
template <> struct TypeTag<int32_t> { static constexpr DataType value = DataType::INT32; };
template <> struct TypeTag<uint32_t> { static constexpr DataType value = DataType::UINT32; };
template <> struct TypeTag<float> { static constexpr DataType value = DataType::FLOAT; };

inline size_t dataSize(DataType t)
{
  switch (t) {
    case DataType::BYTE:
    case DataType::UINT8: return 1;
    case DataType::INT32:
    case DataType::UINT32:
    case DataType::FLOAT: return 4;
    default: return 0;
  }
}
inline DataLocation disown(DataLocation l)
{
  switch (l) {
    case DataLocation::GPU:
    case DataLocation::MANAGED: return DataLocation::FOREIGN_GPU;
    case DataLocation::CPU_PINNED:
    case DataLocation::CPU_MALLOC: return DataLocation::FOREIGN_CPU;
    default: return l;
  }
}
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
/// Recycles pinned host buffers (results of query() / bfQuery()): cudaMallocHost / cudaFreeHost cost more than a whole
/// query batch on a B200 (milliseconds: page pinning + mapping into every device), so freed buffers of up to 64 MB are kept
/// (at most 256 MB in total) and handed out again for the same size.
class PinnedPool {
 public:
  static PinnedPool& instance()
  {
    static PinnedPool pool;
    return pool;
  }
  void* get(size_t bytes)
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = free_.find(bytes);
    if (it == free_.end()) return nullptr;
    void* p = it->second;
    free_.erase(it);
    held -= bytes;
    return p;
  }
  bool put(void* p, size_t bytes)
  {
    if (bytes > MAX_BUFFER) return false;
    std::lock_guard<std::mutex> lock(mu);
    if (held + bytes > MAX_HELD) return false;
    free_.emplace(bytes, p);
    held += bytes;
    return true;
  }
  ~PinnedPool()
  {
    for (auto& kv : free_) cudaFreeHost(kv.second);  // (errors at process teardown are of no interest)
  }

 private:
  static constexpr size_t MAX_BUFFER = size_t(64) << 20, MAX_HELD = size_t(256) << 20;
  std::mutex mu;
  std::unordered_multimap<size_t, void*> free_;
  size_t held{0};
};

struct DeviceGuard {
  int prev{0};
  explicit DeviceGuard(int dev)
  {
    cudaGetDevice(&prev);
    cuda_check(cudaSetDevice(dev), "cudaSetDevice");
  }
  ~DeviceGuard() { cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
}  // namespace detail

template <typename T>
constexpr DataType DataType_v = detail::TypeTag<std::remove_const_t<T>>::value;

template <typename T>
struct Dataset;

/// type-erased 2-D row-major buffer, owning or referencing, on host or device (dataset.cuh:38-91)
struct GenericDataset {
  uint64_t N{0};
  uint32_t D{0};
  DataType type{DataType::UNKNOWN};
  DataLocation location{DataLocation::UNKNOWN};
  int32_t gpu_id{-1};

  GenericDataset() = default;
  GenericDataset(const GenericDataset&) = delete;
  GenericDataset& operator=(const GenericDataset&) = delete;
  GenericDataset(GenericDataset&& o) noexcept { take(o); }
  GenericDataset& operator=(GenericDataset&& o) noexcept
  {
    if (this != &o) {
      release();
      take(o);
    }
    return *this;
  }
  virtual ~GenericDataset() { release(); }

  size_t numel() const { return static_cast<size_t>(N) * D; }
  size_t element_size() const { return detail::dataSize(type); }
  size_t required_size_bytes() const { return element_size() * numel(); }
  bool isCPUAccessible() const
  {
    return location == DataLocation::CPU_MALLOC || location == DataLocation::CPU_PINNED ||
           location == DataLocation::FOREIGN_CPU || location == DataLocation::MANAGED;
  }
  bool isGPUAccessible() const
  {
    return location == DataLocation::GPU || location == DataLocation::FOREIGN_GPU || location == DataLocation::MANAGED;
  }
  /// the memory will not be freed by this object any more
  void releaseOwnership() { location = detail::disown(location); }

  /// non-owning alias of the whole buffer / of rows [from, from + num)
  GenericDataset reference() const { return referenceRange(0, N); }
  GenericDataset referenceRange(uint64_t from, uint64_t num) const
  {
    if (from + num > N) throw std::out_of_range("referenceRange: rows out of bounds");
    GenericDataset r;
    r.N = num;
    r.D = D;
    r.type = type;
    r.location = detail::disown(location);
    r.gpu_id = gpu_id;
    r.mem = static_cast<char*>(mem) + from * D * element_size();
    return r;
  }

  template <typename T>
  std::span<T> reinterpret() { return {reinterpret_cast<T*>(mem), numel()}; }
  template <typename T>
  std::span<const T> reinterpret() const { return {reinterpret_cast<const T*>(mem), numel()}; }
  template <typename T>
  std::span<T> access()
  {
    check_type<T>();
    return reinterpret<T>();
  }
  template <typename T>
  std::span<const T> access() const
  {
    check_type<T>();
    return reinterpret<T>();
  }
  explicit operator void*() { return mem; }
  explicit operator const void*() const { return mem; }

  /// .fvecs / .bvecs / .ivecs by file extension (src/ggnn/base/dataset.cu:118-131)
  static GenericDataset load(const std::filesystem::path& file, uint32_t from = 0,
                             uint32_t num = std::numeric_limits<uint32_t>::max(), bool pin_memory = false);

  /// uninitialised owning buffer
  static GenericDataset allocate(uint64_t N, uint32_t D, DataType type, DataLocation where, int32_t gpu_id = -1)
  {
    GenericDataset d;
    d.N = N;
    d.D = D;
    d.type = type;
    d.gpu_id = gpu_id;
    const size_t bytes = std::max<size_t>(16, d.required_size_bytes());
    switch (where) {
      case DataLocation::GPU: {
        detail::DeviceGuard g(gpu_id);
        detail::cuda_check(cudaMalloc(&d.mem, bytes), "cudaMalloc");
        break;
      }
      case DataLocation::CPU_PINNED:
        d.alloc_bytes = (bytes + 255) / 256 * 256;
        d.mem = detail::PinnedPool::instance().get(d.alloc_bytes);
        if (!d.mem) detail::cuda_check(cudaMallocHost(&d.mem, d.alloc_bytes), "cudaMallocHost");
        break;
      case DataLocation::CPU_MALLOC:
        d.mem = std::malloc(bytes);
        if (!d.mem) throw std::bad_alloc();
        break;
      default: throw std::invalid_argument("allocate: unsupported location");
    }
    d.location = where;
    return d;
  }
  /// non-owning view of foreign memory
  static GenericDataset foreign(void* data, uint64_t N, uint32_t D, DataType type, DataLocation where, int32_t gpu_id = -1)
  {
    GenericDataset d;
    d.N = N;
    d.D = D;
    d.type = type;
    d.location = where;
    d.gpu_id = gpu_id;
    d.mem = data;
    return d;
  }

 protected:
  void* mem{nullptr};
  size_t alloc_bytes{0};  // pinned allocations: the size the buffer was obtained with (pool key)

  template <typename T>
  void check_type() const
  {
    if (mem && DataType_v<T> != type) throw std::invalid_argument("dataset holds a different element type");
  }
  void take(GenericDataset& o) noexcept
  {
    N = o.N; D = o.D; type = o.type; location = o.location; gpu_id = o.gpu_id; mem = o.mem; alloc_bytes = o.alloc_bytes;
    o.mem = nullptr; o.N = 0; o.location = DataLocation::UNKNOWN; o.alloc_bytes = 0;
  }
  void release() noexcept
  {
    if (!mem) return;
    switch (location) {
      case DataLocation::GPU:
      case DataLocation::MANAGED: cudaFree(mem); break;
      case DataLocation::CPU_PINNED:
        if (!alloc_bytes || !detail::PinnedPool::instance().put(mem, alloc_bytes)) cudaFreeHost(mem);
        break;
      case DataLocation::CPU_MALLOC: std::free(mem); break;
      default: break;  // FOREIGN_*: never freed (src/ggnn/base/data.cu:147-149)
    }
    mem = nullptr;
  }
};

/// typed view of a GenericDataset with the std::span-like access of the reference's Dataset<T> (dataset.cuh:93-160)
template <typename T>
struct Dataset : public GenericDataset {
  using element_type = T;
  using value_type = std::remove_cv_t<T>;
  using iterator = T*;

  Dataset() { type = DataType_v<T>; }
  Dataset(GenericDataset&& g) : GenericDataset{std::move(g)}
  {
    if (!mem) type = DataType_v<T>;
    check_type<T>();
  }
  Dataset(Dataset&&) noexcept = default;
  Dataset& operator=(Dataset&&) noexcept = default;

  T* data() const { return static_cast<T*>(mem); }
  size_t size() const { return numel(); }
  size_t size_bytes() const { return numel() * sizeof(T); }
  bool empty() const { return numel() == 0; }
  T* begin() const { return data(); }
  T* end() const { return data() + numel(); }
  T& operator[](size_t i) const { return data()[i]; }
  T& at(size_t i) const
  {
    if (i >= numel()) throw std::out_of_range("Index " + std::to_string(i) + " is out of bounds (size " + std::to_string(numel()) + ").");
    return data()[i];
  }
  T& front() const { return data()[0]; }
  T& back() const { return data()[numel() - 1]; }
  std::span<T> subspan(size_t offset, size_t count = std::dynamic_extent) const { return std::span<T>{data(), numel()}.subspan(offset, count); }
  operator std::span<T>() const { return {data(), numel()}; }
  operator T*() { return data(); }
  operator const T*() const { return data(); }

  static Dataset empty(uint64_t N, uint32_t D, bool pin_memory = false)
  {
    return Dataset{allocate(N, D, DataType_v<T>, pin_memory ? DataLocation::CPU_PINNED : DataLocation::CPU_MALLOC)};
  }
  static Dataset emptyOnGPU(uint64_t N, uint32_t D, int32_t gpu_id)
  {
    return Dataset{allocate(N, D, DataType_v<T>, DataLocation::GPU, gpu_id)};
  }
  static Dataset copy(const std::span<const T>& data, uint32_t D, bool pin_memory = false)
  {
    if (D == 0 || data.size() % D) throw std::invalid_argument("data size is not a multiple of D");
    Dataset d = empty(data.size() / D, D, pin_memory);
    std::memcpy(d.data(), data.data(), d.size_bytes());
    return d;
  }
  static Dataset referenceCPUData(T* data, uint64_t N, uint32_t D)
  {
    return Dataset{foreign(const_cast<value_type*>(data), N, D, DataType_v<T>, DataLocation::FOREIGN_CPU)};
  }
  static Dataset referenceGPUData(T* data, uint64_t N, uint32_t D, int32_t gpu_id)
  {
    return Dataset{foreign(const_cast<value_type*>(data), N, D, DataType_v<T>, DataLocation::FOREIGN_GPU, gpu_id)};
  }

  /// [u]vecs files: per row a 4-byte dimension followed by D values (src/ggnn/base/dataset.cu:133-202); asking for
  /// more rows than the file holds is an error there (CHECK_EQ) and here
  static Dataset load(const std::filesystem::path& file, uint32_t from = 0,
                      uint32_t num = std::numeric_limits<uint32_t>::max(), bool pin_memory = false)
  {
    std::ifstream f(file, std::ios::binary);
    if (!f) throw std::runtime_error("Unable to open file " + file.string() + " for reading.");
    uint32_t dim = 0;
    f.read(reinterpret_cast<char*>(&dim), 4);
    if (!f || dim == 0) throw std::runtime_error("Failed to read vectors from " + file.string() + ".");
    const size_t rec = 4 + static_cast<size_t>(dim) * sizeof(T);
    const size_t total = static_cast<size_t>(std::filesystem::file_size(file)) / rec;
    if (from > total) throw std::out_of_range("Dataset contains fewer vectors than requested.");
    size_t n = total - from;
    if (num != std::numeric_limits<uint32_t>::max()) {
      if (n < num) throw std::out_of_range("Dataset contains fewer vectors than requested.");
      n = num;
    }
    Dataset d = empty(n, dim, pin_memory);
    constexpr size_t rows_per_block = 4096;
    std::vector<char> buf(rows_per_block * rec);
    f.seekg(static_cast<std::streamoff>(from * rec));
    for (size_t r0 = 0; r0 < n; r0 += rows_per_block) {
      const size_t nr = std::min(rows_per_block, n - r0);
      f.read(buf.data(), static_cast<std::streamsize>(nr * rec));
      if (!f) throw std::runtime_error("Failed to read vectors from " + file.string() + ".");
      for (size_t r = 0; r < nr; ++r) std::memcpy(d.data() + (r0 + r) * dim, buf.data() + r * rec + 4, dim * sizeof(T));
    }
    return d;
  }
  void store(const std::filesystem::path& file) const
  {
    if (!isCPUAccessible()) throw std::runtime_error("store() needs CPU-accessible data");
    std::ofstream f(file, std::ios::binary | std::ios::trunc);
    if (!f) throw std::runtime_error("Unable to open file " + file.string() + " for writing.");
    const uint32_t dim = D;
    for (uint64_t r = 0; r < N; ++r) {
      f.write(reinterpret_cast<const char*>(&dim), 4);
      f.write(reinterpret_cast<const char*>(data() + r * D), static_cast<std::streamsize>(D * sizeof(T)));
    }
  }
  void copyTo(Dataset& other, cudaStream_t stream = nullptr) const { copyRangeTo(0, N, other, stream); }
  /// rows [from, from + num) -> the first num rows of `other`
  void copyRangeTo(uint64_t from, uint64_t num, Dataset& other, cudaStream_t stream = nullptr) const
  {
    if (from + num > N || num > other.N || other.D != D) throw std::invalid_argument("copyRangeTo: shape mismatch");
    if (isCPUAccessible() && other.isCPUAccessible()) {  // host to host: no CUDA call needed
      std::memcpy(other.data(), data() + from * D, num * D * sizeof(T));
      return;
    }
    detail::cuda_check(cudaMemcpyAsync(other.data(), data() + from * D, num * D * sizeof(T), cudaMemcpyDefault, stream), "cudaMemcpyAsync");
    if (!isGPUAccessible() || !other.isGPUAccessible()) detail::cuda_check(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  }
  Dataset clone(cudaStream_t stream = nullptr) const
  {
    Dataset d = isGPUAccessible() && !isCPUAccessible() ? emptyOnGPU(N, D, gpu_id) : empty(N, D, location == DataLocation::CPU_PINNED);
    if (numel()) copyTo(d, stream);
    return d;
  }
  /// the data itself if it already lives on `gpu`, else a device copy (dataset.cu:325-334)
  Dataset referenceOnGPU(int gpu, cudaStream_t stream = nullptr) const
  {
    if (isGPUAccessible() && (gpu_id == gpu || location == DataLocation::MANAGED)) return Dataset{reference()};
    Dataset d = emptyOnGPU(N, D, gpu);
    detail::DeviceGuard g(gpu);
    detail::cuda_check(cudaMemcpyAsync(d.data(), data(), size_bytes(), cudaMemcpyDefault, stream), "cudaMemcpyAsync");
    return d;
  }
};

inline GenericDataset GenericDataset::load(const std::filesystem::path& file, uint32_t from, uint32_t num, bool pin_memory)
{
  const std::string name = file.string();
  if (name.ends_with(".fvecs")) return GenericDataset{Dataset<float>::load(file, from, num, pin_memory)};
  if (name.ends_with(".bvecs")) return GenericDataset{Dataset<uint8_t>::load(file, from, num, pin_memory)};
  if (name.ends_with(".ivecs")) return GenericDataset{Dataset<int32_t>::load(file, from, num, pin_memory)};
  throw std::runtime_error("Could not guess file type from " + name + ". fvecs, bvecs, or ivecs file required.");
}

template <typename KeyT, typename ValueT>
struct Results {
  Dataset<KeyT> ids{};
  Dataset<ValueT> dists{};
};

// ------------------------------------------------------------------------------------------------
// Evaluation (include/ggnn/base/eval.h:31-65, src/ggnn/base/eval.cpp:37-242)
// ------------------------------------------------------------------------------------------------
struct GTDuplicates {
  std::vector<uint32_t> top1DuplicateEnd{};
  std::vector<uint32_t> topKDuplicateEnd{};
};

struct Evaluation {
  uint32_t KQuery{};
  float c1{0}, c1_dup{0}, cKQuery{0}, cKQuery_dup{0}, rKQuery{0}, rKQuery_dup{0};
};

inline std::ostream& operator<<(std::ostream& os, const Evaluation& e)
{
  auto dup = [&os](float v, bool newline) {
    if (!std::isnan(v)) os << " +duplicates: " << v;
    else os << " (duplicates unknown)";
    if (newline) os << '\n';
  };
  os << "c@1 (=r@1): " << e.c1;
  dup(e.c1_dup, true);
  os << "c@" << e.KQuery << ": " << e.cKQuery;
  dup(e.cKQuery_dup, true);
  os << "r@" << e.KQuery << ": " << e.rKQuery;
  dup(e.rKQuery_dup, false);
  return os;
}

template <typename KeyT, typename ValueT>
struct Evaluator {
  uint32_t KQuery{0};
  DistanceMeasure measure{};
  Dataset<KeyT> gt;
  GTDuplicates gt_duplicates{};

  Evaluator() = default;

  /// clones the ground truth; if base and query are CPU-accessible, finds for every query how far the ground-truth
  /// prefix extends over distance ties (<= 1e-6) behind rank 1 and rank KQuery (eval.cpp:88-174)
  Evaluator(const GenericDataset& base, const GenericDataset& query, const Dataset<KeyT>& gt_in, uint32_t KQuery_,
            DistanceMeasure measure_)
      : KQuery{KQuery_}, measure{measure_}, gt{gt_in.clone()}
  {
    if (!base.N || !query.N) return;                                   // no duplicate information
    if (!base.isCPUAccessible() || !query.isCPUAccessible()) return;   // ditto
    if (!gt_in.isCPUAccessible()) throw std::runtime_error("Ground truth data needs to be given on the CPU for evaluation.");
    if (base.type != query.type) throw std::invalid_argument("base and query have different data types");
    if (base.type != DataType::FLOAT && base.type != DataType::UINT8) throw std::runtime_error("unsupported data type");
    const bool is_float = base.type == DataType::FLOAT;
    const auto bf = base.reinterpret<float>(), qf = query.reinterpret<float>();
    const auto bu = base.reinterpret<uint8_t>(), qu = query.reinterpret<uint8_t>();
    const size_t Db = base.D;
    auto value = [&](bool of_base, size_t row, size_t d) -> ValueT {
      const size_t i = row * Db + d;
      return is_float ? static_cast<ValueT>(of_base ? bf[i] : qf[i]) : static_cast<ValueT>(of_base ? bu[i] : qu[i]);
    };
    // eval.cpp:37-65 (float accumulation in index order; the cosine b_norm is computed from `a` there, too)
    auto dist_to_query = [&](size_t base_idx, size_t query_idx) -> ValueT {
      if (base_idx >= base.N) throw std::out_of_range("ground truth index out of range");
      ValueT acc = 0.0f, a_norm = 0.0f, b_norm = 0.0f;
      for (size_t d = 0; d < Db; ++d) {
        const ValueT a = value(true, base_idx, d), b = value(false, query_idx, d);
        if (measure == DistanceMeasure::Euclidean) acc += (a - b) * (a - b);
        else {
          acc += a * b;
          a_norm += a * a;
          b_norm += a * a;
        }
      }
      if (measure == DistanceMeasure::Euclidean) return std::sqrt(acc);
      return (a_norm * b_norm > 0.0f) ? std::fabs(1.0f - acc / std::sqrt(a_norm * b_norm)) : 1.0f;
    };
    constexpr float Epsilon = 0.000001f;
    const uint32_t GD = gt.D;
    gt_duplicates.top1DuplicateEnd.reserve(query.N);
    gt_duplicates.topKDuplicateEnd.reserve(query.N);
    for (uint32_t n = 0; n < query.N; ++n) {
      const KeyT* row = gt.data() + static_cast<size_t>(n) * GD;
      auto run_after = [&](uint32_t anchor_rank, uint32_t first) {
        const ValueT anchor = dist_to_query(static_cast<size_t>(row[anchor_rank]), n);
        uint32_t len = 0;
        for (uint32_t k = first; k < GD; ++k) {
          if (dist_to_query(static_cast<size_t>(row[k]), n) - anchor > Epsilon) break;
          ++len;
        }
        return len;
      };
      gt_duplicates.top1DuplicateEnd.push_back(1 + run_after(0, 1));
      gt_duplicates.topKDuplicateEnd.push_back(KQuery <= GD ? KQuery + run_after(KQuery - 1, KQuery) : GD);
    }
  }

  [[nodiscard]] Evaluation evaluateResults(const Dataset<KeyT>& results)
  {
    if (!gt.D) throw std::runtime_error("No ground truth data loaded. cannot compute accuracy.");
    if (gt.N < results.N) throw std::out_of_range("more result rows than ground truth rows");
    if (!results.isCPUAccessible()) throw std::runtime_error("Results need to be given on the CPU for evaluation.");
    const bool has_dup = !gt_duplicates.top1DuplicateEnd.empty() && !gt_duplicates.topKDuplicateEnd.empty();
    uint32_t c1 = 0, c1_dup = 0, cK = 0, cK_dup = 0, rK = 0, rK_dup = 0;
    for (uint32_t n = 0; n < results.N; ++n) {
      const uint32_t endTop1 = has_dup ? gt_duplicates.top1DuplicateEnd.at(n) : 1;
      const uint32_t endTopK = has_dup ? gt_duplicates.topKDuplicateEnd.at(n) : KQuery;
      if (endTopK > gt.D) throw std::out_of_range("ground truth has fewer than KQuery columns");
      const KeyT* g = gt.data() + static_cast<size_t>(n) * gt.D;
      for (uint32_t kr = 0; kr < KQuery; ++kr) {
        const KeyT q = results[static_cast<size_t>(n) * KQuery + kr];
        for (uint32_t kg = 0; kg < endTopK; ++kg) {  // eval.cpp:206-227: every match counts (no early exit)
          if (q != g[kg]) continue;
          if (kg == 0) {
            c1 += (kr == 0);
            rK += (kg < KQuery);
            ++rK_dup;
          }
          if (kg < endTop1) c1_dup += (kr == 0);
          cK += (kg < KQuery);
          ++cK_dup;
        }
      }
    }
    const float inv_q = 1.0f / static_cast<float>(results.N);
    const float inv_r = 1.0f / static_cast<float>(results.N * KQuery);
    const float nan = std::numeric_limits<float>::quiet_NaN();
    Evaluation e;
    e.KQuery = KQuery;
    e.c1 = static_cast<float>(c1) * inv_q;
    e.c1_dup = has_dup ? static_cast<float>(c1_dup) * inv_q : nan;
    e.cKQuery = static_cast<float>(cK) * inv_r;
    e.cKQuery_dup = has_dup ? static_cast<float>(cK_dup) * inv_r : nan;
    e.rKQuery = static_cast<float>(rK) * inv_q;
    e.rKQuery_dup = has_dup ? static_cast<float>(rK_dup) * inv_q : nan;
    return e;
  }
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

  /// A query enqueued by queryAsync(): get() waits for it and hands out the results (pinned host memory).
  class QueryHandle {
   public:
    QueryHandle() = default;
    QueryHandle(QueryHandle&&) noexcept = default;
    QueryHandle& operator=(QueryHandle&&) noexcept = default;
    [[nodiscard]] bool valid() const { return event != nullptr; }
    [[nodiscard]] bool done() const { return !event || cudaEventQuery(event) == cudaSuccess; }
    [[nodiscard]] Results get()
    {
      if (event) detail::cuda_check(cudaEventSynchronize(event), "cudaEventSynchronize");
      event = nullptr;
      return std::move(results);
    }

   private:
    friend class GGNN;
    Results results;
    cudaEvent_t event{nullptr};  // owned by the GGNN instance's pipeline
  };

  GGNN() : st(std::make_unique<State>()) {}
  ~GGNN() = default;
  GGNN(const GGNN&) = delete;
  GGNN& operator=(const GGNN&) = delete;
  // all state lives behind one pointer: a moved GGNN keeps referring to its own base (the reference is movable too,
  // include/ggnn/base/ggnn.cuh:55-58)
  GGNN(GGNN&&) noexcept = default;
  GGNN& operator=(GGNN&&) noexcept = default;

  void setWorkingDirectory(const std::filesystem::path& dir) { st->graph_dir = dir; }
  /// host memory the swapped-out graphs may use; beyond it they live in part_<id>.ggnn files (gpu_instance.cu:177-200)
  void setCPUMemoryLimit(size_t limit) { st->cpu_memory_limit = limit; }
  /// GPU memory left free when counting how many shards fit on a GPU (gpu_instance.cu:150-175)
  void setReservedGPUMemory(size_t reserved) { st->reserved_gpu_memory = reserved; }
  void setGPUs(const std::span<const int>& ids)
  {
    if (!st->shards.empty()) throw std::runtime_error("GPUs cannot be changed after the graph has been set up.");
    if (ids.empty()) throw std::out_of_range("at least one GPU is required");
    st->gpu_ids.assign(ids.begin(), ids.end());
  }
  void setGPUs(const std::vector<int>& ids) { setGPUs(std::span<const int>{ids.data(), ids.size()}); }
  void setShardSize(uint32_t n)
  {
    if (!st->shards.empty()) throw std::runtime_error("The shard size cannot be changed after the graph has been set up.");
    st->N_shard = n;
  }
  /// (the reference allows this for one GPU only, ggnn.cu:299-306; here the merged lists live on the first GPU)
  void setReturnResultsOnGPU(bool flag = true) { st->return_results_on_gpu = flag; }

  void setBase(GenericDataset&& b)
  {
    check_base(b);
    st->owned_base = std::move(b);
    st->base = &st->owned_base;
  }
  void setBaseReference(const GenericDataset& b)
  {
    check_base(b);
    st->base = &b;
  }
  void setBaseReference(GenericDataset&&) = delete;

  void build(uint32_t KBuild, float tau_build, uint32_t refinement_iterations = 2, DistanceMeasure measure = DistanceMeasure::Euclidean)
  {
    prepare(KBuild);
    State& s = *st;
    for (auto& g : s.gpus) {
      detail::DeviceGuard guard(g.id);
      const size_t scratch_bytes = ggnn_b200_build_scratch_bytes(&s.cfg);
      void* scratch = nullptr;
      detail::cuda_check(cudaMalloc(&scratch, scratch_bytes), "cudaMalloc(build scratch)");
      float* d_rng = nullptr;
      detail::cuda_check(cudaMalloc(reinterpret_cast<void**>(&d_rng), (static_cast<size_t>(s.cfg.Ns[0]) + s.cfg.Ns[1] + s.cfg.Ns[2]) * sizeof(float)),
                         "cudaMalloc(build uniforms)");
      int rc = 0;
      for (uint32_t i = 0; i < s.spg && !rc; ++i) {
        Shard& sh = s.shards[g.first_shard + i];
        Slot& slot = acquire(g, sh);
        detail::cuda_check(cudaMemsetAsync(slot.blob.data(), 0, slot.blob.size_bytes(), g.stream), "cudaMemsetAsync");
        // one cuRAND generator per GPU, continuing over its shards like the reference's (graph_construction.cu:96-102,127)
        if (!g.rng) detail::abi_check(ggnn_b200_rng_create(&g.rng, 1234ULL));
        detail::abi_check(ggnn_b200_rng_fill_build(g.rng, &s.cfg, d_rng, g.stream));
        rc = ggnn_b200_build_graph(&s.cfg, rows_f32(g, slot), static_cast<int>(measure), tau_build, refinement_iterations,
                                   d_rng, slot.blob.data(), scratch, scratch_bytes, g.stream);
        cudaStreamSynchronize(g.stream);
        drop_f32(g, slot);
        sh.has_graph = rc == 0;
        sh.dirty = true;
      }
      cudaFree(scratch);
      cudaFree(d_rng);
      detail::abi_check(rc);
    }
  }

  void store()
  {
    State& s = *st;
    if (!has_graph()) throw std::runtime_error("There is no graph to store.");
    std::filesystem::create_directories(s.graph_dir);
    for (auto& g : s.gpus) {
      detail::DeviceGuard guard(g.id);
      for (uint32_t i = 0; i < s.spg; ++i) {  // gpu_instance.cu:86-115: part_<global_shard_id>.ggnn = raw blob
        Shard& sh = s.shards[g.first_shard + i];
        const auto path = part_path(sh.global_id);
        if (sh.slot >= 0) {  // newest copy is on the GPU
          std::vector<uint8_t> h(s.blob_bytes);
          detail::cuda_check(cudaMemcpy(h.data(), g.slots[sh.slot].blob.data(), h.size(), cudaMemcpyDeviceToHost), "cudaMemcpy");
          write_file(path, h.data(), h.size());
        }
        else if (sh.host_blob.data()) write_file(path, sh.host_blob.data(), s.blob_bytes);
        // else: already on disk as part_<id>.ggnn
      }
    }
  }

  void load(uint32_t KBuild)
  {
    prepare(KBuild);
    State& s = *st;
    for (auto& g : s.gpus) {
      detail::DeviceGuard guard(g.id);
      for (uint32_t i = 0; i < s.spg; ++i) {
        Shard& sh = s.shards[g.first_shard + i];
        const auto path = part_path(sh.global_id);
        if (!std::filesystem::exists(path) || std::filesystem::file_size(path) != s.blob_bytes)
          throw std::runtime_error(path.string() + ": missing or unexpected file size");  // gpu_instance.cu:454-455
        if (g.swap) {  // read when the shard is swapped in
          if (sh.slot >= 0) {
            g.slots[sh.slot].resident = -1;
            sh.slot = -1;
          }
          sh.host_blob = Dataset<uint8_t>{};
          sh.on_disk = true;
        }
        else {
          std::vector<uint8_t> h(s.blob_bytes);
          read_file(path, h.data(), h.size());
          detail::cuda_check(cudaMemcpy(g.slots[sh.slot].blob.data(), h.data(), h.size(), cudaMemcpyHostToDevice), "cudaMemcpy");
        }
        sh.has_graph = true;
        sh.dirty = false;
      }
    }
  }

  [[nodiscard]] Results query(const GenericDataset& query, uint32_t KQuery, float tau_query, uint32_t max_iterations = 400,
                              DistanceMeasure measure = DistanceMeasure::Euclidean)
  {
    State& s = *st;
    if (!has_graph()) throw std::runtime_error("There is no graph to query.");
    check_query(query, "unsupported datatype for query");
    const uint32_t n_gpus = static_cast<uint32_t>(s.gpus.size());
    const uint32_t Nq = static_cast<uint32_t>(query.N);
    // one GPU, host query, host results, everything resident: the pipelined path (no allocation per call)
    if (n_gpus == 1 && !s.return_results_on_gpu && query.isCPUAccessible() && !s.gpus[0].swap)
      return queryAsync(query, KQuery, tau_query, max_iterations, measure).get();

    // fused shard-merge exchange for several GPUs: every GPU's traversal kernels store their lists straight into one
    // buffer on the first GPU (peer access); fall back to peer copies when the devices cannot see each other
    const bool gather = n_gpus > 1 && s.peer_gather;
    const uint32_t n_slots = n_gpus * s.spg;
    const size_t list_words = static_cast<size_t>(Nq) * KQuery;
    Gpu& g0 = s.gpus[0];
    Dataset<KeyT> all;
    if (gather) {
      detail::DeviceGuard guard(g0.id);
      all = Dataset<KeyT>::emptyOnGPU(static_cast<uint64_t>(2) * n_slots * Nq, KQuery, g0.id);
    }
    std::vector<Dataset<KeyT>> ids(n_gpus);
    std::vector<Dataset<ValueT>> dists(n_gpus);
    std::vector<Dataset<float>> q_dev(n_gpus);      // fp32 copy of the query batch on each GPU (made when a kernel needs it)
    std::vector<Dataset<uint8_t>> q_u8(n_gpus);     // uint8 copy, for shards searched from native 1-byte rows
    std::vector<cudaEvent_t> evs(n_gpus, nullptr);
    // launch everything asynchronously on every GPU first
    for (uint32_t gi = 0; gi < n_gpus; ++gi) {
      Gpu& g = s.gpus[gi];
      detail::DeviceGuard guard(g.id);
      if (!gather) {
        ids[gi] = Dataset<KeyT>::emptyOnGPU(Nq, KQuery * s.spg, g.id);
        dists[gi] = Dataset<ValueT>::emptyOnGPU(Nq, KQuery * s.spg, g.id);
      }
      else {
        void* dst = all.data();
        detail::cuda_check(cudaMemcpyAsync(g.d_gather_tbl, &dst, sizeof(void*), cudaMemcpyHostToDevice, g.stream), "cudaMemcpyAsync");
      }
      // swap mode: alternate the direction from call to call, so that the shards left on the GPU by the previous call
      // are searched first (gpu_instance.cu:669-670, 740)
      const bool reverse = g.swap && (g.query_calls++ % 2);
      for (uint32_t i = 0; i < s.spg; ++i) {
        const uint32_t sidx = reverse ? s.spg - 1 - i : i;
        Shard& sh = s.shards[g.first_shard + sidx];
        Slot& slot = acquire(g, sh);
        const bool native = native_u8(slot, query, KQuery);
        if (native) {
          if (!q_u8[gi].data()) q_u8[gi] = u8_on_gpu(query, g.id, g.stream);
        }
        else {
          if (!q_dev[gi].data()) q_dev[gi] = float_on_gpu(query, g.id, g.stream);
          rows_f32(g, slot);
        }
        ggnn_b200_query_params p = query_params(slot, q_dev[gi].data(), KQuery, tau_query, max_iterations, measure);
        if (native) {
          p.base_type = GGNN_B200_BASE_U8;
          p.d_base = reinterpret_cast<const float*>(slot.base_u8.data());
          p.d_query = reinterpret_cast<const float*>(q_u8[gi].data());
        }
        if (gather) {
          p.n_scatter = 1;
          p.scatter_slot = gi * s.spg + sidx;
          p.scatter_rows = Nq;
          p.scatter_dists_offset = static_cast<size_t>(n_slots) * list_words * sizeof(KeyT);
          p.d_scatter_dst = reinterpret_cast<void* const*>(g.d_gather_tbl);
        }
        else {
          p.d_query_results = ids[gi].data();
          p.d_query_results_dists = dists[gi].data();
          p.shards_per_gpu = s.spg;
          p.on_gpu_shard_id = sidx;
        }
        p.d_work_counter = g.work_counter;
        detail::abi_check(ggnn_b200_query(&p, Nq, g.stream));
      }
      if (!gather && s.spg > 1) {  // replaces gpu_instance.cu:745-790
        Dataset<KeyT> mi = Dataset<KeyT>::emptyOnGPU(Nq, KQuery, g.id);
        Dataset<ValueT> md = Dataset<ValueT>::emptyOnGPU(Nq, KQuery, g.id);
        detail::abi_check(ggnn_b200_merge_topk(ids[gi].data(), dists[gi].data(), s.spg, KQuery, static_cast<size_t>(KQuery) * s.spg,
                                               KQuery, Nq, KQuery, 0, mi.data(), md.data(), g.stream));
        detail::cuda_check(cudaStreamSynchronize(g.stream), "cudaStreamSynchronize");  // the inputs are freed below
        ids[gi] = std::move(mi);
        dists[gi] = std::move(md);
      }
      if (gather && gi > 0) {
        detail::cuda_check(cudaEventCreateWithFlags(&evs[gi], cudaEventDisableTiming), "cudaEventCreate");
        detail::cuda_check(cudaEventRecord(evs[gi], g.stream), "cudaEventRecord");
      }
    }
    Results out;
    if (n_gpus == 1) {
      detail::DeviceGuard guard(g0.id);
      detail::cuda_check(cudaStreamSynchronize(g0.stream), "cudaStreamSynchronize");
      out.ids = std::move(ids[0]);
      out.dists = std::move(dists[0]);
    }
    else {  // replaces ResultMerger::merge (result_merger.cpp:51-149): one merge kernel on the first GPU
      detail::DeviceGuard guard(g0.id);
      const KeyT* src_i = nullptr;
      const ValueT* src_d = nullptr;
      Dataset<KeyT> all_i;
      Dataset<ValueT> all_d;
      uint32_t lists = n_gpus;
      int64_t id_off = static_cast<int64_t>(s.spg) * s.cfg.N;
      if (gather) {
        for (uint32_t gi = 1; gi < n_gpus; ++gi) detail::cuda_check(cudaStreamWaitEvent(g0.stream, evs[gi], 0), "cudaStreamWaitEvent");
        src_i = all.data();
        src_d = reinterpret_cast<const ValueT*>(all.data() + static_cast<size_t>(n_slots) * list_words);
        lists = n_slots;
        id_off = s.cfg.N;
      }
      else {
        all_i = Dataset<KeyT>::emptyOnGPU(static_cast<uint64_t>(n_gpus) * Nq, KQuery, g0.id);
        all_d = Dataset<ValueT>::emptyOnGPU(static_cast<uint64_t>(n_gpus) * Nq, KQuery, g0.id);
        for (uint32_t gi = 0; gi < n_gpus; ++gi) {
          Gpu& g = s.gpus[gi];
          detail::DeviceGuard guard_i(g.id);
          detail::cuda_check(cudaMemcpyPeerAsync(all_i.data() + gi * list_words, g0.id, ids[gi].data(), g.id, list_words * sizeof(KeyT), g.stream), "cudaMemcpyPeerAsync");
          detail::cuda_check(cudaMemcpyPeerAsync(all_d.data() + gi * list_words, g0.id, dists[gi].data(), g.id, list_words * sizeof(ValueT), g.stream), "cudaMemcpyPeerAsync");
          detail::cuda_check(cudaStreamSynchronize(g.stream), "cudaStreamSynchronize");
        }
        src_i = all_i.data();
        src_d = all_d.data();
      }
      out.ids = Dataset<KeyT>::emptyOnGPU(Nq, KQuery, g0.id);
      out.dists = Dataset<ValueT>::emptyOnGPU(Nq, KQuery, g0.id);
      detail::abi_check(ggnn_b200_merge_topk(src_i, src_d, lists, list_words, KQuery, KQuery, Nq, KQuery, id_off, out.ids.data(),
                                             out.dists.data(), g0.stream));
      detail::cuda_check(cudaStreamSynchronize(g0.stream), "cudaStreamSynchronize");
      for (uint32_t gi = 1; gi < n_gpus; ++gi) {  // the other GPUs' query copies are freed on return
        if (evs[gi]) cudaEventDestroy(evs[gi]);
        detail::DeviceGuard guard_i(s.gpus[gi].id);
        cudaStreamSynchronize(s.gpus[gi].stream);
      }
    }
    return s.return_results_on_gpu ? std::move(out) : to_host(std::move(out));
  }

  /// query() for a host-resident query on one GPU without waiting: host->device copy, traversal and device->host copy
  /// of the results are enqueued on one of the instance's pipelines (own stream and device buffers, nothing is allocated
  /// on the device per call); several batches may be in flight, their copies overlap the other batches' kernels.
  /// (The reference's query is synchronous only: src/ggnn/base/ggnn.cu:506-551.)  `query` must stay valid until get().
  [[nodiscard]] QueryHandle queryAsync(const GenericDataset& query, uint32_t KQuery, float tau_query, uint32_t max_iterations = 400,
                                       DistanceMeasure measure = DistanceMeasure::Euclidean)
  {
    State& s = *st;
    if (!has_graph()) throw std::runtime_error("There is no graph to query.");
    check_query(query, "unsupported datatype for query");
    if (s.gpus.size() != 1 || !query.isCPUAccessible() || s.gpus[0].swap)
      throw std::runtime_error("queryAsync takes a host-resident query and a single GPU holding all of its shards.");
    Gpu& g = s.gpus[0];
    detail::DeviceGuard guard(g.id);
    const uint32_t Nq = static_cast<uint32_t>(query.N);
    if (g.pipes.empty()) {
      g.pipes.resize(4);
      for (auto& p : g.pipes) {
        detail::cuda_check(cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking), "cudaStreamCreate");
        detail::cuda_check(cudaEventCreateWithFlags(&p.done, cudaEventDisableTiming), "cudaEventCreate");
        detail::cuda_check(cudaMalloc(reinterpret_cast<void**>(&p.work_counter), 16), "cudaMalloc");
      }
    }
    Pipe& p = g.pipes[g.next_pipe++ % g.pipes.size()];
    detail::cuda_check(cudaEventSynchronize(p.done), "cudaEventSynchronize");  // the batch that used this pipeline last is complete
    const size_t q_bytes = static_cast<size_t>(Nq) * query.D * sizeof(float);
    const size_t r_words = static_cast<size_t>(Nq) * KQuery * s.spg;
    p.query.reserve(q_bytes);
    p.ids.reserve(r_words * sizeof(KeyT));
    p.dists.reserve(r_words * sizeof(ValueT));
    const float* dq = static_cast<const float*>(p.query.ptr);
    bool widened = false;
    if (query.type == DataType::FLOAT) {
      detail::cuda_check(cudaMemcpyAsync(p.query.ptr, static_cast<const void*>(query), q_bytes, cudaMemcpyHostToDevice, p.stream), "cudaMemcpyAsync(query)");
    }
    else {  // uint8: copy the bytes; widened on the device (exact) only if a shard is not searched from native 1-byte rows
      p.stage.reserve(query.numel());
      detail::cuda_check(cudaMemcpyAsync(p.stage.ptr, static_cast<const void*>(query), query.numel(), cudaMemcpyHostToDevice, p.stream), "cudaMemcpyAsync(query)");
    }
    KeyT* d_ids = static_cast<KeyT*>(p.ids.ptr);
    ValueT* d_dists = static_cast<ValueT*>(p.dists.ptr);
    for (uint32_t i = 0; i < s.spg; ++i) {
      Shard& sh = s.shards[g.first_shard + i];
      Slot& slot = g.slots[sh.slot];
      const bool native = native_u8(slot, query, KQuery);
      if (!native && query.type != DataType::FLOAT && !widened) {
        detail::abi_check(ggnn_b200_widen_u8(static_cast<const uint8_t*>(p.stage.ptr), static_cast<float*>(p.query.ptr), query.numel(), p.stream));
        widened = true;
      }
      if (!native && !slot.base.data()) {  // widened rows are made on the GPU's own stream: order this pipeline behind it
        rows_f32(g, slot);
        detail::cuda_check(cudaStreamSynchronize(g.stream), "cudaStreamSynchronize");
      }
      ggnn_b200_query_params qp = query_params(slot, dq, KQuery, tau_query, max_iterations, measure);
      if (native) {
        qp.base_type = GGNN_B200_BASE_U8;
        qp.d_base = reinterpret_cast<const float*>(slot.base_u8.data());
        qp.d_query = static_cast<const float*>(p.stage.ptr);
      }
      qp.d_query_results = d_ids;
      qp.d_query_results_dists = d_dists;
      qp.shards_per_gpu = s.spg;
      qp.on_gpu_shard_id = i;
      qp.d_work_counter = p.work_counter;
      detail::abi_check(ggnn_b200_query(&qp, Nq, p.stream));
    }
    if (s.spg > 1) {  // replaces gpu_instance.cu:745-790
      p.merged_ids.reserve(static_cast<size_t>(Nq) * KQuery * sizeof(KeyT));
      p.merged_dists.reserve(static_cast<size_t>(Nq) * KQuery * sizeof(ValueT));
      detail::abi_check(ggnn_b200_merge_topk(d_ids, d_dists, s.spg, KQuery, static_cast<size_t>(KQuery) * s.spg, KQuery, Nq, KQuery, 0,
                                             static_cast<KeyT*>(p.merged_ids.ptr), static_cast<ValueT*>(p.merged_dists.ptr), p.stream));
      d_ids = static_cast<KeyT*>(p.merged_ids.ptr);
      d_dists = static_cast<ValueT*>(p.merged_dists.ptr);
    }
    QueryHandle h;
    h.results.ids = Dataset<KeyT>::empty(Nq, KQuery, true);
    h.results.dists = Dataset<ValueT>::empty(Nq, KQuery, true);
    detail::cuda_check(cudaMemcpyAsync(h.results.ids.data(), d_ids, h.results.ids.size_bytes(), cudaMemcpyDeviceToHost, p.stream), "cudaMemcpyAsync(ids)");
    detail::cuda_check(cudaMemcpyAsync(h.results.dists.data(), d_dists, h.results.dists.size_bytes(), cudaMemcpyDeviceToHost, p.stream), "cudaMemcpyAsync(dists)");
    detail::cuda_check(cudaEventRecord(p.done, p.stream), "cudaEventRecord");
    h.event = p.done;
    return h;
  }

  [[nodiscard]] Results bfQuery(const GenericDataset& query, uint32_t KGT = 100, DistanceMeasure measure = DistanceMeasure::Euclidean)
  {
    State& s = *st;
    if (!s.base) throw std::runtime_error("The base needs to be set before running a brute-force query.");
    if (s.gpu_ids.size() > 1) throw std::runtime_error("bfQuery supports only a single GPU.");  // ggnn.cu:338-339
    check_query(query, "unsupported datatype for brute-force query");
    const int gpu = s.gpu_ids[0];
    detail::DeviceGuard g(gpu);
    Dataset<float> b_dev;
    const float* db = nullptr;
    Slot* single = (s.shards.size() == 1 && s.shards[0].slot >= 0) ? &s.gpus[0].slots[s.shards[0].slot] : nullptr;
    // uint8 base: exact integer contraction on the int8 tensor cores (ggnn_b200_bf_query_u8), rows never widened
    // (GGNN_B200_NO_I8_BF=1: widen and take the fp32 path instead -- identical results)
    if (s.base->type == DataType::UINT8 && query.type == DataType::UINT8 && !std::getenv("GGNN_B200_NO_I8_BF")) {
      const size_t ws_bytes = ggnn_b200_bf_query_u8_workspace_bytes(s.base->D, static_cast<int>(measure), KGT, static_cast<uint32_t>(s.base->N),
                                                                    static_cast<uint32_t>(query.N));
      if (ws_bytes) {
        Dataset<uint8_t> b8_dev;
        const uint8_t* b8 = nullptr;
        if (single && single->base_u8.data()) b8 = single->base_u8.data();
        else {
          b8_dev = u8_on_gpu(*s.base, gpu, nullptr);
          b8 = b8_dev.data();
        }
        Dataset<uint8_t> q8 = u8_on_gpu(query, gpu, nullptr);
        Results out;
        out.ids = Dataset<KeyT>::emptyOnGPU(query.N, KGT, gpu);
        out.dists = Dataset<ValueT>::emptyOnGPU(query.N, KGT, gpu);
        ggnn_b200_bf_query_params p{};
        p.D = s.base->D; p.measure = static_cast<int>(measure); p.KQuery = KGT; p.N_base = static_cast<int32_t>(s.base->N);
        p.d_base = reinterpret_cast<const float*>(b8); p.d_query = reinterpret_cast<const float*>(q8.data());
        p.d_query_results = out.ids.data(); p.d_query_results_dists = out.dists.data();
        p.workspace_bytes = ws_bytes;
        detail::cuda_check(cudaMalloc(&p.d_workspace, p.workspace_bytes), "cudaMalloc(bf workspace)");
        if (single) detail::cuda_check(cudaStreamSynchronize(s.gpus[0].stream), "cudaStreamSynchronize");
        const int rc = ggnn_b200_bf_query_u8(&p, static_cast<uint32_t>(query.N), nullptr);
        cudaDeviceSynchronize();
        cudaFree(p.d_workspace);
        detail::abi_check(rc);
        return s.return_results_on_gpu ? std::move(out) : to_host(std::move(out));
      }
    }
    if (single) {
      db = rows_f32(s.gpus[0], *single);
      detail::cuda_check(cudaStreamSynchronize(s.gpus[0].stream), "cudaStreamSynchronize");
    }
    else {
      b_dev = float_on_gpu(*s.base, gpu, nullptr);
      db = b_dev.data();
    }
    Dataset<float> q_dev = float_on_gpu(query, gpu, nullptr);
    Results out;
    out.ids = Dataset<KeyT>::emptyOnGPU(query.N, KGT, gpu);
    out.dists = Dataset<ValueT>::emptyOnGPU(query.N, KGT, gpu);
    ggnn_b200_bf_query_params p{};
    p.D = s.base->D; p.measure = static_cast<int>(measure); p.KQuery = KGT; p.N_base = static_cast<int32_t>(s.base->N);
    p.d_base = db; p.d_query = q_dev.data(); p.d_query_results = out.ids.data(); p.d_query_results_dists = out.dists.data();
    p.workspace_bytes = ggnn_b200_bf_query_workspace_bytes(p.D, p.measure, KGT, static_cast<uint32_t>(s.base->N), static_cast<uint32_t>(query.N));
    if (p.workspace_bytes) detail::cuda_check(cudaMalloc(&p.d_workspace, p.workspace_bytes), "cudaMalloc(bf workspace)");
    const int rc = ggnn_b200_bf_query(&p, static_cast<uint32_t>(query.N), nullptr);
    cudaDeviceSynchronize();
    if (p.d_workspace) cudaFree(p.d_workspace);
    if (single) drop_f32(s.gpus[0], *single);
    detail::abi_check(rc);
    return s.return_results_on_gpu ? std::move(out) : to_host(std::move(out));
  }

  /// the graph of one shard on its GPU (swap mode: valid until the shard is swapped out)
  [[nodiscard]] const Graph& getGraph(uint32_t global_shard_id = 0)
  {
    State& s = *st;
    if (global_shard_id >= s.shards.size()) throw std::out_of_range("no such shard");
    Shard& sh = s.shards[global_shard_id];
    Gpu& g = s.gpus[global_shard_id / s.spg];
    detail::DeviceGuard guard(g.id);
    Slot& slot = acquire(g, sh);
    cudaStreamSynchronize(g.stream);
    s.graph_view.config = s.cfg;
    s.graph_view.offsets = s.off;
    s.graph_view.memory = Dataset<uint8_t>{slot.blob.reference()};
    return s.graph_view;
  }

 private:
  /// grow-only device buffer
  struct DevBuf {
    void* ptr{nullptr};
    size_t cap{0};
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : ptr(o.ptr), cap(o.cap) { o.ptr = nullptr; o.cap = 0; }
    ~DevBuf() { if (ptr) cudaFree(ptr); }
    void reserve(size_t bytes)
    {
      if (bytes <= cap) return;
      if (ptr) cudaFree(ptr);
      ptr = nullptr;
      cap = 0;
      detail::cuda_check(cudaMalloc(&ptr, std::max<size_t>(bytes, 256)), "cudaMalloc(pipeline buffer)");
      cap = bytes;
    }
  };
  struct Pipe {
    cudaStream_t stream{nullptr};
    cudaEvent_t done{nullptr};
    uint32_t* work_counter{nullptr};
    DevBuf query, stage, ids, dists, merged_ids, merged_dists;
  };
  /// one (base rows, graph blob) buffer pair on a GPU
  struct Slot {
    Dataset<float> base;        // fp32 rows (for a uint8 base: widened on demand, see rows_f32 / drop_f32)
    Dataset<uint8_t> base_u8;   // native 1-byte rows of a uint8 base (resident mode)
    Dataset<uint8_t> blob;
    int64_t resident{-1};  // global shard id held, -1 = free
    uint64_t last_use{0};
  };
  struct Shard {
    uint32_t global_id{0};
    int slot{-1};           // index into its GPU's slots, -1 = swapped out
    bool has_graph{false}, dirty{false}, on_disk{false};
    Dataset<uint8_t> host_blob;  // swapped-out graph in pinned host memory (else on disk / none yet)
  };
  struct Gpu {
    int id{0};
    uint32_t first_shard{0};
    cudaStream_t stream{nullptr};
    uint32_t* work_counter{nullptr};
    void** d_gather_tbl{nullptr};
    ggnn_b200_rng* rng{nullptr};
    bool swap{false};
    uint32_t query_calls{0};
    uint64_t clock{0};
    std::vector<Slot> slots;
    std::vector<Pipe> pipes;
    size_t next_pipe{0};
  };
  struct State {
    std::filesystem::path graph_dir{"."};
    std::vector<int> gpu_ids{0};
    uint32_t N_shard{0}, spg{1};
    bool return_results_on_gpu{false};
    bool peer_gather{false};
    size_t cpu_memory_limit{std::numeric_limits<size_t>::max()}, reserved_gpu_memory{0};
    size_t blob_bytes{0}, host_blobs{0};
    GenericDataset owned_base{};
    const GenericDataset* base{nullptr};
    ggnn_b200_graph_config cfg{};
    ggnn_b200_graph_offsets off{};
    std::vector<Shard> shards;
    std::vector<Gpu> gpus;
    Graph graph_view;
    ~State()
    {
      for (auto& g : gpus) {
        cudaSetDevice(g.id);
        cudaDeviceSynchronize();
        for (auto& p : g.pipes) {
          if (p.stream) cudaStreamDestroy(p.stream);
          if (p.done) cudaEventDestroy(p.done);
          if (p.work_counter) cudaFree(p.work_counter);
        }
        g.pipes.clear();
        if (g.stream) cudaStreamDestroy(g.stream);
        if (g.work_counter) cudaFree(g.work_counter);
        if (g.d_gather_tbl) cudaFree(g.d_gather_tbl);
        if (g.rng) ggnn_b200_rng_destroy(g.rng);
      }
    }
  };
  std::unique_ptr<State> st;

  std::filesystem::path part_path(uint32_t global_id) const { return st->graph_dir / ("part_" + std::to_string(global_id) + ".ggnn"); }
  static void write_file(const std::filesystem::path& path, const void* data, size_t bytes)
  {
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    f.write(static_cast<const char*>(data), static_cast<std::streamsize>(bytes));
    if (!f) throw std::runtime_error("cannot write " + path.string());
  }
  static void read_file(const std::filesystem::path& path, void* data, size_t bytes)
  {
    std::ifstream f(path, std::ios::binary);
    f.read(static_cast<char*>(data), static_cast<std::streamsize>(bytes));
    if (!f) throw std::runtime_error("cannot read " + path.string());
  }
  bool has_graph() const
  {
    if (st->shards.empty()) return false;
    for (const auto& sh : st->shards)
      if (!sh.has_graph) return false;
    return true;
  }

  ggnn_b200_query_params query_params(const Slot& slot, const float* dq, uint32_t KQuery, float tau_query, uint32_t max_iterations,
                                      DistanceMeasure measure) const
  {
    const State& s = *st;
    const char* blob = reinterpret_cast<const char*>(slot.blob.data());
    ggnn_b200_query_params p{};
    p.D = s.cfg.D; p.measure = static_cast<int>(measure); p.KQuery = KQuery;
    p.tau_query = tau_query; p.max_iterations = max_iterations;
    p.N_base = static_cast<int32_t>(s.cfg.N); p.KBuild = s.cfg.KBuild; p.num_starting_points = s.cfg.S;
    p.d_base = slot.base.data(); p.d_query = dq;  // (callers override both for native uint8 rows)
    p.d_graph = reinterpret_cast<const KeyT*>(blob + s.off.graph);
    p.d_starting_points = reinterpret_cast<const KeyT*>(blob + s.off.translation) + s.cfg.STs_offsets[GGNN_B200_L - 1];
    p.d_nn1_stats = reinterpret_cast<const ValueT*>(blob + s.off.nn1_stats);
    p.shards_per_gpu = 1;
    return p;
  }

  /// fp32 rows of a slot; the rows of a natively stored uint8 base are widened on the device on demand (exact: the
  /// reference computes on static_cast<float>(value), distance.cuh:104-148)
  const float* rows_f32(Gpu& g, Slot& slot)
  {
    if (!slot.base.data() && slot.base_u8.data()) {
      slot.base = Dataset<float>::emptyOnGPU(slot.base_u8.N, slot.base_u8.D, g.id);
      detail::abi_check(ggnn_b200_widen_u8(slot.base_u8.data(), slot.base.data(), slot.base_u8.numel(), g.stream));
    }
    return slot.base.data();
  }
  /// release the widened copy of a natively stored uint8 shard again (GGNN_B200_KEEP_WIDENED=1 keeps it)
  void drop_f32(Gpu& g, Slot& slot)
  {
    if (slot.base_u8.data() && slot.base.data() && !std::getenv("GGNN_B200_KEEP_WIDENED")) {
      cudaStreamSynchronize(g.stream);
      slot.base = Dataset<float>{};
    }
  }
  /// does the traversal kernel read this slot's 1-byte rows natively for this query?  (include/ggnn_b200.h: D in {32, 64, 96, 128, 256},
  /// KQuery <= 47)
  bool native_u8(const Slot& slot, const GenericDataset& query, uint32_t KQuery) const
  {
    const uint32_t D = st->cfg.D;
    return slot.base_u8.data() && query.type == DataType::UINT8 && (D == 32 || D == 64 || D == 96 || D == 128 || D == 256) && KQuery <= 47 &&
           !std::getenv("GGNN_B200_NO_NATIVE_U8");
  }

  /// the device buffers holding `sh` (current device = g.id).  Resident mode: its own slot.  Swap mode
  /// (gpu_instance.cu:370-467 swapOutPart / swapInPart): the least recently used slot is written back -- graph blob to
  /// pinned host memory up to the CPU memory limit, else to part_<id>.ggnn -- and refilled from the host base + graph;
  /// everything is ordered on the GPU's one stream.
  Slot& acquire(Gpu& g, Shard& sh)
  {
    State& s = *st;
    if (sh.slot >= 0) {
      g.slots[sh.slot].last_use = ++g.clock;
      return g.slots[sh.slot];
    }
    size_t victim = 0;
    for (size_t i = 0; i < g.slots.size(); ++i) {
      if (g.slots[i].resident < 0) { victim = i; break; }
      if (g.slots[i].last_use < g.slots[victim].last_use) victim = i;
    }
    Slot& slot = g.slots[victim];
    detail::cuda_check(cudaStreamSynchronize(g.stream), "cudaStreamSynchronize");  // the victim's kernels are done
    if (slot.resident >= 0) {
      Shard& old = s.shards[static_cast<size_t>(slot.resident)];
      if (old.dirty) {
        if (!old.host_blob.data() && (s.host_blobs + 1) * s.blob_bytes <= s.cpu_memory_limit) {
          old.host_blob = Dataset<uint8_t>::empty(s.blob_bytes, 1, true);
          ++s.host_blobs;
        }
        if (old.host_blob.data()) {
          detail::cuda_check(cudaMemcpy(old.host_blob.data(), slot.blob.data(), s.blob_bytes, cudaMemcpyDeviceToHost), "cudaMemcpy(swap out)");
        }
        else {
          std::vector<uint8_t> h(s.blob_bytes);
          detail::cuda_check(cudaMemcpy(h.data(), slot.blob.data(), s.blob_bytes, cudaMemcpyDeviceToHost), "cudaMemcpy(swap out)");
          std::filesystem::create_directories(s.graph_dir);
          write_file(part_path(old.global_id), h.data(), h.size());
          old.on_disk = true;
        }
        old.dirty = false;
      }
      old.slot = -1;
    }
    // swap in: base rows from the host base, graph from host memory / disk (nothing yet before build)
    const uint64_t n_shard = s.cfg.N;
    Dataset<float> rows = float_on_gpu(*s.base, g.id, g.stream, static_cast<uint64_t>(sh.global_id) * n_shard, n_shard, &slot.base);
    (void)rows;
    if (sh.host_blob.data()) {
      detail::cuda_check(cudaMemcpyAsync(slot.blob.data(), sh.host_blob.data(), s.blob_bytes, cudaMemcpyHostToDevice, g.stream), "cudaMemcpyAsync(swap in)");
    }
    else if (sh.on_disk) {
      std::vector<uint8_t> h(s.blob_bytes);
      read_file(part_path(sh.global_id), h.data(), h.size());
      detail::cuda_check(cudaMemcpy(slot.blob.data(), h.data(), h.size(), cudaMemcpyHostToDevice), "cudaMemcpy(swap in)");
    }
    detail::cuda_check(cudaStreamSynchronize(g.stream), "cudaStreamSynchronize");
    slot.resident = sh.global_id;
    slot.last_use = ++g.clock;
    sh.slot = static_cast<int>(victim);
    return slot;
  }

  static Results to_host(Results r)
  {
    Results h;
    h.ids = Dataset<KeyT>::empty(r.ids.N, r.ids.D, true);
    h.dists = Dataset<ValueT>::empty(r.dists.N, r.dists.D, true);
    detail::cuda_check(cudaMemcpy(h.ids.data(), r.ids.data(), r.ids.size_bytes(), cudaMemcpyDeviceToHost), "cudaMemcpy(ids)");
    detail::cuda_check(cudaMemcpy(h.dists.data(), r.dists.data(), r.dists.size_bytes(), cudaMemcpyDeviceToHost), "cudaMemcpy(dists)");
    return h;
  }

  void check_base(const GenericDataset& b) const
  {
    if (!st->shards.empty()) throw std::runtime_error("The base cannot be changed after the graph has been set up.");
    if (b.type != DataType::FLOAT && b.type != DataType::UINT8) throw std::runtime_error("unsupported datatype for base");  // ggnn.cu:456-491
    if (b.D < MIN_D || b.D > MAX_D) throw std::out_of_range("unsupported dimension");
  }
  void check_query(const GenericDataset& q, const char* what) const
  {
    if (q.type != DataType::FLOAT && q.type != DataType::UINT8) throw std::runtime_error(what);
    // the reference CHECK-aborts here (ggnn.cu:524-540)
    if (q.type != st->base->type) throw std::runtime_error("query data type does not match base data type");
    if (q.D != st->base->D) throw std::out_of_range("query dimension does not match the base");
  }

  /// a uint8 dataset as it is on `gpu` (referenced when already there)
  static Dataset<uint8_t> u8_on_gpu(const GenericDataset& src, int gpu, cudaStream_t stream)
  {
    GenericDataset rows = src.reference();
    if (rows.isGPUAccessible() && (rows.gpu_id == gpu || rows.location == DataLocation::MANAGED)) return Dataset<uint8_t>{std::move(rows)};
    Dataset<uint8_t> d = Dataset<uint8_t>::emptyOnGPU(src.N, src.D, gpu);
    detail::cuda_check(cudaMemcpyAsync(d.data(), static_cast<const void*>(rows), rows.required_size_bytes(), cudaMemcpyDefault, stream), "cudaMemcpyAsync(uint8 query)");
    return d;
  }

  /// rows [from, from + num) of `src` (float or uint8, anywhere) as fp32 on `gpu`; uint8 is widened on the device.
  /// Already-resident float data is referenced, not copied -- unless `into` names the buffer to fill (swap slots).
  /// The current device must be `gpu`.
  static Dataset<float> float_on_gpu(const GenericDataset& src, int gpu, cudaStream_t stream, uint64_t from = 0,
                                     uint64_t num = std::numeric_limits<uint64_t>::max(), Dataset<float>* into = nullptr)
  {
    if (num == std::numeric_limits<uint64_t>::max()) num = src.N - from;
    GenericDataset rows = src.referenceRange(from, num);
    const bool resident = rows.isGPUAccessible() && (rows.gpu_id == gpu || rows.location == DataLocation::MANAGED);
    auto target = [&]() {
      if (into) return Dataset<float>{into->referenceRange(0, num)};
      return Dataset<float>::emptyOnGPU(num, src.D, gpu);
    };
    if (rows.type == DataType::FLOAT) {
      if (resident && !into) return Dataset<float>{std::move(rows)};
      Dataset<float> d = target();
      detail::cuda_check(cudaMemcpyAsync(d.data(), static_cast<const void*>(rows), rows.required_size_bytes(), cudaMemcpyDefault, stream), "cudaMemcpyAsync(float rows)");
      return d;
    }
    Dataset<float> d = target();
    Dataset<uint8_t> staged;
    const uint8_t* u8 = static_cast<const uint8_t*>(static_cast<const void*>(rows));
    if (!resident) {
      staged = Dataset<uint8_t>::emptyOnGPU(num, src.D, gpu);
      detail::cuda_check(cudaMemcpyAsync(staged.data(), u8, rows.required_size_bytes(), cudaMemcpyDefault, stream), "cudaMemcpyAsync(uint8 rows)");
      u8 = staged.data();
    }
    detail::abi_check(ggnn_b200_widen_u8(u8, d.data(), rows.numel(), stream));
    detail::cuda_check(cudaStreamSynchronize(stream), "cudaStreamSynchronize");  // `staged` is freed on return
    return d;
  }

  // src/ggnn/base/ggnn.cu:154-203 (partitioning), gpu_instance.cu:135-227 (how many shards fit on a GPU)
  void prepare(uint32_t KBuild)
  {
    State& s = *st;
    if (!s.base || !static_cast<const void*>(*s.base)) throw std::runtime_error("The base needs to be set before building a graph.");
    if (KBuild < MIN_KBUILD || KBuild > MAX_KBUILD) throw std::out_of_range("KBuild out of range");
    if (!s.shards.empty()) {
      if (s.cfg.KBuild != KBuild) throw std::runtime_error("graph already set up with a different KBuild");
      return;
    }
    const uint64_t N = s.base->N;
    const uint64_t n_shard = s.N_shard ? s.N_shard : N;
    if (N % n_shard) throw std::out_of_range("The base size needs to be divisible by the shard size.");
    const uint64_t num_shards = N / n_shard;
    if (num_shards % s.gpu_ids.size()) throw std::out_of_range("The number of shards needs to be divisible by the number of GPUs.");
    s.spg = static_cast<uint32_t>(num_shards / s.gpu_ids.size());
    detail::abi_check(ggnn_b200_graph_config_init(&s.cfg, static_cast<uint32_t>(n_shard), s.base->D, KBuild));
    ggnn_b200_graph_blob_offsets(&s.cfg, &s.off);
    s.blob_bytes = s.off.total;
    const size_t shard_bytes = n_shard * s.base->D * sizeof(float) + s.blob_bytes;  // (uint8 bases: upper bound)
    const size_t scratch_bytes = ggnn_b200_build_scratch_bytes(&s.cfg);
    s.shards.resize(num_shards);
    s.gpus.resize(s.gpu_ids.size());
    for (uint32_t gi = 0; gi < s.gpu_ids.size(); ++gi) {
      Gpu& g = s.gpus[gi];
      g.id = s.gpu_ids[gi];
      g.first_shard = gi * s.spg;
      detail::DeviceGuard guard(g.id);
      detail::cuda_check(cudaStreamCreate(&g.stream), "cudaStreamCreate");
      detail::cuda_check(cudaMalloc(reinterpret_cast<void**>(&g.work_counter), 16), "cudaMalloc");
      detail::cuda_check(cudaMalloc(reinterpret_cast<void**>(&g.d_gather_tbl), 16), "cudaMalloc");
      // a base that already lives on this GPU stays there (referenced, never swapped)
      const bool base_on_gpu = s.base->isGPUAccessible() && (s.base->gpu_id == g.id || s.base->location == DataLocation::MANAGED);
      uint32_t n_buf = s.spg;
      if (!base_on_gpu) {
        size_t free_b = 0, total_b = 0;
        detail::cuda_check(cudaMemGetInfo(&free_b, &total_b), "cudaMemGetInfo");
        const size_t need = s.reserved_gpu_memory + scratch_bytes;
        const size_t fit = free_b > need ? (free_b - need) / shard_bytes : 0;
        if (const char* e = std::getenv("GGNN_B200_GPU_SHARD_BUFFERS"); e && std::atoi(e) > 0) n_buf = std::min<uint32_t>(s.spg, std::atoi(e));
        else n_buf = static_cast<uint32_t>(std::min<size_t>(s.spg, fit));
        if (n_buf < 1) throw std::runtime_error("not enough GPU memory for a single shard (base + graph + build scratch); use a smaller shard size");
      }
      g.swap = n_buf < s.spg;
      g.slots.resize(n_buf);
      for (uint32_t b = 0; b < n_buf; ++b) {
        Slot& slot = g.slots[b];
        slot.blob = Dataset<uint8_t>::emptyOnGPU(s.blob_bytes, 1, g.id);
        if (g.swap) slot.base = Dataset<float>::emptyOnGPU(n_shard, s.base->D, g.id);
      }
      for (uint32_t i = 0; i < s.spg; ++i) {
        Shard& sh = s.shards[g.first_shard + i];
        sh.global_id = g.first_shard + i;
        if (!g.swap) {  // resident: shard i <-> slot i for good
          Slot& slot = g.slots[i];
          if (s.base->type == DataType::UINT8) {  // 1-byte rows stay 1-byte rows on the device
            GenericDataset rows = s.base->referenceRange(static_cast<uint64_t>(sh.global_id) * n_shard, n_shard);
            const bool on_gpu = rows.isGPUAccessible() && (rows.gpu_id == g.id || rows.location == DataLocation::MANAGED);
            if (on_gpu) slot.base_u8 = Dataset<uint8_t>{std::move(rows)};
            else {
              slot.base_u8 = Dataset<uint8_t>::emptyOnGPU(n_shard, s.base->D, g.id);
              detail::cuda_check(cudaMemcpyAsync(slot.base_u8.data(), static_cast<const void*>(rows), rows.required_size_bytes(),
                                                 cudaMemcpyDefault, g.stream), "cudaMemcpyAsync(uint8 rows)");
            }
          }
          else slot.base = float_on_gpu(*s.base, g.id, g.stream, static_cast<uint64_t>(sh.global_id) * n_shard, n_shard);
          slot.resident = sh.global_id;
          sh.slot = static_cast<int>(i);
        }
      }
      detail::cuda_check(cudaStreamSynchronize(g.stream), "cudaStreamSynchronize");
    }
    // several GPUs: can they all store into the first one's memory?  (fused shard-merge exchange)
    s.peer_gather = false;
    if (s.gpus.size() > 1 && !std::getenv("GGNN_B200_NO_PEER_GATHER")) {
      s.peer_gather = true;
      for (size_t gi = 1; gi < s.gpus.size(); ++gi) {
        detail::DeviceGuard guard(s.gpus[gi].id);
        if (ggnn_b200_peer_enable(s.gpus[0].id) != 0) s.peer_gather = false;
      }
    }
  }
};

}  // namespace ggnn
