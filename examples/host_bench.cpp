// host_bench -- times the C++20 host API (include/ggnn/ggnn.hpp -> C ABI -> sm_100a kernels) on a stored graph:
//   host_bench <dir> <N> <Nq> <D> <KBuild> <KQuery> <tau_query> <max_iterations> <reps>
// <dir> holds base.bin / query.bin (raw fp32 row-major) and part_0.ggnn (a graph stored by any implementation: the blob
// is byte-compatible with the reference's).  Prints one JSON line: synchronous ggnn::GGNN::query() from pinned host memory
// (H2D + traversal + D2H per call, like bench.py's e2e.sync_value through the Python API), the same with the query and
// the results resident on the GPU, and GGNN::queryAsync with two batches in flight.
#include <ggnn/base/ggnn.cuh>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

using namespace ggnn;
using Clock = std::chrono::steady_clock;

static std::vector<float> read_f32(const std::string& path, size_t count)
{
  std::vector<float> v(count);
  std::ifstream f(path, std::ios::binary);
  f.read(reinterpret_cast<char*>(v.data()), static_cast<std::streamsize>(count * sizeof(float)));
  if (!f) {
    std::fprintf(stderr, "cannot read %s\n", path.c_str());
    std::exit(2);
  }
  return v;
}

static uint32_t crc32(const void* data, size_t n)
{
  uint32_t table[256];
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    table[i] = c;
  }
  uint32_t c = 0xFFFFFFFFu;
  const unsigned char* p = static_cast<const unsigned char*>(data);
  for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

int main(int argc, char** argv)
{
  if (argc < 10) {
    std::fprintf(stderr, "usage: host_bench dir N Nq D KBuild KQuery tau_query max_iterations reps\n");
    return 2;
  }
  const std::string dir = argv[1];
  const size_t N = std::strtoull(argv[2], nullptr, 10), Nq = std::strtoull(argv[3], nullptr, 10);
  const uint32_t D = std::atoi(argv[4]), KBuild = std::atoi(argv[5]), KQuery = std::atoi(argv[6]);
  const float tau = static_cast<float>(std::atof(argv[7]));
  const uint32_t max_it = std::atoi(argv[8]);
  const int reps = std::atoi(argv[9]);

  Dataset<float> base = Dataset<float>::copy(read_f32(dir + "/base.bin", N * D), D, true);
  Dataset<float> query = Dataset<float>::copy(read_f32(dir + "/query.bin", Nq * D), D, true);
  GGNN<int32_t, float> ggnn{};
  ggnn.setWorkingDirectory(dir);
  ggnn.setBaseReference(base);
  ggnn.load(KBuild);

  auto ms_since = [](Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); };
  uint32_t crc = 0;
  for (int r = 0; r < 3; ++r) {
    auto res = ggnn.query(query, KQuery, tau, max_it);
    crc = crc32(res.ids.data(), res.ids.size_bytes());
  }
  auto t0 = Clock::now();
  std::vector<double> per_call;
  for (int r = 0; r < reps; ++r) {
    const auto t1 = Clock::now();
    auto res = ggnn.query(query, KQuery, tau, max_it);
    per_call.push_back(ms_since(t1));
  }
  const double sync_ms = ms_since(t0) / reps;
  std::sort(per_call.begin(), per_call.end());
  const double sync_median_ms = per_call[per_call.size() / 2], sync_max_ms = per_call.back();

  // two batches in flight (each: pinned H2D, traversal, D2H of the results)
  double async_ms = -1.0;
  {
    for (int r = 0; r < 2; ++r) ggnn.queryAsync(query, KQuery, tau, max_it).get();
    t0 = Clock::now();
    auto pending = ggnn.queryAsync(query, KQuery, tau, max_it);
    for (int r = 1; r < reps; ++r) {
      auto next = ggnn.queryAsync(query, KQuery, tau, max_it);
      auto res = pending.get();
      pending = std::move(next);
    }
    auto res = pending.get();
    async_ms = ms_since(t0) / reps;
    if (crc32(res.ids.data(), res.ids.size_bytes()) != crc) {
      std::fprintf(stderr, "queryAsync and query disagree\n");
      return 1;
    }
  }

  ggnn.setReturnResultsOnGPU(true);
  Dataset<float> q_gpu = Dataset<float>::emptyOnGPU(Nq, D, 0);
  query.copyTo(q_gpu);
  cudaDeviceSynchronize();
  for (int r = 0; r < 3; ++r) auto res = ggnn.query(q_gpu, KQuery, tau, max_it);
  t0 = Clock::now();
  for (int r = 0; r < reps; ++r) auto res = ggnn.query(q_gpu, KQuery, tau, max_it);
  cudaDeviceSynchronize();
  const double dev_ms = ms_since(t0) / reps;

  std::printf("{\"api\": \"ggnn::GGNN<int32_t,float> (include/ggnn/ggnn.hpp)\", \"reps\": %d, \"sync_ms_per_batch\": %.4f, "
              "\"sync_queries_per_s\": %.1f, \"sync_median_ms\": %.4f, \"sync_max_ms\": %.4f, \"async2_ms_per_batch\": %.4f, \"async2_queries_per_s\": %.1f, "
              "\"device_resident_ms_per_batch\": %.4f, \"device_resident_queries_per_s\": %.1f, \"ids_crc32\": %u}\n",
              reps, sync_ms, Nq / (sync_ms * 1e-3), sync_median_ms, sync_max_ms, async_ms, Nq / (async_ms * 1e-3), dev_ms, Nq / (dev_ms * 1e-3), crc);
  return 0;
}
