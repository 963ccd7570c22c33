// ref_driver -- TEST INFRASTRUCTURE ONLY.
//
// Drives the UNMODIFIED reference library (oracle/_ref/libggnn_ref.so, built from
// /root/reference by oracle/build_ref.sh) through its own public API ggnn::GGNN
// (include/ggnn/base/ggnn.cuh:41-182) to (a) produce golden dumps (graph blob, query / bf_query
// ids + dists) for the parity tests and (b) time the reference CUDA path on the same box for
// bench.py --impl reference.  This file is written for this repo; it contains no reference code.
//
// usage: ref_driver key=value ...
//   dir=<workdir>            working directory: base.bin, query.bin (raw fp32 row-major) are read
//                            from it, part_0.ggnn is stored to / loaded from it, dumps go to it
//   n=<N_base> nq=<N_query> d=<D>
//   measure=0|1              0 Euclidean, 1 Cosine
//   kbuild=24 tau_build=0.5 refine=2
//   build=1|0                1: build() + store();  0: load(kbuild) from dir/part_0.ggnn
//   kquery=10 tau_query=0.64 max_iter=400
//   query_reps=R             run query() R times from pinned host memory (end to end), report each
//   gpu_reps=R               run query() R times with the query resident on the GPU and results
//                            left on the GPU (kernel + result allocation only)
//   bf=K bf_nq=M             run bfQuery(K) once on the first M queries (0 = skip; default M = nq), dump bf_ids.bin / bf_dists.bin
//   dump=1|0                 write query_ids.bin / query_dists.bin
//   gpus=G shard=N_shard     use GPUs 0..G-1 and shards of N_shard rows (default: 1 GPU, one shard)
//   batch=B                  query.bin holds nq = batches x B queries; every query() call takes the next B-query slice
//                            (round robin); default B = nq
//   build_reps=R             build R times with a fresh GGNN object each (reports every build: the first one pays the
//                            CUDA context / module load, the later ones are warm); the last graph is kept
//   sweep=t1,t2,.. sweep_iters=i1,i2,..   after the normal run: query (first batch, results on the host) for every
//                            (tau_query, max_iterations) pair, report recall@kquery against bf (needs bf=kquery) and the
//                            reference's own kernel time -- the operating-point search of examples/cpp-and-cuda/ggnn_benchmark.cpp:186-200
#include <ggnn/base/ggnn.cuh>
#include <ggnn/base/eval.h>

#include <glog/logging.h>

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

using namespace ggnn;
using Clock = std::chrono::steady_clock;

static std::map<std::string, std::string> parse(int argc, char** argv)
{
  std::map<std::string, std::string> m;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto p = a.find('=');
    if (p == std::string::npos) continue;
    m[a.substr(0, p)] = a.substr(p + 1);
  }
  return m;
}
static std::string gets(const std::map<std::string, std::string>& m, const char* k, const char* d)
{
  auto it = m.find(k);
  return it == m.end() ? d : it->second;
}
static double getd(const std::map<std::string, std::string>& m, const char* k, double d)
{
  auto it = m.find(k);
  return it == m.end() ? d : atof(it->second.c_str());
}

static std::vector<float> read_f32(const std::filesystem::path& p, size_t count)
{
  std::vector<float> v(count);
  std::ifstream f(p, std::ios::binary);
  if (!f) { fprintf(stderr, "cannot open %s\n", p.c_str()); exit(2); }
  f.read(reinterpret_cast<char*>(v.data()), count * sizeof(float));
  if (static_cast<size_t>(f.gcount()) != count * sizeof(float)) { fprintf(stderr, "short read %s\n", p.c_str()); exit(2); }
  return v;
}
template <typename T>
static void write_bin(const std::filesystem::path& p, const T* data, size_t count)
{
  std::ofstream f(p, std::ios::binary);
  f.write(reinterpret_cast<const char*>(data), count * sizeof(T));
}

// collects the reference's own "query part ... => ms: X" VLOG(0) lines (gpu_instance.cu:709-712)
struct CerrCapture {
  std::stringstream ss;
  std::streambuf* old;
  CerrCapture() : old(std::cerr.rdbuf(ss.rdbuf())) {}
  ~CerrCapture() { std::cerr.rdbuf(old); }
  std::vector<double> values(const std::string& tag)
  {
    std::vector<double> out;
    std::string s = ss.str(), line;
    std::stringstream in(s);
    while (std::getline(in, line)) {
      auto p = line.find(tag);
      if (p != std::string::npos) out.push_back(atof(line.c_str() + p + tag.size()));
    }
    return out;
  }
};

int main(int argc, char** argv)
{
  auto a = parse(argc, argv);
  const std::filesystem::path dir = gets(a, "dir", ".");
  const size_t N = static_cast<size_t>(getd(a, "n", 10000));
  const size_t Nq = static_cast<size_t>(getd(a, "nq", 10000));
  const uint32_t D = static_cast<uint32_t>(getd(a, "d", 128));
  const auto measure = getd(a, "measure", 0) ? DistanceMeasure::Cosine : DistanceMeasure::Euclidean;
  const uint32_t kbuild = static_cast<uint32_t>(getd(a, "kbuild", 24));
  const float tau_build = static_cast<float>(getd(a, "tau_build", 0.5));
  const uint32_t refine = static_cast<uint32_t>(getd(a, "refine", 2));
  const bool do_build = getd(a, "build", 1) != 0;
  const uint32_t kquery = static_cast<uint32_t>(getd(a, "kquery", 10));
  const float tau_query = static_cast<float>(getd(a, "tau_query", 0.64));
  const uint32_t max_iter = static_cast<uint32_t>(getd(a, "max_iter", 400));
  const int query_reps = static_cast<int>(getd(a, "query_reps", 1));
  const int gpu_reps = static_cast<int>(getd(a, "gpu_reps", 0));
  const uint32_t bf = static_cast<uint32_t>(getd(a, "bf", 0));
  const bool dump = getd(a, "dump", 1) != 0;

  google::SetVLOGLevel("*", 0);

  std::vector<float> base_v = read_f32(dir / "base.bin", N * D);
  std::vector<float> query_v = read_f32(dir / "query.bin", Nq * D);
  Dataset<float> base = Dataset<float>::copy(base_v, D, true);
  Dataset<float> query = Dataset<float>::copy(query_v, D, true);

  printf("{\"impl\": \"reference\", \"n\": %zu, \"nq\": %zu, \"d\": %u", N, Nq, D);

  const size_t batch = std::min(Nq, static_cast<size_t>(getd(a, "batch", static_cast<double>(Nq))));
  const size_t n_batches = std::max<size_t>(1, Nq / batch);
  const int build_reps = std::max(1, static_cast<int>(getd(a, "build_reps", 1)));
  auto configure = [&](GGNN<int32_t, float>& g) {
    g.setWorkingDirectory(dir);
    const int gpus = static_cast<int>(getd(a, "gpus", 1));
    std::vector<int> ids;
    for (int i = 0; i < gpus; ++i) ids.push_back(i);
    g.setGPUs(ids);
    const uint32_t shard = static_cast<uint32_t>(getd(a, "shard", 0));
    if (shard) g.setShardSize(shard);
    g.setBaseReference(base);
  };
  if (do_build && build_reps > 1) {  // throw-away builds: each with its own GGNN object (buffers freed in between)
    printf(", \"build_s_all\": [");
    for (int r = 0; r + 1 < build_reps; ++r) {
      GGNN<int32_t, float> tmp{};
      configure(tmp);
      const auto t0 = Clock::now();
      tmp.build(kbuild, tau_build, refine, measure);
      cudaDeviceSynchronize();
      printf("%s%.6f", r ? ", " : "", std::chrono::duration<double>(Clock::now() - t0).count());
    }
    printf("]");
  }
  GGNN<int32_t, float> ggnn{};
  configure(ggnn);

  if (do_build) {
    const auto t0 = Clock::now();
    ggnn.build(kbuild, tau_build, refine, measure);
    cudaDeviceSynchronize();
    const double s = std::chrono::duration<double>(Clock::now() - t0).count();
    ggnn.store();
    printf(", \"build_s\": %.6f", s);
  }
  else {
    ggnn.load(kbuild);
  }

  std::vector<int32_t> bf_ids;
  if (bf) {  // ground truth for the first bf_nq queries (default: all)
    const size_t bf_nq = std::min(Nq, static_cast<size_t>(getd(a, "bf_nq", static_cast<double>(Nq))));
    Dataset<float> qbf = Dataset<float>::referenceCPUData(query.data(), bf_nq, D);
    const auto t0 = Clock::now();
    auto res = ggnn.bfQuery(qbf, bf, measure);
    cudaDeviceSynchronize();
    const double s = std::chrono::duration<double>(Clock::now() - t0).count();
    printf(", \"bf_s\": %.6f, \"bf_nq\": %zu", s, bf_nq);
    write_bin(dir / "bf_ids.bin", res.ids.data(), bf_nq * bf);
    write_bin(dir / "bf_dists.bin", res.dists.data(), bf_nq * bf);
    bf_ids.assign(res.ids.data(), res.ids.data() + bf_nq * bf);
  }

  // slice b of the query set as a dataset of its own (pinned host memory, not owned)
  auto host_batch = [&](size_t b) {
    return Dataset<float>::referenceCPUData(query.data() + (b % n_batches) * batch * D, batch, D);
  };

  if (query_reps > 0) {
    printf(", \"query_e2e_ms\": [");
    std::vector<double> kernel_ms;
    std::vector<int32_t> all_ids(Nq * kquery, -1);
    std::vector<float> all_d(Nq * kquery, 0.f);
    for (int r = 0; r < query_reps; ++r) {
      CerrCapture cap;
      Dataset<float> qb = host_batch(r);
      const auto t0 = Clock::now();
      auto res = ggnn.query(qb, kquery, tau_query, max_iter, measure);
      const double ms = std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
      printf("%s%.4f", r ? ", " : "", ms);
      for (double v : cap.values("=> ms: ")) kernel_ms.push_back(v);
      if (dump) {
        const size_t off = (static_cast<size_t>(r) % n_batches) * batch * kquery;
        std::copy(res.ids.data(), res.ids.data() + batch * kquery, all_ids.begin() + off);
        std::copy(res.dists.data(), res.dists.data() + batch * kquery, all_d.begin() + off);
      }
    }
    if (dump) {  // batches never queried stay -1
      write_bin(dir / "query_ids.bin", all_ids.data(), Nq * kquery);
      write_bin(dir / "query_dists.bin", all_d.data(), Nq * kquery);
    }
    printf("], \"query_kernel_ms\": [");
    for (size_t i = 0; i < kernel_ms.size(); ++i) printf("%s%.4f", i ? ", " : "", kernel_ms[i]);
    printf("]");
  }

  if (gpu_reps > 0 && getd(a, "gpus", 1) <= 1) {  // results on the GPU are single-GPU only (ggnn.cu:299-306)
    ggnn.setReturnResultsOnGPU(true);
    Dataset<float> q_gpu = Dataset<float>::emptyOnGPU(Nq, D, 0);
    query.copyTo(q_gpu);
    cudaDeviceSynchronize();
    printf(", \"query_gpu_ms\": [");
    std::vector<double> kernel_ms;
    for (int r = 0; r < gpu_reps; ++r) {
      CerrCapture cap;
      Dataset<float> qb = Dataset<float>::referenceGPUData(q_gpu.data() + (static_cast<size_t>(r) % n_batches) * batch * D, batch, D, 0);
      const auto t0 = Clock::now();
      auto res = ggnn.query(qb, kquery, tau_query, max_iter, measure);
      cudaDeviceSynchronize();
      const double ms = std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
      printf("%s%.4f", r ? ", " : "", ms);
      for (double v : cap.values("=> ms: ")) kernel_ms.push_back(v);
    }
    printf("], \"query_gpu_kernel_ms\": [");
    for (size_t i = 0; i < kernel_ms.size(); ++i) printf("%s%.4f", i ? ", " : "", kernel_ms[i]);
    printf("]");
    ggnn.setReturnResultsOnGPU(false);
  }

  // operating-point sweep on the first batch (ggnn_benchmark.cpp:186-200 sweeps tau_query the same way)
  const std::string sweep = gets(a, "sweep", "");
  if (!sweep.empty() && bf >= kquery && bf_ids.size() >= batch * bf) {
    auto split = [](const std::string& s) {
      std::vector<double> v;
      std::stringstream in(s);
      std::string tok;
      while (std::getline(in, tok, ',')) if (!tok.empty()) v.push_back(atof(tok.c_str()));
      return v;
    };
    const std::vector<double> taus = split(sweep);
    const std::vector<double> iters = split(gets(a, "sweep_iters", "400"));
    printf(", \"sweep\": [");
    bool first = true;
    Dataset<float> qb = host_batch(0);
    for (double it : iters)
      for (double tau : taus) {
        ggnn.query(qb, kquery, static_cast<float>(tau), static_cast<uint32_t>(it), measure);  // warm
        CerrCapture cap;
        const auto t0 = Clock::now();
        auto res = ggnn.query(qb, kquery, static_cast<float>(tau), static_cast<uint32_t>(it), measure);
        const double ms = std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
        const auto km = cap.values("=> ms: ");
        size_t hits = 0;
        for (size_t n = 0; n < batch; ++n)
          for (uint32_t i = 0; i < kquery; ++i) {
            const int32_t id = res.ids.data()[n * kquery + i];
            for (uint32_t j = 0; j < kquery; ++j)
              if (bf_ids[n * bf + j] == id) { ++hits; break; }
          }
        printf("%s{\"tau_query\": %.3f, \"max_iterations\": %u, \"recall\": %.5f, \"e2e_ms\": %.4f, \"kernel_ms\": %.4f}",
               first ? "" : ", ", tau, static_cast<uint32_t>(it), static_cast<double>(hits) / (batch * kquery), ms,
               km.empty() ? -1.0 : km.back());
        first = false;
      }
    printf("]");
  }
  printf("}\n");
  fflush(stdout);
  return 0;
}
