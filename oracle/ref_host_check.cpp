// TEST INFRASTRUCTURE ONLY.  Drives the HOST-side functions of the UNMODIFIED reference (oracle/_ref/libggnn_ref.so,
// built by oracle/build_ref.sh from /root/reference) that belong to the hot path's result handling and need no GPU:
//   ggnn::ResultMerger<int32_t, float>::merge   (src/ggnn/base/result_merger.cpp:51-149)  -- SURVEY 8(a) A19
//   ggnn::Evaluator<int32_t, float>             (src/ggnn/base/eval.cpp:88-242)           -- SURVEY 8(a) A20
// tools/gen_host_golden.py feeds it seeded inputs and commits inputs' seeds + outputs as tests/golden/host_merge_eval.npz,
// against which the oracle's restatements (orc_merge_results, orc_eval) and the product's Evaluator are pinned.
//
//   ref_host_check merge <in.bin> <out.bin>
//     in : u32 num_gpus, spg, N_query, KQuery, N_shard; per GPU: int32 ids[N_query][KQuery * spg], float dists[same]
//     out: int32 ids[N_query][KQuery], float dists[N_query][KQuery]
//   ref_host_check eval <in.bin> <out.bin>
//     in : u32 N, N_query, D, K_gt, KQuery, measure, is_uint8; base[N][D], query[N_query][D] (float or uint8);
//          int32 gt[N_query][K_gt]; int32 results[N_query][KQuery]
//     out: float c1, c1_dup, cK, cK_dup, rK, rK_dup; u32 top1DuplicateEnd[N_query]; u32 topKDuplicateEnd[N_query]
//   ref_host_check store <in.bin> <file>   /   ref_host_check load <file> <from> <num> <out.bin>
//     Dataset<T>::store and GenericDataset::load (src/ggnn/base/dataset.cu:118-226): the fvecs / bvecs / ivecs files
//     of SURVEY 8(f) rank 3
#include <ggnn/base/dataset.cuh>
#include <ggnn/base/eval.h>
#include <ggnn/base/result_merger.h>

#include <cuda_runtime_api.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using KeyT = int32_t;
using ValueT = float;

// The reference checks the CUDA error state before EVERY allocation, plain malloc included (src/ggnn/base/data.cu:97-103),
// and a machine without a GPU answers that check with cudaErrorInsufficientDriver.  This program only ever allocates
// host memory, so it answers the one runtime call involved itself (the executable's definition takes precedence over
// libcudart's for the reference library as well); the reference sources stay untouched.
extern "C" cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }

static std::vector<unsigned char> slurp(const char* path)
{
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::perror(path); std::exit(2); }
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<unsigned char> buf(static_cast<size_t>(n));
  if (n && std::fread(buf.data(), 1, buf.size(), f) != buf.size()) { std::perror("fread"); std::exit(2); }
  std::fclose(f);
  return buf;
}

struct Reader {
  const unsigned char* p;
  const unsigned char* end;
  template <typename T>
  const T* take(size_t count)
  {
    const T* r = reinterpret_cast<const T*>(p);
    p += count * sizeof(T);
    if (p > end) { std::fprintf(stderr, "input too short\n"); std::exit(2); }
    return r;
  }
  uint32_t u32() { return *take<uint32_t>(1); }
};

static int run_merge(const char* in, const char* out)
{
  const auto buf = slurp(in);
  Reader r{buf.data(), buf.data() + buf.size()};
  const uint32_t num_gpus = r.u32(), spg = r.u32(), Nq = r.u32(), K = r.u32(), N_shard = r.u32();
  // (the constructor allocates pinned memory, which needs a CUDA device: the fields are filled by hand instead)
  ggnn::ResultMerger<KeyT, ValueT> m;
  m.N_query = Nq;
  m.KQuery = K;
  m.num_gpus = num_gpus;
  m.num_shards_per_gpu = spg;
  const size_t per = static_cast<size_t>(Nq) * K * spg;
  for (uint32_t g = 0; g < num_gpus; ++g) {
    ggnn::Results<KeyT, ValueT> part{ggnn::Dataset<KeyT>::empty(Nq, K * spg), ggnn::Dataset<ValueT>::empty(Nq, K * spg)};
    std::memcpy(part.ids.data(), r.take<KeyT>(per), per * sizeof(KeyT));
    std::memcpy(part.dists.data(), r.take<ValueT>(per), per * sizeof(ValueT));
    m.partial_results_per_gpu.emplace_back(std::move(part));
  }
  const ggnn::Results<KeyT, ValueT> res = std::move(m).merge(N_shard);
  FILE* f = std::fopen(out, "wb");
  if (!f) { std::perror(out); return 2; }
  std::fwrite(res.ids.data(), sizeof(KeyT), static_cast<size_t>(Nq) * K, f);
  std::fwrite(res.dists.data(), sizeof(ValueT), static_cast<size_t>(Nq) * K, f);
  std::fclose(f);
  return 0;
}

static int run_eval(const char* in, const char* out)
{
  const auto buf = slurp(in);
  Reader r{buf.data(), buf.data() + buf.size()};
  const uint32_t N = r.u32(), Nq = r.u32(), D = r.u32(), Kgt = r.u32(), K = r.u32(), measure = r.u32(), is_u8 = r.u32();
  ggnn::GenericDataset base, query;
  if (is_u8) {
    base = ggnn::GenericDataset{ggnn::Dataset<uint8_t>::copy({r.take<uint8_t>(static_cast<size_t>(N) * D), static_cast<size_t>(N) * D}, D)};
    query = ggnn::GenericDataset{ggnn::Dataset<uint8_t>::copy({r.take<uint8_t>(static_cast<size_t>(Nq) * D), static_cast<size_t>(Nq) * D}, D)};
  }
  else {
    base = ggnn::GenericDataset{ggnn::Dataset<float>::copy({r.take<float>(static_cast<size_t>(N) * D), static_cast<size_t>(N) * D}, D)};
    query = ggnn::GenericDataset{ggnn::Dataset<float>::copy({r.take<float>(static_cast<size_t>(Nq) * D), static_cast<size_t>(Nq) * D}, D)};
  }
  ggnn::Dataset<KeyT> gt = ggnn::Dataset<KeyT>::copy({r.take<KeyT>(static_cast<size_t>(Nq) * Kgt), static_cast<size_t>(Nq) * Kgt}, Kgt);
  ggnn::Dataset<KeyT> results = ggnn::Dataset<KeyT>::copy({r.take<KeyT>(static_cast<size_t>(Nq) * K), static_cast<size_t>(Nq) * K}, K);
  ggnn::Evaluator<KeyT, ValueT> ev{base, query, gt, K, static_cast<ggnn::DistanceMeasure>(measure)};
  const ggnn::Evaluation e = ev.evaluateResults(results);
  const float vals[6] = {e.c1, e.c1_dup, e.cKQuery, e.cKQuery_dup, e.rKQuery, e.rKQuery_dup};
  FILE* f = std::fopen(out, "wb");
  if (!f) { std::perror(out); return 2; }
  std::fwrite(vals, sizeof(float), 6, f);
  std::fwrite(ev.gt_duplicates.top1DuplicateEnd.data(), sizeof(uint32_t), ev.gt_duplicates.top1DuplicateEnd.size(), f);
  std::fwrite(ev.gt_duplicates.topKDuplicateEnd.data(), sizeof(uint32_t), ev.gt_duplicates.topKDuplicateEnd.size(), f);
  std::fclose(f);
  return 0;
}

// store: in = u32 type (0 float, 1 uint8, 2 int32), N, D; data[N][D]  ->  Dataset<T>::store(<out file>)   (dataset.cu:215-226)
static int run_store(const char* in, const char* out)
{
  const auto buf = slurp(in);
  Reader r{buf.data(), buf.data() + buf.size()};
  const uint32_t type = r.u32(), N = r.u32(), D = r.u32();
  const size_t n = static_cast<size_t>(N) * D;
  if (type == 0) ggnn::Dataset<float>::copy({r.take<float>(n), n}, D).store(out);
  else if (type == 1) ggnn::Dataset<uint8_t>::copy({r.take<uint8_t>(n), n}, D).store(out);
  else ggnn::Dataset<int32_t>::copy({r.take<int32_t>(n), n}, D).store(out);
  return 0;
}

// load: GenericDataset::load(<in file>.fvecs|.bvecs|.ivecs, from, num)  ->  out = u32 N, D, bytes per element; data
// (dataset.cu:118-213)
static int run_load(const char* in, uint32_t from, uint32_t num, const char* out)
{
  const ggnn::GenericDataset d = ggnn::GenericDataset::load(in, from, num);
  const uint32_t hdr[3] = {static_cast<uint32_t>(d.N), d.D, static_cast<uint32_t>(d.element_size())};
  FILE* f = std::fopen(out, "wb");
  if (!f) { std::perror(out); return 2; }
  std::fwrite(hdr, sizeof(uint32_t), 3, f);
  std::fwrite(d.reinterpret<unsigned char>().data(), 1, d.required_size_bytes(), f);
  std::fclose(f);
  return 0;
}

int main(int argc, char** argv)
{
  if (argc == 6 && std::string(argv[1]) == "load")
    return run_load(argv[2], static_cast<uint32_t>(std::stoul(argv[3])), static_cast<uint32_t>(std::stoul(argv[4])), argv[5]);
  if (argc != 4) {
    std::fprintf(stderr, "usage: %s merge|eval|store <in.bin> <out>  |  load <file> <from> <num> <out.bin>\n", argv[0]);
    return 2;
  }
  const std::string mode = argv[1];
  if (mode == "merge") return run_merge(argv[2], argv[3]);
  if (mode == "eval") return run_eval(argv[2], argv[3]);
  if (mode == "store") return run_store(argv[2], argv[3]);
  std::fprintf(stderr, "unknown mode %s\n", argv[1]);
  return 2;
}
