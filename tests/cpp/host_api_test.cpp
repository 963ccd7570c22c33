// CPU-only checks of the header-only C++ host API (include/ggnn/ggnn.hpp): dataset factories, [fbi]vecs IO,
// type erasure, Evaluator.  No CUDA call is made (pageable host memory only).  Driven by tests/test_host_logic.py:
//   host_api_test <dir>   reads <dir>/base.fvecs, query.fvecs, gt.ivecs, res.ivecs (+ .bvecs variants) written by the
//   test, prints one JSON object with what it loaded and evaluated; exit code != 0 on any failed check.
#include <ggnn/base/eval.h>
#include <ggnn/base/ggnn.cuh>

#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>

using namespace ggnn;

static int failures = 0;
#define EXPECT(cond)                                                     \
  do {                                                                   \
    if (!(cond)) {                                                       \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      ++failures;                                                        \
    }                                                                    \
  } while (0)
template <typename E, typename F>
static bool throws(F&& f)
{
  try {
    f();
  }
  catch (const E&) {
    return true;
  }
  catch (...) {
  }
  return false;
}

static std::string eval_json(const Evaluation& e)
{
  std::ostringstream os;
  os.precision(9);
  os << "{\"KQuery\": " << e.KQuery << ", \"c1\": " << e.c1 << ", \"c1_dup\": " << (std::isnan(e.c1_dup) ? -1.f : e.c1_dup)
     << ", \"cK\": " << e.cKQuery << ", \"cK_dup\": " << (std::isnan(e.cKQuery_dup) ? -1.f : e.cKQuery_dup)
     << ", \"rK\": " << e.rKQuery << ", \"rK_dup\": " << (std::isnan(e.rKQuery_dup) ? -1.f : e.rKQuery_dup) << "}";
  return os.str();
}

int main(int argc, char** argv)
{
  if (argc < 2) return 2;
  const std::filesystem::path dir{argv[1]};
  const uint32_t KQuery = argc > 2 ? static_cast<uint32_t>(std::atoi(argv[2])) : 10;

  // ---- factories, ownership, typed access ----
  std::vector<float> v(12);
  for (size_t i = 0; i < v.size(); ++i) v[i] = static_cast<float>(i);
  Dataset<float> a = Dataset<float>::copy(v, 4);
  EXPECT(a.N == 3 && a.D == 4 && a.type == DataType::FLOAT && a.location == DataLocation::CPU_MALLOC);
  EXPECT(a.isCPUAccessible() && !a.isGPUAccessible() && a.required_size_bytes() == 48 && a.element_size() == 4);
  EXPECT(a[5] == 5.f && a.at(11) == 11.f && a.size() == 12 && a.data() == static_cast<const float*>(a));
  EXPECT(throws<std::out_of_range>([&] { (void)a.at(12); }));
  EXPECT(throws<std::invalid_argument>([&] { (void)Dataset<float>::copy(v, 5); }));
  float sum = 0;
  for (float x : a) sum += x;
  EXPECT(sum == 66.f);
  Dataset<float> r = Dataset<float>::referenceCPUData(v.data(), 3, 4);
  EXPECT(r.location == DataLocation::FOREIGN_CPU && r.data() == v.data());
  GenericDataset mid = a.referenceRange(1, 2);
  EXPECT(mid.N == 2 && mid.location == DataLocation::FOREIGN_CPU && mid.access<float>()[0] == 4.f);
  EXPECT(throws<std::out_of_range>([&] { (void)a.referenceRange(2, 2); }));
  EXPECT(throws<std::invalid_argument>([&] { Dataset<int32_t> wrong{a.reference()}; }));
  Dataset<float> moved = std::move(a);
  EXPECT(moved.N == 3 && a.data() == nullptr && a.N == 0);
  Dataset<float> c = moved.clone();
  EXPECT(c.data() != moved.data() && c[7] == 7.f && c.location == DataLocation::CPU_MALLOC);
  Dataset<float> part = Dataset<float>::empty(2, 4);
  moved.copyRangeTo(1, 2, part);
  EXPECT(part[0] == 4.f && part[7] == 11.f);
  GenericDataset g = std::move(c);  // type erasure keeps the element type
  EXPECT(g.type == DataType::FLOAT && g.numel() == 12);
  Dataset<float> back{std::move(g)};
  EXPECT(back[3] == 3.f);
  back.releaseOwnership();
  EXPECT(back.location == DataLocation::FOREIGN_CPU);
  std::free(back.data());

  // ---- IO: store -> load round trips, ranges, extension dispatch ----
  moved.store(dir / "rt.fvecs");
  Dataset<float> l = Dataset<float>::load(dir / "rt.fvecs");
  EXPECT(l.N == 3 && l.D == 4 && l[11] == 11.f);
  Dataset<float> l2 = Dataset<float>::load(dir / "rt.fvecs", 1, 2);
  EXPECT(l2.N == 2 && l2[0] == 4.f);
  EXPECT(throws<std::out_of_range>([&] { (void)Dataset<float>::load(dir / "rt.fvecs", 2, 2); }));  // CHECK_EQ(N, num) there
  EXPECT(throws<std::runtime_error>([&] { (void)Dataset<float>::load(dir / "missing.fvecs"); }));
  EXPECT(throws<std::runtime_error>([&] { (void)GenericDataset::load(dir / "rt.txt"); }));

  GenericDataset base = GenericDataset::load(dir / "base.fvecs");
  GenericDataset query = GenericDataset::load(dir / "query.fvecs");
  GenericDataset base_u8 = GenericDataset::load(dir / "base.bvecs");
  GenericDataset query_u8 = GenericDataset::load(dir / "query.bvecs");
  Dataset<int32_t> gt = Dataset<int32_t>::load(dir / "gt.ivecs");
  Dataset<int32_t> res = Dataset<int32_t>::load(dir / "res.ivecs");
  EXPECT(base.type == DataType::FLOAT && base_u8.type == DataType::UINT8 && GenericDataset::load(dir / "gt.ivecs").type == DataType::INT32);
  EXPECT(base_u8.element_size() == 1 && base_u8.required_size_bytes() == base_u8.N * base_u8.D);

  // ---- GGNN front end: argument checks that need no device ----
  {
    GGNN<int32_t, float> idx;
    EXPECT(throws<std::runtime_error>([&] { idx.build(24, 0.5f); }));                  // base not set
    EXPECT(throws<std::runtime_error>([&] { (void)idx.query(query, 10, 0.5f); }));     // no graph
    EXPECT(throws<std::runtime_error>([&] { idx.store(); }));
    EXPECT(throws<std::runtime_error>([&] { idx.setBaseReference(gt); }));             // int32 base
    EXPECT(throws<std::out_of_range>([&] { idx.setGPUs(std::vector<int>{}); }));
    idx.setBaseReference(base_u8);
    EXPECT(throws<std::runtime_error>([&] { (void)idx.bfQuery(query, 10); }));         // float query on a uint8 base
    idx.setBaseReference(base);
    idx.setShardSize(7);
    EXPECT(throws<std::out_of_range>([&] { idx.build(24, 0.5f); }));                   // N % N_shard != 0
    EXPECT(throws<std::out_of_range>([&] { idx.build(1000, 0.5f); }));                 // KBuild out of range
  }

  // ---- Evaluator ----
  std::cout << "{\"base\": [" << base.N << ", " << base.D << "], \"gt\": [" << gt.N << ", " << gt.D << "]";
  for (int m = 0; m < 2; ++m) {
    const DistanceMeasure measure = m ? DistanceMeasure::Cosine : DistanceMeasure::Euclidean;
    Evaluator<int32_t, float> ev{base, query, gt, KQuery, measure};
    std::cout << ", \"float_" << m << "\": " << eval_json(ev.evaluateResults(res));
    Evaluator<int32_t, float> ev8{base_u8, query_u8, gt, KQuery, measure};
    std::cout << ", \"uint8_" << m << "\": " << eval_json(ev8.evaluateResults(res));
  }
  Evaluator<int32_t, float> nodup{GenericDataset{}, GenericDataset{}, gt, KQuery, DistanceMeasure::Euclidean};
  const Evaluation e = nodup.evaluateResults(res);
  std::cout << ", \"nodup\": " << eval_json(e) << "}" << std::endl;
  std::ostringstream os;
  os << e;
  EXPECT(os.str().find("(duplicates unknown)") != std::string::npos && os.str().find("c@" + std::to_string(KQuery)) != std::string::npos);
  Evaluator<int32_t, float> none;
  EXPECT(throws<std::runtime_error>([&] { (void)none.evaluateResults(res); }));        // no ground truth

  return failures ? 1 : 0;
}
