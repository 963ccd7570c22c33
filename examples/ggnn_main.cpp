// Minimal end-to-end use of the C++ API (same flow as the reference's README / ggnn_main example): random base and
// query vectors on the host, build, ANN query, brute-force ground truth, recall.
//   g++ -std=c++20 -Iinclude -I/usr/local/cuda/include examples/ggnn_main.cpp -Lggnn_b200 -lggnn_b200 \
//       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/ggnn_b200 -o examples/ggnn_main
#include <ggnn/base/ggnn.cuh>

#include <cstdio>
#include <random>
#include <vector>

int main()
{
  using namespace ggnn;
  const size_t N_base = 10'000, N_query = 10'000;
  const uint32_t dim = 128, KQuery = 10;
  std::vector<float> base_data(N_base * dim), query_data(N_query * dim);
  std::default_random_engine prng{};
  std::uniform_real_distribution<float> uniform{0.0f, 1.0f};
  for (float& x : base_data) x = uniform(prng);
  for (float& x : query_data) x = uniform(prng);

  GGNN<int32_t, float> ggnn{};
  Dataset<float> base = Dataset<float>::copy(base_data, dim, true);
  Dataset<float> query = Dataset<float>::copy(query_data, dim, true);
  ggnn.setBaseReference(base);
  ggnn.build(24, 0.5f);
  const auto [indices, dists] = ggnn.query(query, KQuery, 0.5f);
  const auto [gt, gt_dists] = ggnn.bfQuery(query, KQuery);

  size_t hits = 0;
  for (size_t n = 0; n < N_query; ++n)
    for (uint32_t i = 0; i < KQuery; ++i)
      for (uint32_t j = 0; j < KQuery; ++j) hits += indices[n * KQuery + i] == gt[n * KQuery + j];
  std::printf("first query: nearest base[%d] at squared distance %f (exact: base[%d] %f)\n", indices[0], dists[0], gt[0], gt_dists[0]);
  std::printf("recall@%u = %.4f\n", KQuery, static_cast<double>(hits) / static_cast<double>(N_query * KQuery));
  return hits > N_query * KQuery * 9 / 10 ? 0 : 1;
}
