// graph.cu -- host-side graph parameters and blob layout.
// Replaces ggnn::GraphConfig (src/ggnn/base/graph_config.cpp:39-98) and the ggnn::Graph blob layout
// (include/ggnn/base/graph.h:38-55, src/ggnn/base/graph.cpp:33-92); byte-compatible with the
// reference's part_<id>.ggnn files.
#include "host_util.h"
#include "../../include/ggnn_b200.h"

#include <cmath>
#include <cstring>

using namespace g200;

extern "C" int ggnn_b200_graph_config_init(ggnn_b200_graph_config* c, uint32_t N, uint32_t D, uint32_t KBuild)
{
  if (!c) return set_error(GGNN_B200_ERR_INVALID, "null config");
  // include/ggnn/base/ggnn.cuh:48-52, src/ggnn/base/ggnn.cu:163-183
  if (D < 1 || D > 4096) return set_error(GGNN_B200_ERR_INVALID, "D must be in [1, 4096]");
  if (KBuild < 2 || KBuild > 512) return set_error(GGNN_B200_ERR_INVALID, "KBuild must be in [2, 512]");
  memset(c, 0, sizeof(*c));
  c->N = N;
  c->D = D;
  c->KBuild = KBuild;
  c->KF = KBuild / 2;
  c->S = next_multiple32(c->KF + 1);
  if (N < c->S) return set_error(GGNN_B200_ERR_INVALID, "N must be at least one segment (S) large");
  constexpr int L = GGNN_B200_L;
  // the growth factor is picked between floor and ceil of (N/S)^(1/3) by the resulting layer-0
  // segment size (graph_config.cpp:66-87); float math as in the reference
  const float growth = std::pow(static_cast<float>(N) / static_cast<float>(c->S), 1.f / (L - 1));
  const uint32_t Gf = static_cast<uint32_t>(growth);
  const uint32_t Gc = Gf + 1;
  const float S0f = static_cast<float>(N) / std::pow(static_cast<float>(Gf), L - 1.0f);
  const float S0c = static_cast<float>(N) / std::pow(static_cast<float>(Gc), L - 1.0f);
  const bool is_floor = (static_cast<uint32_t>(S0c) < KBuild) ||
                        (std::abs(S0f - static_cast<float>(c->S)) < std::abs(S0c - static_cast<float>(c->S)));
  c->G = is_floor ? Gf : Gc;
  c->S0 = is_floor ? static_cast<uint32_t>(S0f) : static_cast<uint32_t>(S0c);
  c->S0_off = N - c->G * c->G * c->G * c->S0;
  c->SG = c->S / c->G;
  c->SG_off = c->S - c->SG * c->G;

  uint32_t B = 1;
  for (int l = L - 1; l >= 0; --l, B *= c->G) {
    c->Bs[l] = B;
    c->Ns[l] = B * c->S;
  }
  c->Ns[0] = N;
  c->Ns_offsets[0] = 0;
  c->STs_offsets[0] = 0;
  c->STs_offsets[1] = 0;
  c->Ns_offsets[1] = N;
  for (int l = 2; l < L; ++l) {
    c->Ns_offsets[l] = c->Ns_offsets[l - 1] + c->Ns[l - 1];
    c->STs_offsets[l] = c->STs_offsets[l - 1] + c->Ns[l - 1];
  }
  c->N_all = c->Ns_offsets[L - 1] + c->Ns[L - 1];
  c->ST_all = c->STs_offsets[L - 1] + c->Ns[L - 1];
  return 0;
}

extern "C" void ggnn_b200_graph_blob_offsets(const ggnn_b200_graph_config* c, ggnn_b200_graph_offsets* o)
{
  o->graph = 0;
  o->translation = static_cast<size_t>(c->N_all) * c->KBuild * 4;
  o->selection = o->translation + static_cast<size_t>(c->ST_all) * 4;
  o->nn1_stats = o->translation + static_cast<size_t>(c->ST_all) * 8;
  o->total = align8(static_cast<size_t>(c->N_all) * c->KBuild * 4) + 2 * align8(static_cast<size_t>(c->ST_all) * 4) + 8;
}

extern "C" size_t ggnn_b200_graph_blob_bytes(const ggnn_b200_graph_config* c)
{
  ggnn_b200_graph_offsets o;
  ggnn_b200_graph_blob_offsets(c, &o);
  return o.total;
}
