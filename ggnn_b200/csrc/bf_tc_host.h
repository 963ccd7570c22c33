// bf_tc_host.h -- host-side sizing helpers shared by the two tensor-core brute-force paths (bf_tc.cu, bf_i8.cu)
#pragma once
#include <stdint.h>

namespace g200 {

constexpr uint32_t TC_MAX_SPLITS = 32;  // 64 published best lists per query
// candidate capacity per query (speed only: overflowing queries are re-done by an exact scan)
uint32_t tc_cap(uint32_t Nq);
// number of base splits: fill the SMs in whole waves, within the candidate capacity
uint32_t tc_pick_splits(uint32_t q_tiles, uint32_t n_tiles, uint32_t num_sms, uint32_t cap, uint32_t K);

}  // namespace g200
