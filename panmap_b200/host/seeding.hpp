// seeding.hpp -- host-side mirror of /root/reference/src/seeding.hpp for the place stage.
//   rollingSyncmers(seq, k, s, open, t, returnAll) == seeding::rollingSyncmers (seeding.hpp:126-127): vector of
//   (hash, isReverse, isSyncmer, startPos); the syncmers come from the GPU kernel (pm_rolling_syncmers), the
//   returnAll=true filler tuples (UINT64_MAX,false,false,pos) are added on the host exactly as seeding.cpp:196-225 does.
#pragma once
#include <cstdint>
#include <string_view>
#include <tuple>
#include <vector>

namespace seeding {
std::vector<std::tuple<size_t, bool, bool, int64_t>> rollingSyncmers(std::string_view seq, int k, int s, bool open, int t = 0,
                                                                     bool returnAll = true, int device = 0);
}
