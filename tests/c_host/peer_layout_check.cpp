// Mailbox layout of the peer-memory collectives (csrc/kernels/peer.hpp): every slot of every parity is
// 8-byte aligned, inside the mailbox, and disjoint from every other slot -- for all supported rank
// counts and a few halo capacities.  Host-only (the offset functions are __host__ __device__).
#include <algorithm>
#include <cstdio>
#include <utility>
#include <vector>

#include "kernels/peer.hpp"

using namespace dcb::peer;

int main() {
  for (int size = 2; size <= kMaxRanks; ++size)
    for (long long cap : {1LL, 7LL, 1000LL, 132098LL}) {
      std::vector<std::pair<size_t, size_t>> spans;   // [begin, end)
      for (int par = 0; par < 2; ++par)
        for (int src = 0; src < size; ++src) {
          spans.push_back({ar_val_offset(size, par, src), ar_val_offset(size, par, src) + kMaxWords * sizeof(double)});
          spans.push_back({ar_flag_offset(size, par, src), ar_flag_offset(size, par, src) + 8});
        }
      for (int slot = 0; slot < size; ++slot)
        for (int par = 0; par < 2; ++par) {
          spans.push_back({halo_flag_offset(size, slot, par), halo_flag_offset(size, slot, par) + 8});
          spans.push_back({halo_data_offset(size, cap, slot, par),
                           halo_data_offset(size, cap, slot, par) + (size_t)cap * sizeof(double)});
        }
      const size_t total = mailbox_bytes(size, cap, size);
      std::sort(spans.begin(), spans.end());
      for (size_t i = 0; i < spans.size(); ++i) {
        if (spans[i].first % 8 != 0 || spans[i].second > total) { printf("bad span size=%d cap=%lld\n", size, cap); return 1; }
        if (i && spans[i].first < spans[i - 1].second) { printf("overlap size=%d cap=%lld\n", size, cap); return 1; }
      }
    }
  printf("ok\n");
  return 0;
}
