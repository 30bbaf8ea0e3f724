// tests/host/hostbatch_check.cpp -- CPU check of the host-side batch plumbing (capycrypt_b200/csrc/hostbatch.h):
// shard/chunk boundaries of split_items and the chunk counts of fixed and ragged batches.  Compiled with g++ against
// the CUDA headers only (nothing here touches a device); run by tests/test_hostbatch.py.
#include <cstdio>
#include <numeric>
#include "../../capycrypt_b200/csrc/hostbatch.h"
using namespace capy;

#define EXPECT(c) do { if (!(c)) { printf("FAIL line %d: %s\n", __LINE__, #c); return 1; } } while (0)

static bool covers(const std::vector<Range>& r, uint64_t i0, uint64_t i1) {
  uint64_t at = i0;
  for (const Range& x : r) {
    if (x.i0 != at || x.i1 <= x.i0) return false;
    at = x.i1;
  }
  return at == i1;
}

int main() {
  // fixed-size items: equal parts, contiguous, complete
  auto r = split_items(nullptr, 64, 0, 1000, 8, 0);
  EXPECT(r.size() == 8 && covers(r, 0, 1000));
  for (const Range& x : r) EXPECT(x.i1 - x.i0 == 125);
  // more parts than items, empty range, sub-range
  EXPECT(split_items(nullptr, 64, 0, 3, 8, 0).size() == 3);
  EXPECT(split_items(nullptr, 64, 5, 5, 4, 0).empty());
  EXPECT(covers(split_items(nullptr, 1, 10, 20, 3, 0), 10, 20));
  // ragged items are balanced by bytes (+ a per-item cost): one 1 MB item against 1000 small ones
  std::vector<uint64_t> off(1002, 0);
  off[1] = 1000000;
  for (int i = 2; i <= 1001; i++) off[i] = off[i - 1] + 1000;
  r = split_items(off.data(), 0, 0, 1001, 2, 0);
  EXPECT(r.size() == 2 && covers(r, 0, 1001) && r[0].i1 == 1);  // the big item alone is half of the bytes
  // zero-length items everywhere: falls back to the per-item cost
  std::vector<uint64_t> zeros(101, 0);
  r = split_items(zeros.data(), 0, 0, 100, 4, 200);
  EXPECT(r.size() == 4 && covers(r, 0, 100));
  // chunk counts: 8 MiB chunks, never fewer than 1024 items per chunk
  EXPECT(chunk_count(64ull << 20, 1 << 20) == 8);
  EXPECT(chunk_count(1, 1) == 1);
  EXPECT(chunk_count(1ull << 30, 2048) == 2);
  // ragged: a chunk must hold ~8192 x its longest message, so long messages mean few chunks
  std::vector<uint64_t> big(3001, 0);
  for (int i = 1; i <= 3000; i++) big[i] = big[i - 1] + (1u << 20);  // 3000 x 1 MiB = 2.9 GiB
  EXPECT(ragged_chunk_count(big.data(), 0, 3000, 0) == 1);
  std::vector<uint64_t> small(1 << 20 | 1, 0);
  for (size_t i = 1; i < small.size(); i++) small[i] = small[i - 1] + 64;  // 2^20 x 64 B
  EXPECT(ragged_chunk_count(small.data(), 0, 1 << 20, 0) == 8);
  // ---- LPT shares of a ragged batch over several devices ----
  {
    // a batch sorted by length, longest first: a contiguous split by bytes gives device 0 all the long chains
    std::vector<uint64_t> so(1 + 40 + 100000, 0);
    for (int i = 1; i <= 40; i++) so[i] = so[i - 1] + (1u << 20);
    for (size_t i = 41; i < so.size(); i++) so[i] = so[i - 1] + 400;
    const uint64_t n = so.size() - 1;
    for (size_t parts : {2, 4, 8}) {
      auto sh = lpt_shares(so.data(), n, parts, 72, 2);
      EXPECT(sh.size() == parts);
      std::vector<int> seen(n, 0);
      uint64_t max_cost = 0, sum_cost = 0, max_long = 0, min_long = ~0ull;
      for (const DeviceShare& d : sh) {
        uint64_t prev = 0, longs = 0, items = 0;
        for (const Run& r : d.runs) {
          EXPECT(r.i0 < r.i1 && r.i0 >= prev);  // increasing, non-overlapping
          prev = r.i1;
          for (uint64_t i = r.i0; i < r.i1; i++) { seen[i]++; items++; if (i < 40) longs++; }
        }
        EXPECT(items == d.items);
        max_cost = std::max(max_cost, d.cost);
        sum_cost += d.cost;
        max_long = std::max(max_long, longs);
        min_long = std::min(min_long, longs);
        EXPECT(d.runs.size() <= 64);  // a handful of copies per device, not one per message
      }
      for (uint64_t i = 0; i < n; i++) EXPECT(seen[i] == 1);  // a partition
      EXPECT(max_long - min_long <= 1);                        // the long chains are dealt out evenly
      EXPECT(max_cost * parts <= sum_cost + sum_cost / 20 + parts * 15000);  // and the total work is balanced
    }
    // uniform batch: no outliers -> plain contiguous ranges, one run per device
    auto sh = lpt_shares(small.data(), 1 << 20, 4, 136, 1);
    for (const DeviceShare& d : sh) EXPECT(d.runs.size() == 1 && d.items == (1u << 18));
    // fewer items than devices, empty batch
    EXPECT(lpt_shares(so.data(), 3, 8, 72, 2).size() == 3);
    EXPECT(lpt_shares(so.data(), 0, 4, 72, 2).size() == 1);
  }
  printf("hostbatch ok\n");
  return 0;
}
