// Host-only check of SpillIndex (csrc/table.h): open addressing with backward-shift deletion against
// std::unordered_map over a long random stream of put / erase / find, including heavy collision chains.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <unordered_map>
#include <vector>

#include "../../meepoembedding_b200/csrc/table.h"

int main() {
  meepo::SpillIndex idx;
  std::unordered_map<uint64_t, meepo::SpillTuple> ref;
  std::mt19937_64 rng(12345);
  const uint64_t universe = 5000;  // small: forces re-insertion, long chains and many deletions
  for (int round = 0; round < 3; round++) {
    for (int step = 0; step < 400000; step++) {
      const uint64_t key = rng() % universe * 0x9E3779B97F4A7C15ull;  // spread, but repeatable
      const int op = (int)(rng() % 10);
      if (op < 5) {
        meepo::SpillTuple v{rng(), rng() % 1000};
        idx.put(key, v);
        ref[key] = v;
      } else if (op < 8) {
        const bool a = idx.erase(key);
        const bool b = ref.erase(key) != 0;
        if (a != b) return printf("erase mismatch at step %d\n", step), 1;
      } else {
        meepo::SpillTuple* f = idx.find(key);
        auto it = ref.find(key);
        if ((f != nullptr) != (it != ref.end())) return printf("find presence mismatch at step %d\n", step), 1;
        if (f && (f->seq != it->second.seq || f->ring_index != it->second.ring_index))
          return printf("find value mismatch at step %d\n", step), 1;
      }
      if (idx.size() != ref.size()) return printf("size mismatch at step %d\n", step), 1;
    }
    for (auto& kv : ref) {  // everything that should be there is, with the right value
      meepo::SpillTuple* f = idx.find(kv.first);
      if (!f || f->seq != kv.second.seq) return printf("final sweep mismatch\n"), 1;
    }
    if (round == 1) {
      idx.clear();
      ref.clear();
      idx.reserve(100000);
    }
  }
  // replace_all on several threads == the same puts one by one; the freed slabs are those of the older copies
  for (int threads : {1, 3, 8}) {
    std::vector<uint64_t> keys;
    std::vector<uint32_t> slabs;
    std::unordered_map<uint64_t, int> seen;
    while (keys.size() < 50000) {
      const uint64_t key = rng() % 120000 * 0x9E3779B97F4A7C15ull;
      if (seen.emplace(key, 1).second) {
        keys.push_back(key);
        slabs.push_back((uint32_t)(rng() % 1000000));
      }
    }
    std::vector<uint32_t> freed, want;
    for (size_t j = 0; j < keys.size(); j++) {
      auto it = ref.find(keys[j]);
      if (it != ref.end()) want.push_back((uint32_t)it->second.ring_index);
      ref[keys[j]] = meepo::SpillTuple{777000 + j, slabs[j]};
    }
    idx.replace_all(keys.data(), slabs.data(), 777000, keys.size(), freed, threads);
    std::sort(freed.begin(), freed.end());
    std::sort(want.begin(), want.end());
    if (freed != want) return printf("replace_all: freed slabs differ (threads=%d)\n", threads), 1;
    if (idx.size() != ref.size()) return printf("replace_all: size differs (threads=%d)\n", threads), 1;
    for (auto& kv : ref) {
      meepo::SpillTuple* f = idx.find(kv.first);
      if (!f || f->seq != kv.second.seq || f->ring_index != kv.second.ring_index)
        return printf("replace_all: content differs (threads=%d)\n", threads), 1;
    }
  }
  printf("spill index ok: %zu keys\n", idx.size());
  return 0;
}
