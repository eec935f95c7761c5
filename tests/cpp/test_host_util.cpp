// CPU unit test of finitevolume.jl_b200/csrc/host_util.h against the straightforward formulations.
#include <cstdio>
#include <map>
#include <random>

#include "../../finitevolume.jl_b200/csrc/host_util.h"

#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static std::vector<int64_t> ref_sort_unique(std::vector<int64_t> v) {
  std::sort(v.begin(), v.end());
  v.erase(std::unique(v.begin(), v.end()), v.end());
  return v;
}

int main() {
  std::mt19937_64 rng(3);
  // ---- sort_unique_i64 -------------------------------------------------------------------------------
  std::vector<std::vector<int64_t>> cases = {{}, {7}, {3, 3}, {5, 4}, {-9, 4, -9, 0}, {INT64_MAX, INT64_MIN, 0},
                                             {INT64_MAX - 1, INT64_MAX, INT64_MAX - 1}};
  {  // two dense planes far apart, shuffled (the middle rank of a slab partition), with and without duplicates
    std::vector<int64_t> v;
    for (int64_t i = 0; i < 4096; ++i) { v.push_back(1000000 + i); v.push_back(1000000 + 700000 + i); }
    std::shuffle(v.begin(), v.end(), rng);
    cases.push_back(v);
    for (int i = 0; i < 500; ++i) v.push_back(v[rng() % v.size()]);
    cases.push_back(v);
  }
  for (int rep = 0; rep < 200; ++rep) {  // random spans: dense (bitmap path) and scattered (fallback)
    const size_t n = 1 + rng() % 3000;
    const int64_t span = (rep % 2) ? (int64_t)(1 + rng() % (4 * n)) : (int64_t)(rng() % (1ull << 50));
    const int64_t base = (int64_t)(rng() % 2000000) - 1000000;
    std::vector<int64_t> v(n);
    for (auto &x : v) x = base + (int64_t)(rng() % (uint64_t)std::max<int64_t>(span, 1));
    if (rep % 5 == 0) std::sort(v.begin(), v.end());
    cases.push_back(v);
  }
  for (auto &c : cases) {
    std::vector<int64_t> got = c;
    fvb::sort_unique_i64(got);
    CHECK(got == ref_sort_unique(c));
  }
  // ---- dirichlet_table -------------------------------------------------------------------------------
  for (int rep = 0; rep < 300; ++rep) {
    const size_t n = rng() % 400;
    std::vector<int64_t> dn(n);
    if (rep % 3 == 0) { int64_t x = 1 + (int64_t)(rng() % 50); for (auto &d : dn) { d = x; x += 1 + (int64_t)(rng() % 3); } }  // ascending
    else for (auto &d : dn) d = 1 + (int64_t)(rng() % 300);                                                             // duplicates
    std::map<int64_t, int> last;  // node (0-based) -> last position
    for (size_t k = 0; k < n; ++k) last[dn[k] - 1] = (int)k;
    std::vector<int64_t> nodes;
    std::vector<int> slot;
    fvb::dirichlet_table(dn, nodes, slot);
    CHECK(nodes.size() == last.size() && slot.size() == last.size());
    size_t i = 0;
    for (auto &kv : last) { CHECK(nodes[i] == kv.first && slot[i] == kv.second); ++i; }
  }
  std::printf("host_util ok\n");
  return 0;
}
