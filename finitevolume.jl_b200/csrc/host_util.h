// host_util.h -- small host-side helpers of the multi-rank assembly (plain C++; unit-tested on the CPU by
// tests/cpp/test_host_util.cpp).  They exist because every millisecond the host spends between two kernels
// of fvb_assemble is device idle time inside the step, on every rank: at 8 GPUs a 512^3 step is ~0.8 s, and
// two std::sort calls over half a million entries were 7 % of it.
#pragma once
#include <algorithm>
#include <cstdint>
#include <utility>
#include <vector>

namespace fvb {

// v <- its distinct values in ascending order.  The halo references of a slab are one or two dense runs
// (whole planes of the neighbouring ranks), so when the value span is small next to the count a bitmap
// pass replaces the comparison sort: O(n + span/64) instead of O(n log n).
inline void sort_unique_i64(std::vector<int64_t> &v) {
  const size_t n = v.size();
  if (n < 2) return;
  int64_t lo = v[0], hi = v[0];
  bool ascending = true;
  for (size_t i = 1; i < n; ++i) {
    ascending &= v[i - 1] < v[i];
    lo = std::min(lo, v[i]);
    hi = std::max(hi, v[i]);
  }
  if (ascending) return;  // already strictly ascending: sorted and unique
  const uint64_t span = (uint64_t)hi - (uint64_t)lo + 1;  // no overflow: computed modulo 2^64, hi >= lo
  if (span != 0 && span <= 64 * (uint64_t)n + 4096) {
    std::vector<uint64_t> bits((size_t)((span + 63) / 64), 0);
    for (size_t i = 0; i < n; ++i) {
      const uint64_t d = (uint64_t)v[i] - (uint64_t)lo;
      bits[(size_t)(d >> 6)] |= uint64_t(1) << (d & 63);
    }
    size_t out = 0;
    for (size_t w = 0; w < bits.size(); ++w) {
      uint64_t x = bits[w];
      while (x) {
        const int b = __builtin_ctzll(x);
        v[out++] = (int64_t)((uint64_t)lo + (uint64_t)w * 64 + (uint64_t)b);
        x &= x - 1;
      }
    }
    v.resize(out);
    return;
  }
  std::sort(v.begin(), v.end());
  v.erase(std::unique(v.begin(), v.end()), v.end());
}

// The Dirichlet table of the multi-rank assembly: the 1-based node list `dn` as (0-based node, position)
// sorted by node, one entry per distinct node, the LAST position winning on duplicates
// (getnodei2dirichleti, src/FiniteVolume.jl:20-30, overwrites earlier entries).  Callers usually pass whole
// planes of a regular grid in ascending order: that case is a single linear pass.
inline void dirichlet_table(const std::vector<int64_t> &dn, std::vector<int64_t> &nodes, std::vector<int> &slot) {
  const size_t n = dn.size();
  nodes.clear();
  slot.clear();
  nodes.reserve(n);
  slot.reserve(n);
  bool ascending = true;
  for (size_t k = 1; k < n && ascending; ++k) ascending = dn[k - 1] < dn[k];
  if (ascending) {
    for (size_t k = 0; k < n; ++k) { nodes.push_back(dn[k] - 1); slot.push_back((int)k); }
    return;
  }
  std::vector<std::pair<int64_t, int>> pr(n);
  for (size_t k = 0; k < n; ++k) pr[k] = {dn[k] - 1, (int)k};
  std::sort(pr.begin(), pr.end());
  for (size_t k = 0; k < n; ++k) {
    if (!nodes.empty() && nodes.back() == pr[k].first) slot.back() = pr[k].second;  // last duplicate wins
    else { nodes.push_back(pr[k].first); slot.push_back(pr[k].second); }
  }
}

}  // namespace fvb
