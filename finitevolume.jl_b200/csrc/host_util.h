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

// ---- partition planning of the single-process multi-GPU front end (multi_impl.h); pure functions, mirrored by
// finitevolume.jl_b200/distributed.py (slab_planes, halo_plan_from_ranges, send_destinations) and checked against it
// on the CPU (tests/test_host_logic.py::test_cpp_partition_planning_matches_python) -------------------------------------

// x-planes 1..n1 -> P contiguous slabs balanced by FREE planes: with Dirichlet end planes (and room for it) the end
// ranks take one plane more.  1-based inclusive.
inline void slab_planes(int64_t n1, int P, bool dirichlet_ends, std::vector<int64_t> &plo, std::vector<int64_t> &phi) {
  const int64_t fixed = (dirichlet_ends && n1 >= 2 + P) ? 2 : 0;
  const int64_t free_planes = n1 - fixed, base = free_planes / P, extra = free_planes % P;
  plo.assign((size_t)P, 0);
  phi.assign((size_t)P, 0);
  int64_t at = 1;
  for (int r = 0; r < P; ++r) {
    int64_t c = base + (r < extra ? 1 : 0);
    if (fixed && r == 0) ++c;
    if (fixed && r == P - 1) ++c;
    plo[(size_t)r] = at;
    phi[(size_t)r] = at + c - 1;
    at += c;
  }
}

// index of the first face node (i1,i2,i3) emits in regulargrid's list (src/grid.jl:72-108; grid.cuh: faces_before)
inline int64_t regulargrid_faces_before(int64_t n1, int64_t n2, int64_t n3, int64_t i1, int64_t i2, int64_t i3) {
  const int64_t hx = i1 < n1, hy = i2 < n2;
  const int64_t pfull = n2 * n3 + (n2 - 1) * n3 + n2 * (n3 - 1);
  return (i1 - 1) * pfull + (i2 - 1) * (hx * n3 + n3 + (n3 - 1)) + (i3 - 1) * (hx + hy + 1);
}

// What one rank needs to exchange its halo: per peer (ascending rank) how many rows it sends / receives, which local
// rows it sends (concatenated per peer), and where its first value lands in the peer's vector
// (= the peer's row count + the number of the peer's halo columns owned by lower ranks).
struct RankHaloPlan {
  std::vector<int32_t> peers;
  std::vector<int64_t> send_counts, recv_counts, send_dst;
  std::vector<int32_t> send_rows;
};

// start[r], nf[r]: 0-based first global free row and row count of rank r (ascending, contiguous);
// halo[r]: ascending distinct global free columns rank r references outside its own range.
// Returns "" or the reason the input is inconsistent.
inline const char *plan_halo_exchange(const std::vector<int64_t> &start, const std::vector<int64_t> &nf,
                                      const std::vector<std::vector<int64_t>> &halo, std::vector<RankHaloPlan> &plans) {
  const int P = (int)start.size();
  auto owner = [&](int64_t g) {
    int o = (int)(std::upper_bound(start.begin(), start.end(), g) - start.begin()) - 1;
    while (o >= 0 && nf[(size_t)o] == 0) --o;  // ranks without rows share their start with the next one
    return (o >= 0 && g < start[(size_t)o] + nf[(size_t)o]) ? o : -1;
  };
  plans.assign((size_t)P, RankHaloPlan());
  std::vector<std::vector<std::vector<int32_t>>> send((size_t)P, std::vector<std::vector<int32_t>>((size_t)P));
  std::vector<std::vector<int64_t>> recv((size_t)P, std::vector<int64_t>((size_t)P, 0));
  for (int r = 0; r < P; ++r) {
    const auto &hc = halo[(size_t)r];
    for (size_t k = 0; k < hc.size(); ++k) {
      if (k && hc[k] <= hc[k - 1]) return "halo columns are not strictly ascending";
      const int o = owner(hc[k]);
      if (o < 0 || o == r) return "a halo column has no owner among the other ranks";
      recv[(size_t)r][(size_t)o]++;
      send[(size_t)o][(size_t)r].push_back((int32_t)(hc[k] - start[(size_t)o]));
    }
  }
  for (int r = 0; r < P; ++r) {
    RankHaloPlan &pl = plans[(size_t)r];
    for (int p = 0; p < P; ++p) {
      if (p == r || (send[(size_t)r][(size_t)p].empty() && recv[(size_t)r][(size_t)p] == 0)) continue;
      pl.peers.push_back(p);
      pl.send_counts.push_back((int64_t)send[(size_t)r][(size_t)p].size());
      pl.recv_counts.push_back(recv[(size_t)r][(size_t)p]);
      pl.send_rows.insert(pl.send_rows.end(), send[(size_t)r][(size_t)p].begin(), send[(size_t)r][(size_t)p].end());
      const auto &ph = halo[(size_t)p];
      const int64_t before = std::lower_bound(ph.begin(), ph.end(), start[(size_t)r]) - ph.begin();
      pl.send_dst.push_back(nf[(size_t)p] + before);
    }
  }
  return "";
}

}  // namespace fvb
