// CPU unit test of finitevolume.jl_b200/csrc/arena.h (host-only bookkeeping; chunks come from malloc
// here, with a byte budget to provoke the out-of-memory path).  Built and run by tests/test_host_logic.py.
#include <algorithm>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../finitevolume.jl_b200/csrc/arena.h"

static size_t g_budget = size_t(1) << 40, g_held = 0, g_chunk_calls = 0;
static std::unordered_map<void *, size_t> g_sizes;
static void *chunk_alloc(size_t b) {
  if (g_held + b > g_budget) return nullptr;
  void *p = std::aligned_alloc(256, b);
  if (!p) return nullptr;
  g_held += b;
  g_sizes[p] = b;
  ++g_chunk_calls;
  return p;
}
static void chunk_free(void *p) {
  g_held -= g_sizes.at(p);
  g_sizes.erase(p);
  std::free(p);
}

#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

struct Live { char *p; size_t n; unsigned char tag; };

static bool overlaps(const std::vector<Live> &v) {
  std::vector<std::pair<char *, size_t>> s;
  for (auto &l : v) s.push_back({l.p, l.n});
  std::sort(s.begin(), s.end());
  for (size_t i = 1; i < s.size(); ++i)
    if (s[i - 1].first + s[i - 1].second > s[i].first) return true;
  return false;
}

int main() {
  using fvb::Arena;
  // 1. a repeated sequence of requests reaches a steady state after the first pass:
  //    same addresses, no new chunks (the property the hot path relies on)
  {
    Arena A(chunk_alloc, chunk_free);
    // sizes like one assemble -> solve step (scaled down), interleaved temporaries
    const size_t seq[] = {6400000, 3200000, 3200000, 1000000, 4, 500000, 500000, 6400000, 500000, 3700000, 7500000,
                          1070000, 1070000, 1070000, 1070000, 1070000, 1070000, 1070000, 4, 8, 16, 1070000};
    std::vector<std::vector<char *>> addr(4);
    size_t chunks_after_first = 0;
    for (int pass = 0; pass < 4; ++pass) {
      std::vector<char *> live;
      for (size_t i = 0; i < sizeof(seq) / sizeof(seq[0]); ++i) {
        char *p = (char *)A.alloc(seq[i]);
        CHECK(p && ((uintptr_t)p % Arena::kAlign) == 0);
        addr[pass].push_back(p);
        live.push_back(p);
        if (i % 5 == 4) { CHECK(A.free(live[live.size() - 2])); live.erase(live.end() - 2); }  // a temporary dies early
      }
      for (char *p : live) CHECK(A.free(p));
      CHECK(A.in_use() == 0 && A.live_blocks() == 0);
      if (pass == 0) chunks_after_first = g_chunk_calls;
    }
    CHECK(g_chunk_calls == chunks_after_first);   // no growth after the first pass
    CHECK(addr[1] == addr[2] && addr[2] == addr[3]);  // identical placement in steady state
    CHECK(A.trim() == A.reserved() + 0 || A.reserved() == 0);
    CHECK(A.reserved() == 0 && A.chunks() == 0);
  }
  CHECK(g_held == 0);
  // 2. random alloc/free: no overlap, contents survive, accounting is exact, everything coalesces back
  {
    Arena A(chunk_alloc, chunk_free);
    std::mt19937_64 rng(7);
    std::vector<Live> live;
    size_t expect_in_use = 0;
    for (int step = 0; step < 20000; ++step) {
      if (live.empty() || rng() % 100 < 55) {
        const size_t req = (rng() % 3 == 0) ? rng() % (8u << 20) : rng() % 4096;
        char *p = (char *)A.alloc(req);
        CHECK(p);
        const size_t n = std::max<size_t>(req, 1);  // a zero-byte request still owns a block
        unsigned char tag = (unsigned char)(rng() & 255);
        size_t touch = std::min<size_t>(n, 64);
        for (size_t k = 0; k < touch; ++k) { p[k] = (char)tag; p[n - 1 - k] = (char)tag; }
        live.push_back({p, n, tag});
        expect_in_use += (n + 255) / 256 * 256;
      } else {
        size_t i = rng() % live.size();
        Live l = live[i];
        size_t touch = std::min<size_t>(l.n, 64);
        for (size_t k = 0; k < touch; ++k) CHECK((unsigned char)l.p[k] == l.tag && (unsigned char)l.p[l.n - 1 - k] == l.tag);
        CHECK(A.free(l.p));
        expect_in_use -= (l.n + 255) / 256 * 256;
        live[i] = live.back();
        live.pop_back();
      }
      CHECK(A.in_use() == expect_in_use);
      if (step % 997 == 0) CHECK(!overlaps(live));
    }
    CHECK(!overlaps(live));
    for (auto &l : live) CHECK(A.free(l.p));
    CHECK(A.in_use() == 0);
    CHECK(!A.free((void *)0x1000));      // foreign pointer is refused, not corrupted
    CHECK(A.free(nullptr));
    const size_t res = A.reserved();
    CHECK(A.trim() == res && A.reserved() == 0);  // every chunk coalesced back into one free block
  }
  CHECK(g_held == 0);
  // 3. out of memory: fully free chunks are given back and the request retried; a request that cannot
  //    be met returns null and leaves the arena usable
  {
    g_budget = size_t(300) << 20;
    Arena A(chunk_alloc, chunk_free);
    void *a = A.alloc(size_t(100) << 20), *b = A.alloc(size_t(100) << 20);
    CHECK(a && b);
    CHECK(!A.alloc(size_t(150) << 20));          // 200 held + 150 > 300, nothing to trim
    CHECK(A.free(a));
    CHECK(!A.alloc(size_t(250) << 20));          // even after trimming a: 100 + 250 > 300
    void *c = A.alloc(size_t(150) << 20);        // trims a's chunk (100 MB, too small) and grows by 150
    CHECK(c && g_held == (size_t(250) << 20));
    CHECK(A.free(b) && A.free(c));
    g_budget = size_t(1) << 40;
  }
  CHECK(g_held == 0);
  std::printf("arena ok\n");
  return 0;
}
