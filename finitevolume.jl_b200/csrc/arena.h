// arena.h -- the device-memory arena of one handle (host-side bookkeeping only; plain C++).
//
// Why not cudaMallocAsync: a step of the hot path (assemble -> solve) allocates and frees ~40 arrays
// between 4 bytes and 7.5 GB.  On the stream-ordered pool the same sequence, repeated, cost anything
// between 0 and 1.3 s of device idle time per step at 512^3 (the pool remaps physical memory to
// defragment its virtual ranges, at random points of the sequence: profiles/r1b_step_diag.md).
// Here every request is rounded to 256 bytes and served best-fit from the free blocks of a few big
// cudaMalloc chunks; a request that fits nowhere gets a chunk of its own size (at least kMinChunk).
// A repeated sequence of requests therefore finds, from the second pass on, a free block of exactly
// its size: no device call at all in steady state, same addresses every pass.
//
// Ordering contract (same as cudaFreeAsync on one stream): a block may be handed out again right after
// free(), so all work touching it must have been enqueued on the handle's single stream before the
// free, and all work of the next owner is enqueued on that stream after the alloc.  One handle = one
// owner thread, so there is no locking.
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <set>
#include <unordered_map>
#include <vector>

namespace fvb {

class Arena {
 public:
  using ChunkAlloc = void *(*)(size_t bytes);  // nullptr on failure
  using ChunkFree = void (*)(void *);
  static constexpr size_t kAlign = 256;
  static constexpr size_t kMinChunk = size_t(64) << 20;

  Arena(ChunkAlloc a, ChunkFree f) : chunk_alloc_(a), chunk_free_(f) {}
  ~Arena() { release_all(); }
  Arena(const Arena &) = delete;
  Arena &operator=(const Arena &) = delete;

  void *alloc(size_t bytes) {
    const size_t need = round_up(bytes ? bytes : 1);
    auto it = by_size_.lower_bound({need, nullptr});  // smallest free block that fits
    if (it == by_size_.end()) {
      if (!grow(need)) {
        // out of device memory: give back the chunks nobody uses and try once more
        if (trim() == 0 || !grow(need)) return nullptr;
      }
      it = by_size_.lower_bound({need, nullptr});
      if (it == by_size_.end()) return nullptr;
    }
    char *p = it->second;
    const size_t have = it->first;
    by_size_.erase(it);
    by_addr_.erase(p);
    if (have > need) insert_free(p + need, have - need);
    used_[p] = need;
    in_use_ += need;
    return p;
  }

  // false if p was not handed out by this arena
  bool free(void *ptr) {
    if (!ptr) return true;
    char *p = static_cast<char *>(ptr);
    auto u = used_.find(p);
    if (u == used_.end()) return false;
    size_t size = u->second;
    used_.erase(u);
    in_use_ -= size;
    // coalesce with the free neighbours inside the same chunk
    auto nxt = by_addr_.find(p + size);
    if (nxt != by_addr_.end() && !chunk_bases_.count(nxt->first)) {
      size += nxt->second;
      by_size_.erase({nxt->second, nxt->first});
      by_addr_.erase(nxt);
    }
    if (!chunk_bases_.count(p)) {
      auto prv = by_addr_.lower_bound(p);
      if (prv != by_addr_.begin()) {
        --prv;
        if (prv->first + prv->second == p) {
          p = prv->first;
          size += prv->second;
          by_size_.erase({prv->second, prv->first});
          by_addr_.erase(prv);
        }
      }
    }
    insert_free(p, size);
    return true;
  }

  // Give chunks that are entirely free back to the driver; returns the bytes released.
  size_t trim() {
    size_t released = 0;
    for (size_t i = 0; i < chunks_.size();) {
      auto f = by_addr_.find(chunks_[i].base);
      if (f != by_addr_.end() && f->second == chunks_[i].size) {
        by_size_.erase({f->second, f->first});
        by_addr_.erase(f);
        chunk_bases_.erase(chunks_[i].base);
        chunk_free_(chunks_[i].base);
        released += chunks_[i].size;
        reserved_ -= chunks_[i].size;
        chunks_[i] = chunks_.back();
        chunks_.pop_back();
      } else {
        ++i;
      }
    }
    return released;
  }

  void release_all() {
    for (auto &c : chunks_) chunk_free_(c.base);
    chunks_.clear();
    chunk_bases_.clear();
    by_addr_.clear();
    by_size_.clear();
    used_.clear();
    reserved_ = in_use_ = 0;
  }

  size_t reserved() const { return reserved_; }   // bytes held from the driver
  size_t in_use() const { return in_use_; }       // bytes handed out
  size_t chunks() const { return chunks_.size(); }
  size_t live_blocks() const { return used_.size(); }

 private:
  struct Chunk { char *base; size_t size; };
  static size_t round_up(size_t b) { return (b + kAlign - 1) / kAlign * kAlign; }

  bool grow(size_t need) {
    const size_t size = need > kMinChunk ? need : kMinChunk;
    char *base = static_cast<char *>(chunk_alloc_(size));
    if (!base) return false;
    chunks_.push_back({base, size});
    chunk_bases_.insert(base);
    reserved_ += size;
    insert_free(base, size);
    return true;
  }
  void insert_free(char *p, size_t size) {
    by_addr_[p] = size;
    by_size_.insert({size, p});
  }

  ChunkAlloc chunk_alloc_;
  ChunkFree chunk_free_;
  std::vector<Chunk> chunks_;
  std::set<char *> chunk_bases_;
  std::map<char *, size_t> by_addr_;                 // free blocks by address (coalescing)
  std::set<std::pair<size_t, char *>> by_size_;      // free blocks by (size, address) (best fit)
  std::unordered_map<char *, size_t> used_;          // live blocks
  size_t reserved_ = 0, in_use_ = 0;
};

}  // namespace fvb
