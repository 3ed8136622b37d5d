// Chunked best-fit device allocator with coalescing (header-only; the backend is a pair of function pointers so that the
// free-list logic is unit-tested on the CPU with malloc, tests/test_abi_cpu.py).
//
// All device memory of a context is carved out of a few large chunks obtained from the backend (cudaMalloc); blocks are
// split on allocation and merged with their free neighbours on release, so the multi-GB workspaces of a sweep, whose
// sizes drift from site to site, are recycled without driver calls in steady state.  (Round 1's exact-size free list
// missed on drifting sizes and was trimmed with cudaFree storms; CUDA's stream-ordered pool, tried in round 2, showed
// run-to-run swings of 10x in allocation-heavy phases once tens of GB were resident.)  Use is stream-ordered: a block
// released by the host may be handed out again immediately because all work of a context is issued on one stream (the
// factorisation pool's streams only run between synchronisations of that stream).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <map>
#include <set>
#include <unordered_map>
#include <vector>

namespace mpsk {

struct BlockAllocator {
  typedef void* (*chunk_alloc_fn)(size_t);
  typedef void (*chunk_free_fn)(void*);
  chunk_alloc_fn chunk_alloc = nullptr;
  chunk_free_fn chunk_free = nullptr;
  size_t chunk_bytes = (size_t)8 << 30;       // growth granularity
  static constexpr size_t ALIGN = 512;

  struct Chunk { char* base; size_t size; };
  std::vector<Chunk> chunks;
  std::set<char*> chunk_starts;                // a block never merges across a chunk boundary
  std::map<char*, size_t> free_by_addr;
  std::multimap<size_t, char*> free_by_size;
  std::unordered_map<void*, size_t> live;
  size_t bytes_live = 0, bytes_reserved = 0;
  int64_t n_chunk_allocs = 0;

  static size_t round_up(size_t b) { return b < ALIGN ? ALIGN : (b + ALIGN - 1) / ALIGN * ALIGN; }

  void insert_free(char* p, size_t sz) { free_by_addr[p] = sz; free_by_size.emplace(sz, p); }
  void erase_free(char* p, size_t sz) {
    free_by_addr.erase(p);
    auto rg = free_by_size.equal_range(sz);
    for (auto it = rg.first; it != rg.second; ++it) if (it->second == p) { free_by_size.erase(it); break; }
  }

  void* alloc(size_t bytes) {
    const size_t sz = round_up(bytes);
    auto it = free_by_size.lower_bound(sz);
    if (it == free_by_size.end()) {
      // grow: a new chunk (at least chunk_bytes; an oversized request gets its own chunk)
      size_t want = sz > chunk_bytes ? sz : chunk_bytes;
      char* base = (char*)chunk_alloc(want);
      if (!base && want > sz) { want = sz; base = (char*)chunk_alloc(want); }
      if (!base) return nullptr;
      ++n_chunk_allocs;
      chunks.push_back(Chunk{base, want}); chunk_starts.insert(base); bytes_reserved += want;
      insert_free(base, want);
      it = free_by_size.lower_bound(sz);
    }
    char* p = it->second; const size_t have = it->first;
    free_by_addr.erase(p); free_by_size.erase(it);
    if (have - sz >= ALIGN) insert_free(p + sz, have - sz);
    const size_t got = have - sz >= ALIGN ? sz : have;
    live[p] = got; bytes_live += got;
    return p;
  }

  bool owns(void* p) const { return live.count(p) != 0; }

  void free(void* ptr) {
    auto it = live.find(ptr);
    if (it == live.end()) return;
    char* p = (char*)ptr; size_t sz = it->second;
    live.erase(it); bytes_live -= sz;
    // merge with the next free block (same chunk)
    auto nx = free_by_addr.find(p + sz);
    if (nx != free_by_addr.end() && !chunk_starts.count(nx->first)) { const size_t s2 = nx->second; erase_free(nx->first, s2); sz += s2; }
    // merge with the previous free block (same chunk)
    if (!chunk_starts.count(p)) {
      auto pv = free_by_addr.lower_bound(p);
      if (pv != free_by_addr.begin()) {
        --pv;
        if (pv->first + pv->second == p) { char* q = pv->first; const size_t s1 = pv->second; erase_free(q, s1); p = q; sz += s1; }
      }
    }
    insert_free(p, sz);
  }

  // return completely free chunks to the backend (called on allocation failure and at context destruction)
  void trim() {
    std::vector<Chunk> keep;
    for (const Chunk& c : chunks) {
      auto f = free_by_addr.find(c.base);
      if (f != free_by_addr.end() && f->second == c.size) {
        erase_free(c.base, c.size); chunk_starts.erase(c.base); bytes_reserved -= c.size;
        chunk_free(c.base);
      } else keep.push_back(c);
    }
    chunks.swap(keep);
  }

  // invariants (tests): free blocks are disjoint, inside chunks, never adjacent within a chunk; live + free == reserved
  bool check() const {
    size_t fr = 0; const char* prev_end = nullptr;
    for (auto& kv : free_by_addr) {
      fr += kv.second;
      if (prev_end && kv.first < prev_end) return false;
      if (prev_end && kv.first == prev_end && !chunk_starts.count(kv.first)) return false;     // should have been merged
      prev_end = kv.first + kv.second;
      bool inside = false;
      for (const Chunk& c : chunks) if (kv.first >= c.base && kv.first + kv.second <= c.base + c.size) inside = true;
      if (!inside) return false;
    }
    if (free_by_size.size() != free_by_addr.size()) return false;
    return fr + bytes_live == bytes_reserved;
  }
};

}  // namespace mpsk
