// Page-granular radix tree for the RadixAttention prefix cache (host-side, C++).
//
// Same semantics as sglang_b200/mem_cache/radix_cache.py (the executable specification) and the reference's
// protocol (python/sglang/srt/mem_cache/radix_cache.py:377-705; the reference's own C++ tree lives in
// mem_cache/cpp_radix_tree/tree_v2.h:15-56).  Nodes live in a flat vector and are addressed by int32 ids,
// keys/values are contiguous int32 arrays, children are indexed by the bytes of their first page.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace sglb_rt {

enum class EvictPolicy : int { LRU = 0, LFU = 1, FIFO = 2, MRU = 3, FILO = 4, PRIORITY = 5 };

struct Node {
  int32_t parent = -1;
  std::unordered_map<std::string, int32_t> children;
  std::vector<int32_t> key;     // token ids (len == pages.size() * page_size)
  std::vector<int32_t> pages;   // device page ids
  std::vector<int32_t> host_pages;
  bool evicted = false;         // device copy dropped (host tier keeps it)
  int32_t lock_ref = 0;
  uint64_t last_access = 0, created = 0;
  int64_t hit_count = 0;
  int32_t priority = 0;
  bool alive = false;
};

class RadixTree {
 public:
  RadixTree(int page_size, int policy) : ps_(page_size), policy_((EvictPolicy)policy) { reset(); }
  void reset();
  // returns matched page ids; *last_node receives the deepest matched node id
  std::vector<int32_t> match_prefix(const int32_t* tokens, int64_t n, int64_t extra_key, int32_t* last_node);
  int64_t insert(const int32_t* tokens, int64_t n_tokens, const int32_t* pages, int64_t n_pages, int64_t extra_key,
                 int32_t priority);
  void inc_lock_ref(int32_t node);
  void dec_lock_ref(int32_t node);
  int64_t lock_cost(int32_t node) const;
  // evict >= num_pages pages of unlocked leaves; returns the freed device page ids
  std::vector<int32_t> evict(int64_t num_pages);
  int64_t evictable_pages() const { return evictable_; }
  int64_t protected_pages() const { return protected_; }
  std::vector<int32_t> all_pages() const;
  int64_t num_nodes() const;
  int page_size() const { return ps_; }

 private:
  int32_t new_node();
  int32_t root_for(int64_t extra_key);
  std::string child_key(const int32_t* p) const { return std::string(reinterpret_cast<const char*>(p), sizeof(int32_t) * ps_); }
  int64_t match_pages(const std::vector<int32_t>& a, const int32_t* b, int64_t nb) const;
  int32_t split(int32_t child, int64_t m_pages);
  double policy_value(const Node& n) const;

  int ps_;
  EvictPolicy policy_;
  std::vector<Node> nodes_;
  std::vector<int32_t> free_ids_;
  std::unordered_map<int64_t, int32_t> roots_;
  int64_t evictable_ = 0, protected_ = 0;
  uint64_t tick_ = 0;
};

}  // namespace sglb_rt
