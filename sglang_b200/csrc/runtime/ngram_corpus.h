// N-gram draft corpus for speculative decoding (host C++).
// Reference: python/sglang/kernels/jit/csrc/ngram_corpus/{trie.cpp,ngram.cpp,result.cpp} (trie keyed on token
// n-grams with frequency-ranked continuations, async insert queue).  Ours: a bounded-depth counting trie stored
// in a flat node pool; `match` walks the longest suffix of the context that exists in the trie and then follows the
// most frequent continuation greedily (top-1 chain) or returns a breadth-limited tree (parents array).
#pragma once
#include <cstdint>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace sglb_rt {

class NgramCorpus {
 public:
  NgramCorpus(int max_depth, int64_t max_nodes) : max_depth_(max_depth), max_nodes_(max_nodes) { reset(); }
  void reset();
  void insert(const int32_t* tokens, int64_t n);
  // chain draft: up to `max_draft` tokens following `context`; min_match = shortest suffix length accepted
  std::vector<int32_t> match_chain(const int32_t* context, int64_t n, int max_draft, int min_match) const;
  // tree draft: returns tokens and parent indices (-1 = attached to the last context token), BFS order
  void match_tree(const int32_t* context, int64_t n, int max_nodes, int branch, int min_match, std::vector<int32_t>* tokens,
                  std::vector<int32_t>* parents) const;
  int64_t num_nodes() const { return (int64_t)nodes_.size(); }

 private:
  struct TNode {
    std::unordered_map<int32_t, int32_t> next;  // token -> node
    int64_t count = 0;
  };
  int32_t find_suffix_node(const int32_t* context, int64_t n, int min_match, int* matched) const;
  int max_depth_;
  int64_t max_nodes_;
  std::vector<TNode> nodes_;
  mutable std::mutex mu_;
};

}  // namespace sglb_rt
