// N-gram draft corpus for speculative decoding (host C++).
// Reference: python/sglang/kernels/jit/csrc/ngram_corpus/{trie.cpp,ngram.cpp,result.cpp} (trie keyed on token
// n-grams with frequency-ranked continuations, async insert queue).  Ours: a bounded-depth counting trie stored
// in a flat node pool; `match` walks the longest suffix of the context that exists in the trie and then follows the
// most frequent continuation greedily (top-1 chain) or returns a breadth-limited tree (parents array).
#pragma once
#include <cstdint>
#include <mutex>
#include <unordered_map>
#include <utility>
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

// Online suffix automaton over the whole token stream (reference kernels/jit/csrc/ngram_corpus/suffix_automaton.cpp): unlike the trie
// its matches are not bounded by a fixed n-gram depth — the longest suffix of the context that occurred ANYWHERE earlier is found in time
// linear in the context tail, and the draft is the text that followed its first occurrence.  Sequences are separated by a token that never
// occurs in a query, so no match spans two requests.  O(1) amortised per inserted token; a full automaton is rebuilt from empty.
class SuffixAutomaton {
 public:
  explicit SuffixAutomaton(int64_t max_states) : max_states_(max_states) { reset(); }
  void reset();
  void insert(const int32_t* tokens, int64_t n, bool terminate);       // terminate: close the sequence (append the separator)
  std::vector<int32_t> match_chain(const int32_t* context, int64_t n, int max_draft, int min_match, int* matched_len) const;
  int64_t num_states() const { return (int64_t)st_.size(); }
  int64_t num_tokens() const { return (int64_t)text_.size(); }

 private:
  struct State {
    int32_t len = 0, link = -1;
    int64_t end = -1;                                   // end position (index into text_) of the first occurrence
    std::vector<std::pair<int32_t, int32_t>> next;      // sorted by token: most states have one or two transitions
  };
  int32_t get(const State& s, int32_t tok) const;
  void set(State& s, int32_t tok, int32_t to);
  void extend(int32_t tok);
  int64_t max_states_;
  int32_t last_ = 0;
  std::vector<State> st_;
  std::vector<int32_t> text_;
  mutable std::mutex mu_;
};

}  // namespace sglb_rt
