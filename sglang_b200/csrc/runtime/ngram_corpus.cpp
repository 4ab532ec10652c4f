#include "ngram_corpus.h"

#include <algorithm>
#include <queue>

namespace sglb_rt {

void NgramCorpus::reset() {
  std::lock_guard<std::mutex> g(mu_);
  nodes_.clear();
  nodes_.emplace_back();  // root
}

void NgramCorpus::insert(const int32_t* tokens, int64_t n) {
  std::lock_guard<std::mutex> g(mu_);
  // every position starts an n-gram of up to max_depth tokens
  for (int64_t i = 0; i < n; ++i) {
    int32_t cur = 0;
    const int64_t end = std::min<int64_t>(n, i + max_depth_);
    for (int64_t j = i; j < end; ++j) {
      auto it = nodes_[cur].next.find(tokens[j]);
      int32_t nxt;
      if (it == nodes_[cur].next.end()) {
        if ((int64_t)nodes_.size() >= max_nodes_) break;
        nxt = (int32_t)nodes_.size();
        nodes_.emplace_back();
        nodes_[cur].next[tokens[j]] = nxt;
      } else {
        nxt = it->second;
      }
      nodes_[nxt].count += 1;
      cur = nxt;
    }
  }
}

int32_t NgramCorpus::find_suffix_node(const int32_t* context, int64_t n, int min_match, int* matched) const {
  // longest suffix (<= max_depth-1 tokens, leaving room for at least one continuation) present in the trie
  const int longest = (int)std::min<int64_t>(n, max_depth_ - 1);
  for (int len = longest; len >= std::max(1, min_match); --len) {
    int32_t cur = 0;
    bool ok = true;
    for (int64_t j = n - len; j < n; ++j) {
      auto it = nodes_[cur].next.find(context[j]);
      if (it == nodes_[cur].next.end()) { ok = false; break; }
      cur = it->second;
    }
    if (ok && !nodes_[cur].next.empty()) { *matched = len; return cur; }
  }
  *matched = 0;
  return -1;
}

std::vector<int32_t> NgramCorpus::match_chain(const int32_t* context, int64_t n, int max_draft, int min_match) const {
  std::lock_guard<std::mutex> g(mu_);
  std::vector<int32_t> out;
  std::vector<int32_t> ctx(context, context + n);
  while ((int)out.size() < max_draft) {
    int matched = 0;
    int32_t node = find_suffix_node(ctx.data(), (int64_t)ctx.size(), min_match, &matched);
    if (node < 0) break;
    // follow the most frequent continuation as deep as the trie goes
    bool advanced = false;
    while ((int)out.size() < max_draft && !nodes_[node].next.empty()) {
      int32_t best_tok = -1, best_node = -1;
      int64_t best_cnt = -1;
      for (auto& kv : nodes_[node].next)
        if (nodes_[kv.second].count > best_cnt || (nodes_[kv.second].count == best_cnt && kv.first < best_tok)) {
          best_cnt = nodes_[kv.second].count; best_tok = kv.first; best_node = kv.second;
        }
      out.push_back(best_tok);
      ctx.push_back(best_tok);
      node = best_node;
      advanced = true;
    }
    if (!advanced) break;
  }
  return out;
}

void NgramCorpus::match_tree(const int32_t* context, int64_t n, int max_nodes, int branch, int min_match,
                             std::vector<int32_t>* tokens, std::vector<int32_t>* parents) const {
  std::lock_guard<std::mutex> g(mu_);
  tokens->clear();
  parents->clear();
  int matched = 0;
  int32_t start = find_suffix_node(context, n, min_match, &matched);
  if (start < 0) return;
  std::queue<std::pair<int32_t, int32_t>> q;  // (trie node, index in output or -1)
  q.emplace(start, -1);
  while (!q.empty() && (int)tokens->size() < max_nodes) {
    auto [node, parent_idx] = q.front();
    q.pop();
    std::vector<std::pair<int64_t, int32_t>> kids;  // (-count, token)
    for (auto& kv : nodes_[node].next) kids.emplace_back(-nodes_[kv.second].count, kv.first);
    std::sort(kids.begin(), kids.end());
    for (int i = 0; i < (int)kids.size() && i < branch && (int)tokens->size() < max_nodes; ++i) {
      const int32_t tok = kids[i].second;
      tokens->push_back(tok);
      parents->push_back(parent_idx);
      q.emplace(nodes_[node].next.at(tok), (int32_t)tokens->size() - 1);
    }
  }
}

// ------------------------------------------------------------------------------------------------ suffix automaton
static const int32_t kSep = -1;

void SuffixAutomaton::reset() {
  std::lock_guard<std::mutex> g(mu_);
  st_.clear();
  st_.emplace_back();
  text_.clear();
  last_ = 0;
}

int32_t SuffixAutomaton::get(const State& s, int32_t tok) const {
  auto it = std::lower_bound(s.next.begin(), s.next.end(), std::make_pair(tok, INT32_MIN));
  return (it != s.next.end() && it->first == tok) ? it->second : -1;
}

void SuffixAutomaton::set(State& s, int32_t tok, int32_t to) {
  auto it = std::lower_bound(s.next.begin(), s.next.end(), std::make_pair(tok, INT32_MIN));
  if (it != s.next.end() && it->first == tok) it->second = to;
  else s.next.insert(it, std::make_pair(tok, to));
}

void SuffixAutomaton::extend(int32_t tok) {
  text_.push_back(tok);
  int32_t cur = (int32_t)st_.size();
  st_.emplace_back();
  st_[cur].len = st_[last_].len + 1;
  st_[cur].end = (int64_t)text_.size() - 1;
  int32_t p = last_;
  while (p != -1 && get(st_[p], tok) == -1) {
    set(st_[p], tok, cur);
    p = st_[p].link;
  }
  if (p == -1) {
    st_[cur].link = 0;
  } else {
    int32_t q = get(st_[p], tok);
    if (st_[p].len + 1 == st_[q].len) {
      st_[cur].link = q;
    } else {
      int32_t cl = (int32_t)st_.size();
      st_.push_back(st_[q]);                     // clone keeps q's transitions, link and first-occurrence end
      st_[cl].len = st_[p].len + 1;
      while (p != -1 && get(st_[p], tok) == q) {
        set(st_[p], tok, cl);
        p = st_[p].link;
      }
      st_[q].link = st_[cur].link = cl;
    }
  }
  last_ = cur;
}

void SuffixAutomaton::insert(const int32_t* tokens, int64_t n, bool terminate) {
  std::lock_guard<std::mutex> g(mu_);
  if ((int64_t)st_.size() + 2 * (n + 1) > max_states_) {          // full: start over (the corpus is a cache, not a record)
    st_.clear();
    st_.emplace_back();
    text_.clear();
    last_ = 0;
  }
  for (int64_t i = 0; i < n; ++i) extend(tokens[i]);
  if (terminate) extend(kSep);
}

std::vector<int32_t> SuffixAutomaton::match_chain(const int32_t* context, int64_t n, int max_draft, int min_match, int* matched_len) const {
  std::lock_guard<std::mutex> g(mu_);
  int32_t s = 0;
  int l = 0;
  for (int64_t i = 0; i < n; ++i) {
    int32_t t = context[i];
    while (s != 0 && get(st_[s], t) == -1) {
      s = st_[s].link;
      l = st_[s].len;
    }
    int32_t to = get(st_[s], t);
    if (to != -1) {
      s = to;
      ++l;
    } else {
      s = 0;
      l = 0;
    }
  }
  if (matched_len) *matched_len = l;
  std::vector<int32_t> out;
  if (l < min_match || s == 0) return out;
  // prefer an occurrence that has something after it: walk up the suffix links while the first occurrence sits at the very end of the text
  while (s != 0 && (st_[s].end + 1 >= (int64_t)text_.size() || text_[st_[s].end + 1] == kSep) && st_[st_[s].link].len >= min_match) s = st_[s].link;
  for (int64_t p = st_[s].end + 1; p < (int64_t)text_.size() && (int)out.size() < max_draft; ++p) {
    if (text_[p] == kSep) break;
    out.push_back(text_[p]);
  }
  return out;
}

}  // namespace sglb_rt
