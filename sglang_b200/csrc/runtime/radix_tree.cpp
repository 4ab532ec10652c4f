#include "radix_tree.h"

#include <algorithm>
#include <queue>
#include <stdexcept>
#include <tuple>

namespace sglb_rt {

void RadixTree::reset() {
  nodes_.clear();
  free_ids_.clear();
  roots_.clear();
  evictable_ = protected_ = 0;
  tick_ = 0;
}

int32_t RadixTree::new_node() {
  int32_t id;
  if (!free_ids_.empty()) {
    id = free_ids_.back();
    free_ids_.pop_back();
    nodes_[id] = Node();
  } else {
    id = (int32_t)nodes_.size();
    nodes_.emplace_back();
  }
  nodes_[id].alive = true;
  return id;
}

int32_t RadixTree::root_for(int64_t extra_key) {
  auto it = roots_.find(extra_key);
  if (it != roots_.end()) return it->second;
  int32_t r = new_node();
  nodes_[r].lock_ref = 1;  // roots are never evicted
  roots_[extra_key] = r;
  return r;
}

int64_t RadixTree::match_pages(const std::vector<int32_t>& a, const int32_t* b, int64_t nb) const {
  int64_t n = std::min<int64_t>((int64_t)a.size(), nb);
  n -= n % ps_;
  int64_t i = 0;
  while (i < n && a[i] == b[i]) ++i;
  return i / ps_;
}

int32_t RadixTree::split(int32_t child_id, int64_t m_pages) {
  int32_t nid = new_node();
  Node& child = nodes_[child_id];
  Node& nn = nodes_[nid];
  const int64_t mt = m_pages * ps_;
  nn.parent = child.parent;
  nn.key.assign(child.key.begin(), child.key.begin() + mt);
  nn.pages.assign(child.pages.begin(), child.pages.begin() + m_pages);
  nn.lock_ref = child.lock_ref;
  nn.last_access = child.last_access;
  nn.created = child.created;
  nn.hit_count = child.hit_count;
  nn.priority = child.priority;
  if (!child.host_pages.empty()) {
    nn.host_pages.assign(child.host_pages.begin(), child.host_pages.begin() + m_pages);
    child.host_pages.erase(child.host_pages.begin(), child.host_pages.begin() + m_pages);
  }
  nodes_[child.parent].children[child_key(nn.key.data())] = nid;
  child.key.erase(child.key.begin(), child.key.begin() + mt);
  child.pages.erase(child.pages.begin(), child.pages.begin() + m_pages);
  child.parent = nid;
  nn.children[child_key(child.key.data())] = child_id;
  return nid;
}

std::vector<int32_t> RadixTree::match_prefix(const int32_t* tokens, int64_t n, int64_t extra_key, int32_t* last_node) {
  int32_t node = root_for(extra_key);
  std::vector<int32_t> out;
  n -= n % ps_;
  const uint64_t now = ++tick_;
  int64_t off = 0;
  while (off < n) {
    auto& ch = nodes_[node].children;
    auto it = ch.find(child_key(tokens + off));
    if (it == ch.end()) break;
    int32_t c = it->second;
    if (nodes_[c].evicted) break;
    int64_t m = match_pages(nodes_[c].key, tokens + off, n - off);
    if (m == 0) break;
    if (m * ps_ < (int64_t)nodes_[c].key.size()) c = split(c, m);
    Node& cn = nodes_[c];
    cn.last_access = now;
    cn.hit_count += 1;
    out.insert(out.end(), cn.pages.begin(), cn.pages.end());
    node = c;
    off += m * ps_;
  }
  *last_node = node;
  return out;
}

int64_t RadixTree::insert(const int32_t* tokens, int64_t n_tokens, const int32_t* pages, int64_t n_pages, int64_t extra_key,
                          int32_t priority) {
  int64_t np = std::min<int64_t>(n_tokens / ps_, n_pages);
  const int64_t n = np * ps_;
  int32_t node = root_for(extra_key);
  int64_t off = 0, matched = 0;
  const uint64_t now = ++tick_;
  while (off < n) {
    auto& ch = nodes_[node].children;
    auto it = ch.find(child_key(tokens + off));
    if (it == ch.end()) break;
    int32_t c = it->second;
    int64_t m = match_pages(nodes_[c].key, tokens + off, n - off);
    if (m == 0) break;
    if (m * ps_ < (int64_t)nodes_[c].key.size()) c = split(c, m);
    Node& cn = nodes_[c];
    if (cn.evicted) {
      cn.pages.assign(pages + off / ps_, pages + off / ps_ + m);
      cn.evicted = false;
      if (cn.lock_ref == 0) evictable_ += m; else protected_ += m;
    } else {
      matched += m;
    }
    cn.last_access = now;
    cn.priority = std::max(cn.priority, priority);
    node = c;
    off += m * ps_;
  }
  if (off < n) {
    int32_t nid = new_node();
    Node& nn = nodes_[nid];
    nn.parent = node;
    nn.key.assign(tokens + off, tokens + n);
    nn.pages.assign(pages + off / ps_, pages + np);
    nn.last_access = nn.created = now;
    nn.priority = priority;
    nodes_[node].children[child_key(nn.key.data())] = nid;
    evictable_ += (int64_t)nn.pages.size();
  }
  return matched;
}

void RadixTree::inc_lock_ref(int32_t node) {
  while (node >= 0 && nodes_[node].parent >= 0) {
    Node& n = nodes_[node];
    if (n.lock_ref == 0 && !n.evicted) {
      evictable_ -= (int64_t)n.pages.size();
      protected_ += (int64_t)n.pages.size();
    }
    n.lock_ref += 1;
    node = n.parent;
  }
}

void RadixTree::dec_lock_ref(int32_t node) {
  while (node >= 0 && nodes_[node].parent >= 0) {
    Node& n = nodes_[node];
    if (n.lock_ref <= 0) throw std::runtime_error("dec_lock_ref underflow");
    n.lock_ref -= 1;
    if (n.lock_ref == 0 && !n.evicted) {
      evictable_ += (int64_t)n.pages.size();
      protected_ -= (int64_t)n.pages.size();
    }
    node = n.parent;
  }
}

int64_t RadixTree::lock_cost(int32_t node) const {
  int64_t cost = 0;
  while (node >= 0 && nodes_[node].parent >= 0 && nodes_[node].lock_ref == 0) {
    if (!nodes_[node].evicted) cost += (int64_t)nodes_[node].pages.size();
    node = nodes_[node].parent;
  }
  return cost;
}

double RadixTree::policy_value(const Node& n) const {
  switch (policy_) {
    case EvictPolicy::LRU: return (double)n.last_access;
    case EvictPolicy::LFU: return (double)n.hit_count * 1e12 + (double)n.last_access;
    case EvictPolicy::FIFO: return (double)n.created;
    case EvictPolicy::MRU: return -(double)n.last_access;
    case EvictPolicy::FILO: return -(double)n.created;
    case EvictPolicy::PRIORITY: return (double)n.priority * 1e12 + (double)n.last_access;
  }
  return 0;
}

std::vector<int32_t> RadixTree::evict(int64_t num_pages) {
  std::vector<int32_t> freed;
  if (num_pages <= 0) return freed;
  using Item = std::tuple<double, int32_t>;
  std::priority_queue<Item, std::vector<Item>, std::greater<Item>> heap;
  auto live_children = [&](const Node& n) {
    for (auto& kv : n.children)
      if (!nodes_[kv.second].evicted) return true;
    return false;
  };
  for (int32_t i = 0; i < (int32_t)nodes_.size(); ++i) {
    const Node& n = nodes_[i];
    if (!n.alive || n.parent < 0 || n.evicted || n.lock_ref > 0) continue;
    if (!live_children(n)) heap.emplace(policy_value(n), i);
  }
  while (!heap.empty() && (int64_t)freed.size() < num_pages) {
    int32_t id = std::get<1>(heap.top());
    heap.pop();
    Node& n = nodes_[id];
    if (!n.alive || n.lock_ref > 0 || n.evicted) continue;
    freed.insert(freed.end(), n.pages.begin(), n.pages.end());
    evictable_ -= (int64_t)n.pages.size();
    int32_t parent = n.parent;
    nodes_[parent].children.erase(child_key(n.key.data()));
    n.alive = false;
    n.children.clear();
    n.key.clear();
    n.pages.clear();
    free_ids_.push_back(id);
    Node& pn = nodes_[parent];
    if (pn.parent >= 0 && pn.lock_ref == 0 && !pn.evicted && !live_children(pn)) heap.emplace(policy_value(pn), parent);
  }
  return freed;
}

std::vector<int32_t> RadixTree::all_pages() const {
  std::vector<int32_t> out;
  for (auto& n : nodes_)
    if (n.alive && n.parent >= 0 && !n.evicted) out.insert(out.end(), n.pages.begin(), n.pages.end());
  return out;
}

int64_t RadixTree::num_nodes() const {
  int64_t c = 0;
  for (auto& n : nodes_) c += n.alive ? 1 : 0;
  return c;
}

}  // namespace sglb_rt
