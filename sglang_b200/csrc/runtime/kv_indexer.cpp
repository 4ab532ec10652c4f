#include "kv_indexer.h"

#include <algorithm>

namespace sglb {

void KvIndexer::stored(int worker, const std::vector<uint64_t>& hashes) {
  std::lock_guard<std::mutex> g(mu_);
  for (uint64_t h : hashes) {
    auto& v = owners_[h];
    if (std::find(v.begin(), v.end(), worker) == v.end()) {
      v.push_back(worker);
      ++per_worker_[worker];
    }
  }
}

void KvIndexer::removed(int worker, const std::vector<uint64_t>& hashes) {
  std::lock_guard<std::mutex> g(mu_);
  for (uint64_t h : hashes) {
    auto it = owners_.find(h);
    if (it == owners_.end()) continue;
    auto& v = it->second;
    auto p = std::find(v.begin(), v.end(), worker);
    if (p != v.end()) {
      v.erase(p);
      --per_worker_[worker];
    }
    if (v.empty()) owners_.erase(it);
  }
}

void KvIndexer::clear_worker(int worker) {
  std::lock_guard<std::mutex> g(mu_);
  for (auto it = owners_.begin(); it != owners_.end();) {
    auto& v = it->second;
    v.erase(std::remove(v.begin(), v.end(), worker), v.end());
    it = v.empty() ? owners_.erase(it) : std::next(it);
  }
  per_worker_[worker] = 0;
}

std::vector<std::pair<int, int>> KvIndexer::match(const std::vector<uint64_t>& hashes) const {
  std::lock_guard<std::mutex> g(mu_);
  std::vector<std::pair<int, int>> out;      // (worker, depth)
  std::vector<int> alive;
  for (size_t i = 0; i < hashes.size(); ++i) {
    auto it = owners_.find(hashes[i]);
    if (it == owners_.end()) break;
    if (i == 0) {
      alive = it->second;
      for (int w : alive) out.emplace_back(w, 1);
    } else {
      std::vector<int> next;
      for (int w : alive)
        if (std::find(it->second.begin(), it->second.end(), w) != it->second.end()) next.push_back(w);
      if (next.empty()) break;
      for (auto& pr : out)
        if (std::find(next.begin(), next.end(), pr.first) != next.end()) pr.second = (int)i + 1;
      alive.swap(next);
    }
  }
  return out;
}

int64_t KvIndexer::num_blocks() const {
  std::lock_guard<std::mutex> g(mu_);
  return (int64_t)owners_.size();
}

int64_t KvIndexer::worker_blocks(int worker) const {
  std::lock_guard<std::mutex> g(mu_);
  auto it = per_worker_.find(worker);
  return it == per_worker_.end() ? 0 : it->second;
}

}  // namespace sglb
