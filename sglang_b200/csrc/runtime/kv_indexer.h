// Fleet-level KV index: which worker holds which cached prefix, keyed by chained page hashes.
//
// Workers publish BlockStored / BlockRemoved / AllBlocksCleared events carrying the chain hashes of the pages involved (a chain hash
// identifies the whole prefix up to and including that page), the router asks for the longest cached prefix of a request on every worker.
// Reference: experimental/sgl-router + sgl-kv-indexer (Rust; KV-event-driven hash-tree index over ZMQ).  Flat open hashing is enough
// here because a chain hash already encodes its ancestry.
#pragma once
#include <cstdint>
#include <mutex>
#include <unordered_map>
#include <utility>
#include <vector>

namespace sglb {

class KvIndexer {
 public:
  void stored(int worker, const std::vector<uint64_t>& hashes);
  void removed(int worker, const std::vector<uint64_t>& hashes);
  void clear_worker(int worker);
  // leading pages of `hashes` cached on each worker that has at least one: (worker, pages)
  std::vector<std::pair<int, int>> match(const std::vector<uint64_t>& hashes) const;
  int64_t num_blocks() const;
  int64_t worker_blocks(int worker) const;

 private:
  mutable std::mutex mu_;
  std::unordered_map<uint64_t, std::vector<int>> owners_;     // hash -> workers (small vectors: fleets are tens of workers)
  std::unordered_map<int, int64_t> per_worker_;
};

}  // namespace sglb
