// pybind11 module `sglb_rt`: host-side C++ runtime (radix tree, page allocator, n-gram corpus, page hashing).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstring>
#include <stdexcept>

#include <thread>

#include "http_frontend.h"
#include "image_ops.h"
#include "kv_indexer.h"
#include "ngram_corpus.h"
#include "radix_tree.h"

namespace py = pybind11;
using namespace sglb_rt;
using i32arr = py::array_t<int32_t, py::array::c_style | py::array::forcecast>;

static py::array_t<int32_t> to_np(const std::vector<int32_t>& v) {
  py::array_t<int32_t> a((py::ssize_t)v.size());
  if (!v.empty()) std::memcpy(a.mutable_data(), v.data(), v.size() * sizeof(int32_t));
  return a;
}

// ---------------------------------------------------------------- LIFO page allocator (page 0 reserved)
class PageAllocator {
 public:
  explicit PageAllocator(int64_t num_pages) : n_(num_pages) { clear(); }
  void clear() {
    free_.clear();
    free_.reserve(n_);
    for (int64_t p = n_ - 1; p >= 1; --p) free_.push_back((int32_t)p);
    is_free_.assign(n_, 1);
    is_free_[0] = 0;
  }
  int64_t available() const { return (int64_t)free_.size(); }
  py::object alloc(int64_t n) {
    if (n > (int64_t)free_.size()) return py::none();
    py::array_t<int32_t> out((py::ssize_t)n);
    int32_t* o = out.mutable_data();
    for (int64_t i = 0; i < n; ++i) {
      o[i] = free_.back();
      free_.pop_back();
      is_free_[o[i]] = 0;
    }
    return std::move(out);
  }
  void free_pages(i32arr pages) {
    const int32_t* p = pages.data();
    for (py::ssize_t i = 0; i < pages.size(); ++i) {
      if (p[i] <= 0 || p[i] >= n_ || is_free_[p[i]]) throw std::runtime_error("double free / invalid free of KV page " + std::to_string(p[i]));
      is_free_[p[i]] = 1;
      free_.push_back(p[i]);
    }
  }
  bool is_free(int64_t p) const { return is_free_[p] != 0; }

 private:
  int64_t n_;
  std::vector<int32_t> free_;
  std::vector<uint8_t> is_free_;
};

// 64-bit FNV-1a chained page hash (parent hash folded in) — keys for KV events / L3 storage
// (reference: mem_cache/cpp_utils/hash_binding.cpp)
static std::vector<uint64_t> hash_pages(i32arr tokens, int page_size, uint64_t seed) {
  const int32_t* t = tokens.data();
  const int64_t n = tokens.size() / page_size;
  std::vector<uint64_t> out((size_t)n);
  uint64_t prev = seed;
  for (int64_t p = 0; p < n; ++p) {
    uint64_t h = 1469598103934665603ull ^ prev;
    const unsigned char* b = reinterpret_cast<const unsigned char*>(t + p * page_size);
    for (size_t i = 0; i < sizeof(int32_t) * (size_t)page_size; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    out[(size_t)p] = h;
    prev = h;
  }
  return out;
}

PYBIND11_MODULE(sglb_rt, m) {
  m.doc() = "sglang_b200 host runtime (C++)";
  py::class_<RadixTree>(m, "RadixTree")
      .def(py::init<int, int>(), py::arg("page_size"), py::arg("policy") = 0)
      .def("reset", &RadixTree::reset)
      .def("match_prefix", [](RadixTree& t, i32arr tokens, int64_t extra_key) {
        int32_t last = -1;
        auto pages = t.match_prefix(tokens.data(), tokens.size(), extra_key, &last);
        return py::make_tuple(to_np(pages), last);
      })
      .def("insert", [](RadixTree& t, i32arr tokens, i32arr pages, int64_t extra_key, int32_t priority) {
        return t.insert(tokens.data(), tokens.size(), pages.data(), pages.size(), extra_key, priority);
      })
      .def("inc_lock_ref", &RadixTree::inc_lock_ref)
      .def("dec_lock_ref", &RadixTree::dec_lock_ref)
      .def("lock_cost", &RadixTree::lock_cost)
      .def("evict", [](RadixTree& t, int64_t n) { return to_np(t.evict(n)); })
      .def("evictable_pages", &RadixTree::evictable_pages)
      .def("protected_pages", &RadixTree::protected_pages)
      .def("all_pages", [](RadixTree& t) { return to_np(t.all_pages()); })
      .def("num_nodes", &RadixTree::num_nodes);

  py::class_<PageAllocator>(m, "PageAllocator")
      .def(py::init<int64_t>())
      .def("clear", &PageAllocator::clear)
      .def("available", &PageAllocator::available)
      .def("alloc", &PageAllocator::alloc)
      .def("free", &PageAllocator::free_pages)
      .def("is_free", &PageAllocator::is_free);

  py::class_<NgramCorpus>(m, "NgramCorpus")
      .def(py::init<int, int64_t>(), py::arg("max_depth") = 8, py::arg("max_nodes") = 1 << 22)
      .def("reset", &NgramCorpus::reset)
      .def("insert", [](NgramCorpus& c, i32arr toks) { py::gil_scoped_release r; c.insert(toks.data(), toks.size()); })
      .def("match_chain", [](NgramCorpus& c, i32arr ctx, int max_draft, int min_match) {
        return to_np(c.match_chain(ctx.data(), ctx.size(), max_draft, min_match));
      }, py::arg("context"), py::arg("max_draft"), py::arg("min_match") = 1)
      .def("match_tree", [](NgramCorpus& c, i32arr ctx, int max_nodes, int branch, int min_match) {
        std::vector<int32_t> toks, parents;
        c.match_tree(ctx.data(), ctx.size(), max_nodes, branch, min_match, &toks, &parents);
        return py::make_tuple(to_np(toks), to_np(parents));
      }, py::arg("context"), py::arg("max_nodes"), py::arg("branch") = 2, py::arg("min_match") = 1)
      .def("num_nodes", &NgramCorpus::num_nodes);
  py::class_<SuffixAutomaton>(m, "SuffixAutomaton")
      .def(py::init<int64_t>(), py::arg("max_states") = (int64_t)1 << 23)
      .def("reset", &SuffixAutomaton::reset)
      .def("insert", [](SuffixAutomaton& c, i32arr toks, bool terminate) { py::gil_scoped_release r; c.insert(toks.data(), toks.size(), terminate); },
           py::arg("tokens"), py::arg("terminate") = true)
      .def("match_chain", [](SuffixAutomaton& c, i32arr ctx, int max_draft, int min_match) {
        int matched = 0;
        auto v = c.match_chain(ctx.data(), ctx.size(), max_draft, min_match, &matched);
        return py::make_tuple(to_np(v), matched);
      })
      .def("num_states", &SuffixAutomaton::num_states)
      .def("num_tokens", &SuffixAutomaton::num_tokens);

  py::class_<sglb::KvIndexer>(m, "KvIndexer")
      .def(py::init<>())
      .def("stored", &sglb::KvIndexer::stored, py::call_guard<py::gil_scoped_release>())
      .def("removed", &sglb::KvIndexer::removed, py::call_guard<py::gil_scoped_release>())
      .def("clear_worker", &sglb::KvIndexer::clear_worker, py::call_guard<py::gil_scoped_release>())
      .def("match", &sglb::KvIndexer::match, py::call_guard<py::gil_scoped_release>())
      .def("num_blocks", &sglb::KvIndexer::num_blocks)
      .def("worker_blocks", &sglb::KvIndexer::worker_blocks);

  m.def("hash_pages", &hash_pages, py::arg("tokens"), py::arg("page_size"), py::arg("seed") = 0);

  // ---- image preprocessing (image_ops.cpp)
  using u8arr = py::array_t<uint8_t, py::array::c_style | py::array::forcecast>;
  using f32arr = py::array_t<float, py::array::c_style | py::array::forcecast>;
  m.def("image_resize", [](u8arr img, int out_h, int out_w, int filter) {
    if (img.ndim() != 3) throw std::runtime_error("image_resize: expected HWC uint8");
    const int h = (int)img.shape(0), w = (int)img.shape(1), c = (int)img.shape(2);
    py::array_t<uint8_t> out({out_h, out_w, c});
    {
      py::gil_scoped_release r;
      resize_u8(img.data(), h, w, c, out.mutable_data(), out_h, out_w, (Filter)filter);
    }
    return out;
  }, py::arg("img"), py::arg("out_h"), py::arg("out_w"), py::arg("filter") = 3);
  m.def("clip_preprocess_batch", [](std::vector<u8arr> imgs, int size, f32arr mean, f32arr std_, int filter, int threads) {
    const int n = (int)imgs.size();
    py::array_t<float> out({n, 3, size, size});
    float* o = out.mutable_data();
    std::vector<const uint8_t*> ptr(n);
    std::vector<int> hs(n), ws(n);
    for (int i = 0; i < n; ++i) {
      if (imgs[i].ndim() != 3 || imgs[i].shape(2) != 3) throw std::runtime_error("clip_preprocess_batch: expected HxWx3 uint8 images");
      ptr[i] = imgs[i].data(); hs[i] = (int)imgs[i].shape(0); ws[i] = (int)imgs[i].shape(1);
    }
    const float* mp = mean.data();
    const float* sp = std_.data();
    {
      py::gil_scoped_release r;
      const int nt = std::max(1, std::min(threads, n));
      std::vector<std::thread> pool;
      for (int t = 0; t < nt; ++t)
        pool.emplace_back([&, t] {
          for (int i = t; i < n; i += nt) clip_preprocess(ptr[i], hs[i], ws[i], o + (size_t)i * 3 * size * size, size, mp, sp, (Filter)filter);
        });
      for (auto& th : pool) th.join();
    }
    return out;
  }, py::arg("imgs"), py::arg("size"), py::arg("mean"), py::arg("std"), py::arg("filter") = 3, py::arg("threads") = 4);
  m.def("patchify", [](f32arr chw, int patch, int merge, int temporal) {
    if (chw.ndim() != 3 || chw.shape(0) != 3) throw std::runtime_error("patchify: expected [3, H, W] float32");
    const int H = (int)chw.shape(1), W = (int)chw.shape(2);
    if (H % (patch * merge) || W % (patch * merge)) throw std::runtime_error("patchify: H and W must be multiples of patch * merge");
    py::array_t<float> out({(H / patch) * (W / patch), 3 * temporal * patch * patch});
    patchify(chw.data(), H, W, patch, merge, temporal, out.mutable_data());
    return out;
  }, py::arg("chw"), py::arg("patch"), py::arg("merge") = 1, py::arg("temporal") = 1);
  // ---- native HTTP front end (http_frontend.cpp): sockets and HTTP framing live on a C++ epoll thread, Python sees whole requests
  py::class_<HttpFrontend>(m, "HttpFrontend")
      .def(py::init<const std::string&, int, int, int64_t, bool>(), py::arg("host"), py::arg("port"), py::arg("backlog") = 2048,
           py::arg("max_body") = (int64_t)1 << 30, py::arg("reuse_port") = false)
      .def_property_readonly("port", &HttpFrontend::port)
      .def_property_readonly("notify_fd", &HttpFrontend::notify_fd)
      .def("start", &HttpFrontend::start)
      .def("stop", &HttpFrontend::stop, py::arg("drain_seconds") = 0.0, py::call_guard<py::gil_scoped_release>())
      .def("set_static", [](HttpFrontend& f, const std::string& path, int status, const std::string& ctype, py::bytes body) {
        f.set_static(path, status, ctype, (std::string)body);
      })
      .def("remove_static", &HttpFrontend::remove_static)
      .def("poll", [](HttpFrontend& f, int timeout_ms, int max_events) {
        std::vector<HttpEvent> evs;
        {
          py::gil_scoped_release rel;
          evs = f.poll(timeout_ms, max_events);
        }
        py::list out;
        for (auto& e : evs) {
          py::list hs;
          for (auto& kv : e.headers) hs.append(py::make_tuple(py::bytes(kv.first), py::bytes(kv.second)));
          out.append(py::make_tuple(e.type, e.id, e.method, py::bytes(e.target), e.version, e.peer, hs, py::bytes(e.body)));
        }
        return out;
      }, py::arg("timeout_ms") = 100, py::arg("max_events") = 64)
      .def("respond", [](HttpFrontend& f, uint64_t id, int status, std::vector<std::pair<std::string, std::string>> headers, py::bytes body) {
        return f.respond(id, status, headers, (std::string)body);
      })
      .def("start_stream", &HttpFrontend::start_stream)
      .def("send_chunk", [](HttpFrontend& f, uint64_t id, py::bytes data) { return f.send_chunk(id, (std::string)data); })
      .def("end_stream", &HttpFrontend::end_stream)
      .def("ws_accept", &HttpFrontend::ws_accept)
      .def("ws_send", [](HttpFrontend& f, uint64_t id, bool binary, py::bytes data) { return f.ws_send(id, binary, (std::string)data); })
      .def("ws_close", &HttpFrontend::ws_close)
      .def("stats", [](HttpFrontend& f) {
        HttpStats s = f.stats();
        py::dict d;
        d["accepted"] = s.accepted; d["requests"] = s.requests; d["static_hits"] = s.static_hits; d["bad_requests"] = s.bad_requests;
        d["disconnects"] = s.disconnects; d["bytes_in"] = s.bytes_in; d["bytes_out"] = s.bytes_out; d["open_conns"] = s.open_conns;
        return d;
      });
  m.def("hash_bytes", [](py::buffer b) {
    py::buffer_info info = b.request();
    return hash_bytes((const uint8_t*)info.ptr, (int64_t)info.size * info.itemsize);
  });
}
