// Native HTTP/1.1 front end: one epoll I/O thread owns every socket (accept, parse, keep-alive, chunked/SSE writes, disconnect detection);
// Python only sees complete requests and hands back response bytes.  This is the in-process C++ counterpart of the reference's embedded
// Rust server (rust/sglang-server: axum API server inside the scheduler process) — the socket path never holds the GIL.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

namespace sglb_rt {

struct HttpEvent {
  int type = 0;  // 0 = request, 1 = client went away before the response finished, 2 = WebSocket message (method = "text" | "binary"), 3 = WebSocket closed
  uint64_t id = 0;
  std::string method, target, version, peer;
  std::vector<std::pair<std::string, std::string>> headers;
  std::string body;
};

struct HttpStats {
  uint64_t accepted = 0, requests = 0, static_hits = 0, bad_requests = 0, disconnects = 0, bytes_in = 0, bytes_out = 0, open_conns = 0;
};

class HttpFrontend {
 public:
  HttpFrontend(const std::string& host, int port, int backlog, int64_t max_body, bool reuse_port);
  ~HttpFrontend();
  int port() const { return port_; }
  int notify_fd() const { return nfd_; }      // eventfd, readable while events are queued: lets an asyncio loop poll without a helper thread
  void start();
  void stop(double drain_seconds);       // stop accepting; wait for in-flight responses up to the deadline; close everything
  // responses answered by the I/O thread itself (health probes, model info): no Python, no GIL
  void set_static(const std::string& path, int status, const std::string& content_type, const std::string& body);
  void remove_static(const std::string& path);
  // Python side
  std::vector<HttpEvent> poll(int timeout_ms, int max_events);
  bool respond(uint64_t id, int status, const std::vector<std::pair<std::string, std::string>>& headers, const std::string& body);
  bool start_stream(uint64_t id, int status, const std::vector<std::pair<std::string, std::string>>& headers);
  bool send_chunk(uint64_t id, const std::string& data);
  bool end_stream(uint64_t id);
  // WebSocket (RFC 6455): accept answers the upgrade request `id` with 101 and switches the connection to frame mode; messages then arrive as
  // events of type 2 (fragments reassembled, pings answered by the I/O thread) and the peer's close as type 3
  bool ws_accept(uint64_t id, const std::vector<std::pair<std::string, std::string>>& headers);
  bool ws_send(uint64_t id, bool binary, const std::string& payload);
  bool ws_close(uint64_t id, int code);
  HttpStats stats() const;

 private:
  struct Conn {
    int fd = -1;
    std::string peer, in;
    std::deque<std::string> out;
    size_t out_off = 0;
    uint64_t cur = 0;          // request being answered (0 = idle, parsing allowed)
    bool streaming = false, keep_alive = true, close_after = false, want_write = false, sent_continue = false;
    bool ws = false;            // upgraded: `in` holds WebSocket frames, `cur` stays the id of the upgrade request
    std::string ws_key, ws_msg; // Sec-WebSocket-Key of the pending upgrade; fragments of the message being reassembled
    int ws_opcode = 0;
  };
  struct Cmd {
    int op;  // 0 respond, 1 start, 2 chunk, 3 end, 4 ws accept, 5 ws frame, 6 ws close
    uint64_t id;
    std::string data;          // fully formatted bytes to append to the connection's output
  };
  void loop();
  void on_accept();
  void on_readable(Conn& c);
  void try_parse(Conn& c);
  void parse_ws(Conn& c);
  void flush(Conn& c);
  void close_conn(int fd, bool notify);
  void finish_response(Conn& c);
  void apply(Cmd& cmd);
  void push_cmd(Cmd&& c);
  void update_interest(Conn& c);
  void simple_reply(Conn& c, int status, const std::string& text, bool close);

  std::string host_;
  int port_ = 0, backlog_ = 1024, lfd_ = -1, efd_ = -1, wfd_ = -1, nfd_ = -1;
  void post_event(HttpEvent&& e);              // (mu_ held by the caller)
  int64_t max_body_;
  bool reuse_port_;
  std::thread th_;
  std::atomic<bool> running_{false}, accepting_{false};
  std::unordered_map<int, Conn> conns_;
  std::unordered_map<uint64_t, int> req_fd_;          // I/O thread only
  uint64_t next_id_ = 1;
  mutable std::mutex mu_;                              // events_, cmds_, statics_, stats_
  std::condition_variable cv_;
  std::deque<HttpEvent> events_;
  std::deque<Cmd> cmds_;
  std::unordered_set<uint64_t> live_;                  // requests handed to Python whose response is not complete yet
  bool live_empty_ = true;
  struct Static { int status; std::string ctype, body; };
  std::unordered_map<std::string, Static> statics_;
  HttpStats stats_;
};

}  // namespace sglb_rt
