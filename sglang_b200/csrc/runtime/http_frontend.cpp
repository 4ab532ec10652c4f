#include "http_frontend.h"

#include <arpa/inet.h>
#include <fcntl.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/epoll.h>
#include <sys/eventfd.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cstring>
#include <stdexcept>

namespace sglb_rt {

static const char* reason(int s) {
  switch (s) {
    case 100: return "Continue";
    case 200: return "OK";
    case 201: return "Created";
    case 202: return "Accepted";
    case 204: return "No Content";
    case 206: return "Partial Content";
    case 301: return "Moved Permanently";
    case 302: return "Found";
    case 304: return "Not Modified";
    case 307: return "Temporary Redirect";
    case 400: return "Bad Request";
    case 401: return "Unauthorized";
    case 403: return "Forbidden";
    case 404: return "Not Found";
    case 405: return "Method Not Allowed";
    case 408: return "Request Timeout";
    case 411: return "Length Required";
    case 413: return "Payload Too Large";
    case 422: return "Unprocessable Entity";
    case 426: return "Upgrade Required";
    case 429: return "Too Many Requests";
    case 431: return "Request Header Fields Too Large";
    case 500: return "Internal Server Error";
    case 501: return "Not Implemented";
    case 503: return "Service Unavailable";
    default: return "Status";
  }
}

static void set_nonblock(int fd) { fcntl(fd, F_SETFL, fcntl(fd, F_GETFL, 0) | O_NONBLOCK); }

static std::string lower(std::string s) {
  for (auto& ch : s) ch = (char)std::tolower((unsigned char)ch);
  return s;
}

static std::string head(int status, const std::vector<std::pair<std::string, std::string>>& headers, int64_t content_length, bool chunked,
                        bool keep_alive) {
  std::string h = "HTTP/1.1 " + std::to_string(status) + " " + reason(status) + "\r\n";
  for (auto& kv : headers) {
    std::string k = lower(kv.first);
    if (k == "content-length" || k == "transfer-encoding" || k == "connection") continue;      // framing is ours
    h += kv.first + ": " + kv.second + "\r\n";
  }
  if (chunked) h += "Transfer-Encoding: chunked\r\n";
  else if (content_length >= 0) h += "Content-Length: " + std::to_string(content_length) + "\r\n";
  h += keep_alive ? "Connection: keep-alive\r\n\r\n" : "Connection: close\r\n\r\n";
  return h;
}


// ---- WebSocket helpers (RFC 6455): SHA-1 and base64 for the accept key, server-to-client frames are never masked
static void sha1(const std::string& in, unsigned char out[20]) {
  uint32_t h0 = 0x67452301, h1 = 0xEFCDAB89, h2 = 0x98BADCFE, h3 = 0x10325476, h4 = 0xC3D2E1F0;
  std::string m = in;
  uint64_t bits = (uint64_t)in.size() * 8;
  m.push_back((char)0x80);
  while (m.size() % 64 != 56) m.push_back((char)0);
  for (int i = 7; i >= 0; --i) m.push_back((char)((bits >> (8 * i)) & 0xFF));
  auto rol = [](uint32_t v, int s) { return (v << s) | (v >> (32 - s)); };
  for (size_t off = 0; off < m.size(); off += 64) {
    uint32_t w[80];
    for (int i = 0; i < 16; ++i)
      w[i] = ((uint32_t)(unsigned char)m[off + 4 * i] << 24) | ((uint32_t)(unsigned char)m[off + 4 * i + 1] << 16) |
             ((uint32_t)(unsigned char)m[off + 4 * i + 2] << 8) | (uint32_t)(unsigned char)m[off + 4 * i + 3];
    for (int i = 16; i < 80; ++i) w[i] = rol(w[i - 3] ^ w[i - 8] ^ w[i - 14] ^ w[i - 16], 1);
    uint32_t a = h0, b = h1, c = h2, d = h3, e = h4;
    for (int i = 0; i < 80; ++i) {
      uint32_t f, k;
      if (i < 20) { f = (b & c) | (~b & d); k = 0x5A827999; }
      else if (i < 40) { f = b ^ c ^ d; k = 0x6ED9EBA1; }
      else if (i < 60) { f = (b & c) | (b & d) | (c & d); k = 0x8F1BBCDC; }
      else { f = b ^ c ^ d; k = 0xCA62C1D6; }
      uint32_t t = rol(a, 5) + f + e + k + w[i];
      e = d; d = c; c = rol(b, 30); b = a; a = t;
    }
    h0 += a; h1 += b; h2 += c; h3 += d; h4 += e;
  }
  uint32_t hs[5] = {h0, h1, h2, h3, h4};
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 4; ++j) out[4 * i + j] = (unsigned char)((hs[i] >> (24 - 8 * j)) & 0xFF);
}

static std::string b64(const unsigned char* p, size_t n) {
  static const char* T = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  std::string o;
  for (size_t i = 0; i < n; i += 3) {
    uint32_t v = (uint32_t)p[i] << 16 | (i + 1 < n ? (uint32_t)p[i + 1] << 8 : 0) | (i + 2 < n ? (uint32_t)p[i + 2] : 0);
    o.push_back(T[(v >> 18) & 63]);
    o.push_back(T[(v >> 12) & 63]);
    o.push_back(i + 1 < n ? T[(v >> 6) & 63] : '=');
    o.push_back(i + 2 < n ? T[v & 63] : '=');
  }
  return o;
}

static std::string ws_frame(int opcode, const std::string& payload) {
  std::string f;
  f.push_back((char)(0x80 | (opcode & 0x0F)));
  size_t n = payload.size();
  if (n < 126) f.push_back((char)n);
  else if (n < 65536) { f.push_back((char)126); f.push_back((char)(n >> 8)); f.push_back((char)(n & 0xFF)); }
  else { f.push_back((char)127); for (int i = 7; i >= 0; --i) f.push_back((char)((n >> (8 * i)) & 0xFF)); }
  return f + payload;
}

HttpFrontend::HttpFrontend(const std::string& host, int port, int backlog, int64_t max_body, bool reuse_port)
    : host_(host), port_(port), backlog_(backlog), max_body_(max_body), reuse_port_(reuse_port) {
  lfd_ = ::socket(AF_INET, SOCK_STREAM | SOCK_CLOEXEC, 0);
  if (lfd_ < 0) throw std::runtime_error("socket() failed");
  int one = 1;
  setsockopt(lfd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  if (reuse_port_) setsockopt(lfd_, SOL_SOCKET, SO_REUSEPORT, &one, sizeof(one));
  sockaddr_in a{};
  a.sin_family = AF_INET;
  a.sin_port = htons((uint16_t)port);
  if (host.empty() || host == "0.0.0.0") a.sin_addr.s_addr = INADDR_ANY;
  else if (inet_pton(AF_INET, host == "localhost" ? "127.0.0.1" : host.c_str(), &a.sin_addr) != 1) {
    ::close(lfd_);
    throw std::runtime_error("native HTTP front end: IPv4 literal expected for --host, got " + host);
  }
  if (::bind(lfd_, (sockaddr*)&a, sizeof(a)) < 0 || ::listen(lfd_, backlog_) < 0) {
    int e = errno;
    ::close(lfd_);
    throw std::runtime_error(std::string("bind/listen failed: ") + std::strerror(e));
  }
  socklen_t len = sizeof(a);
  getsockname(lfd_, (sockaddr*)&a, &len);
  port_ = ntohs(a.sin_port);
  set_nonblock(lfd_);
  efd_ = epoll_create1(EPOLL_CLOEXEC);
  wfd_ = eventfd(0, EFD_NONBLOCK | EFD_CLOEXEC);
  nfd_ = eventfd(0, EFD_NONBLOCK | EFD_CLOEXEC);
  epoll_event ev{};
  ev.events = EPOLLIN;
  ev.data.fd = lfd_;
  epoll_ctl(efd_, EPOLL_CTL_ADD, lfd_, &ev);
  ev.data.fd = wfd_;
  epoll_ctl(efd_, EPOLL_CTL_ADD, wfd_, &ev);
}

HttpFrontend::~HttpFrontend() {
  stop(0.0);
  if (efd_ >= 0) ::close(efd_);
  if (wfd_ >= 0) ::close(wfd_);
  if (nfd_ >= 0) ::close(nfd_);
}

void HttpFrontend::post_event(HttpEvent&& e) {
  bool was_empty = events_.empty();
  events_.push_back(std::move(e));
  if (was_empty) {
    uint64_t one = 1;
    (void)!::write(nfd_, &one, 8);
  }
}

void HttpFrontend::start() {
  if (running_.exchange(true)) return;
  accepting_ = true;
  th_ = std::thread([this] { loop(); });
}

void HttpFrontend::stop(double drain_seconds) {
  if (!running_) return;
  accepting_ = false;
  uint64_t one = 1;
  (void)!::write(wfd_, &one, 8);
  auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds((int64_t)(drain_seconds * 1000));
  while (std::chrono::steady_clock::now() < deadline) {
    {
      std::lock_guard<std::mutex> g(mu_);
      if (stats_.open_conns == 0) break;
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(10));      // (the loop closes connections as their responses complete)
  }
  running_ = false;
  (void)!::write(wfd_, &one, 8);
  if (th_.joinable()) th_.join();
  cv_.notify_all();
}

void HttpFrontend::set_static(const std::string& path, int status, const std::string& ctype, const std::string& body) {
  std::lock_guard<std::mutex> g(mu_);
  statics_[path] = Static{status, ctype, body};
}

void HttpFrontend::remove_static(const std::string& path) {
  std::lock_guard<std::mutex> g(mu_);
  statics_.erase(path);
}

HttpStats HttpFrontend::stats() const {
  std::lock_guard<std::mutex> g(mu_);
  return stats_;
}

std::vector<HttpEvent> HttpFrontend::poll(int timeout_ms, int max_events) {
  std::unique_lock<std::mutex> g(mu_);
  if (events_.empty() && timeout_ms != 0)
    cv_.wait_for(g, std::chrono::milliseconds(timeout_ms), [this] { return !events_.empty() || !running_; });
  std::vector<HttpEvent> out;
  while (!events_.empty() && (int)out.size() < max_events) {
    out.push_back(std::move(events_.front()));
    events_.pop_front();
  }
  if (events_.empty()) {          // level-triggered readers: the fd stays readable exactly while something is queued
    uint64_t v;
    (void)!::read(nfd_, &v, 8);
  }
  return out;
}

void HttpFrontend::push_cmd(Cmd&& c) {
  {
    std::lock_guard<std::mutex> g(mu_);
    cmds_.push_back(std::move(c));
  }
  uint64_t one = 1;
  (void)!::write(wfd_, &one, 8);
}

bool HttpFrontend::respond(uint64_t id, int status, const std::vector<std::pair<std::string, std::string>>& headers, const std::string& body) {
  {
    std::lock_guard<std::mutex> g(mu_);
    auto it = live_.find(id);
    if (it == live_.end()) return false;
  }
  // keep-alive is decided by the I/O thread (it knows what the client asked for): the marker byte is patched in apply()
  push_cmd(Cmd{0, id, head(status, headers, (int64_t)body.size(), false, true) + body});
  return true;
}

bool HttpFrontend::start_stream(uint64_t id, int status, const std::vector<std::pair<std::string, std::string>>& headers) {
  {
    std::lock_guard<std::mutex> g(mu_);
    if (live_.find(id) == live_.end()) return false;
  }
  push_cmd(Cmd{1, id, head(status, headers, -1, true, true)});
  return true;
}

bool HttpFrontend::send_chunk(uint64_t id, const std::string& data) {
  if (data.empty()) return true;
  {
    std::lock_guard<std::mutex> g(mu_);
    if (live_.find(id) == live_.end()) return false;
  }
  char hex[32];
  int n = snprintf(hex, sizeof(hex), "%zx\r\n", data.size());
  std::string s;
  s.reserve(data.size() + n + 2);
  s.append(hex, n).append(data).append("\r\n");
  push_cmd(Cmd{2, id, std::move(s)});
  return true;
}

bool HttpFrontend::end_stream(uint64_t id) {
  {
    std::lock_guard<std::mutex> g(mu_);
    if (live_.find(id) == live_.end()) return false;
  }
  push_cmd(Cmd{3, id, "0\r\n\r\n"});
  return true;
}


bool HttpFrontend::ws_accept(uint64_t id, const std::vector<std::pair<std::string, std::string>>& headers) {
  {
    std::lock_guard<std::mutex> g(mu_);
    if (live_.find(id) == live_.end()) return false;
  }
  std::string extra;
  for (auto& kv : headers) extra += kv.first + ": " + kv.second + "\r\n";
  push_cmd(Cmd{4, id, extra});          // the accept key is known to the I/O thread (it parsed the upgrade request)
  return true;
}

bool HttpFrontend::ws_send(uint64_t id, bool binary, const std::string& payload) {
  {
    std::lock_guard<std::mutex> g(mu_);
    if (live_.find(id) == live_.end()) return false;
  }
  push_cmd(Cmd{5, id, ws_frame(binary ? 2 : 1, payload)});
  return true;
}

bool HttpFrontend::ws_close(uint64_t id, int code) {
  {
    std::lock_guard<std::mutex> g(mu_);
    if (live_.find(id) == live_.end()) return false;
  }
  std::string p;
  p.push_back((char)((code >> 8) & 0xFF));
  p.push_back((char)(code & 0xFF));
  push_cmd(Cmd{6, id, ws_frame(8, p)});
  return true;
}

// ------------------------------------------------------------------------------------------------ I/O thread
void HttpFrontend::update_interest(Conn& c) {
  epoll_event ev{};
  ev.events = EPOLLIN | EPOLLRDHUP | (c.want_write ? EPOLLOUT : 0);
  ev.data.fd = c.fd;
  epoll_ctl(efd_, EPOLL_CTL_MOD, c.fd, &ev);
}

void HttpFrontend::close_conn(int fd, bool notify) {
  auto it = conns_.find(fd);
  if (it == conns_.end()) return;
  Conn& c = it->second;
  {
    std::lock_guard<std::mutex> g(mu_);
    if (c.cur) {
      live_.erase(c.cur);
      live_empty_ = live_.empty();
      if (notify) {
        HttpEvent e;
        e.type = 1;
        e.id = c.cur;
        post_event(std::move(e));
        stats_.disconnects++;
      }
    }
    stats_.open_conns--;
  }
  if (c.cur) {
    req_fd_.erase(c.cur);
    cv_.notify_one();
  }
  epoll_ctl(efd_, EPOLL_CTL_DEL, fd, nullptr);
  ::close(fd);
  conns_.erase(it);
}

void HttpFrontend::on_accept() {
  for (;;) {
    sockaddr_in a{};
    socklen_t len = sizeof(a);
    int fd = ::accept4(lfd_, (sockaddr*)&a, &len, SOCK_NONBLOCK | SOCK_CLOEXEC);
    if (fd < 0) return;
    if (!accepting_) {
      ::close(fd);
      continue;
    }
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    char ip[64] = {0};
    inet_ntop(AF_INET, &a.sin_addr, ip, sizeof(ip));
    Conn c;
    c.fd = fd;
    c.peer = std::string(ip) + ":" + std::to_string(ntohs(a.sin_port));
    conns_.emplace(fd, std::move(c));
    epoll_event ev{};
    ev.events = EPOLLIN | EPOLLRDHUP;
    ev.data.fd = fd;
    epoll_ctl(efd_, EPOLL_CTL_ADD, fd, &ev);
    std::lock_guard<std::mutex> g(mu_);
    stats_.accepted++;
    stats_.open_conns++;
  }
}

void HttpFrontend::simple_reply(Conn& c, int status, const std::string& text, bool close) {
  std::string body = "{\"error\": \"" + text + "\"}";
  c.out.push_back(head(status, {{"Content-Type", "application/json"}}, (int64_t)body.size(), false, !close) + body);
  c.close_after = close;
  flush(c);
}

void HttpFrontend::try_parse(Conn& c) {
  // one request at a time per connection: pipelined bytes wait in c.in until the current response has been written
  while (c.cur == 0 && !c.close_after) {
    size_t he = c.in.find("\r\n\r\n");
    if (he == std::string::npos) {
      if (c.in.size() > 65536) {
        { std::lock_guard<std::mutex> g(mu_); stats_.bad_requests++; }
        simple_reply(c, 431, "header block too large", true);
      }
      return;
    }
    // request line
    size_t le = c.in.find("\r\n");
    std::string line = c.in.substr(0, le);
    size_t s1 = line.find(' '), s2 = line.rfind(' ');
    if (s1 == std::string::npos || s2 == s1) {
      { std::lock_guard<std::mutex> g(mu_); stats_.bad_requests++; }
      simple_reply(c, 400, "malformed request line", true);
      return;
    }
    HttpEvent e;
    e.method = line.substr(0, s1);
    e.target = line.substr(s1 + 1, s2 - s1 - 1);
    e.version = line.substr(s2 + 1);
    int64_t clen = 0;
    bool chunked = false, expect = false;
    bool keep = e.version != "HTTP/1.0";
    size_t p = le + 2;
    while (p < he) {
      size_t q = c.in.find("\r\n", p);
      if (q == std::string::npos || q > he) q = he;
      size_t colon = c.in.find(':', p);
      if (colon != std::string::npos && colon < q) {
        std::string k = c.in.substr(p, colon - p);
        size_t vs = colon + 1;
        while (vs < q && (c.in[vs] == ' ' || c.in[vs] == '\t')) vs++;
        size_t ve = q;
        while (ve > vs && (c.in[ve - 1] == ' ' || c.in[ve - 1] == '\t')) ve--;
        std::string v = c.in.substr(vs, ve - vs);
        std::string kl = lower(k);
        if (kl == "content-length") clen = std::strtoll(v.c_str(), nullptr, 10);
        else if (kl == "transfer-encoding" && lower(v).find("chunked") != std::string::npos) chunked = true;
        else if (kl == "connection") {
          std::string vl = lower(v);
          if (vl.find("close") != std::string::npos) keep = false;
          else if (vl.find("keep-alive") != std::string::npos) keep = true;
        } else if (kl == "expect" && lower(v).find("100-continue") != std::string::npos) expect = true;
        else if (kl == "sec-websocket-key") c.ws_key = v;
        e.headers.emplace_back(std::move(k), std::move(v));
      }
      p = q + 2;
    }
    if (chunked) {
      { std::lock_guard<std::mutex> g(mu_); stats_.bad_requests++; }
      simple_reply(c, 411, "chunked request bodies are not accepted: send Content-Length", true);
      return;
    }
    if (clen < 0 || clen > max_body_) {
      { std::lock_guard<std::mutex> g(mu_); stats_.bad_requests++; }
      simple_reply(c, 413, "request body too large", true);
      return;
    }
    size_t total = he + 4 + (size_t)clen;
    if (c.in.size() < total) {
      if (expect && !c.sent_continue) {
        c.sent_continue = true;
        c.out.push_back("HTTP/1.1 100 Continue\r\n\r\n");
        flush(c);
      }
      return;
    }
    e.body = c.in.substr(he + 4, (size_t)clen);
    c.in.erase(0, total);
    c.sent_continue = false;
    c.keep_alive = keep && accepting_;
    // answered here?
    std::string path = e.target.substr(0, e.target.find('?'));
    bool served = false;
    if (e.method == "GET" || e.method == "HEAD") {
      std::unique_lock<std::mutex> g(mu_);
      auto it = statics_.find(path);
      if (it != statics_.end()) {
        std::string h = head(it->second.status, {{"Content-Type", it->second.ctype}}, (int64_t)it->second.body.size(), false, c.keep_alive);
        std::string b = e.method == "HEAD" ? std::string() : it->second.body;
        stats_.static_hits++;
        stats_.requests++;
        g.unlock();
        c.out.push_back(h + b);
        c.close_after = !c.keep_alive;
        served = true;
      }
    }
    if (served) {
      flush(c);
      if (conns_.find(c.fd) == conns_.end()) return;      // flush may have closed it
      continue;
    }
    e.id = next_id_++;
    e.peer = c.peer;
    c.cur = e.id;
    c.streaming = false;
    req_fd_[e.id] = c.fd;
    {
      std::lock_guard<std::mutex> g(mu_);
      live_.insert(e.id);
      live_empty_ = false;
      stats_.requests++;
      post_event(std::move(e));
    }
    cv_.notify_one();
  }
}

void HttpFrontend::on_readable(Conn& c) {
  char buf[65536];
  for (;;) {
    ssize_t n = ::recv(c.fd, buf, sizeof(buf), 0);
    if (n > 0) {
      c.in.append(buf, (size_t)n);
      { std::lock_guard<std::mutex> g(mu_); stats_.bytes_in += (uint64_t)n; }
      if ((size_t)n < sizeof(buf)) break;
      continue;
    }
    if (n == 0) {          // peer closed: an unanswered request becomes a disconnect event (Python aborts the generation)
      close_conn(c.fd, true);
      return;
    }
    if (errno == EAGAIN || errno == EWOULDBLOCK) break;
    if (errno == EINTR) continue;
    close_conn(c.fd, true);
    return;
  }
  if (c.ws) parse_ws(c);
  else try_parse(c);
}

void HttpFrontend::parse_ws(Conn& c) {
  for (;;) {
    if (c.in.size() < 2) return;
    const unsigned char* b = (const unsigned char*)c.in.data();
    bool fin = b[0] & 0x80, masked = b[1] & 0x80;
    int op = b[0] & 0x0F;
    uint64_t n = b[1] & 0x7F;
    size_t hdr = 2;
    if (n == 126) {
      if (c.in.size() < 4) return;
      n = ((uint64_t)b[2] << 8) | b[3];
      hdr = 4;
    } else if (n == 127) {
      if (c.in.size() < 10) return;
      n = 0;
      for (int i = 0; i < 8; ++i) n = (n << 8) | b[2 + i];
      hdr = 10;
    }
    if ((int64_t)n > max_body_ || !masked) {          // clients MUST mask; oversized messages are refused
      close_conn(c.fd, true);
      return;
    }
    if (c.in.size() < hdr + 4 + n) return;
    const unsigned char* mask = b + hdr;
    std::string payload((size_t)n, '\0');
    for (uint64_t i = 0; i < n; ++i) payload[i] = (char)(b[hdr + 4 + i] ^ mask[i & 3]);
    c.in.erase(0, hdr + 4 + (size_t)n);
    if (op == 9) {          // ping -> pong, answered here
      c.out.push_back(ws_frame(10, payload));
      flush(c);
      if (conns_.find(c.fd) == conns_.end()) return;
      continue;
    }
    if (op == 10) continue;
    if (op == 8) {          // close: echo it, tell Python, drop the connection once written
      c.out.push_back(ws_frame(8, payload.substr(0, std::min<size_t>(2, payload.size()))));
      uint64_t id = c.cur;
      {
        std::lock_guard<std::mutex> g(mu_);
        HttpEvent e;
        e.type = 3;
        e.id = id;
        e.body = payload;
        post_event(std::move(e));
        live_.erase(id);
        live_empty_ = live_.empty();
      }
      cv_.notify_one();
      req_fd_.erase(id);
      c.cur = 0;
      c.close_after = true;
      flush(c);
      return;
    }
    if (op == 1 || op == 2) {
      c.ws_opcode = op;
      c.ws_msg = payload;
    } else if (op == 0) {
      c.ws_msg += payload;
    }
    if (fin) {
      std::lock_guard<std::mutex> g(mu_);
      HttpEvent e;
      e.type = 2;
      e.id = c.cur;
      e.method = c.ws_opcode == 2 ? "binary" : "text";
      e.body = std::move(c.ws_msg);
      c.ws_msg.clear();
      post_event(std::move(e));
      cv_.notify_one();
    }
  }
}

void HttpFrontend::finish_response(Conn& c) {
  {
    std::lock_guard<std::mutex> g(mu_);
    live_.erase(c.cur);
    live_empty_ = live_.empty();
  }
  req_fd_.erase(c.cur);
  c.cur = 0;
  c.streaming = false;
  if (!c.keep_alive) c.close_after = true;
}

void HttpFrontend::flush(Conn& c) {
  int fd = c.fd;
  while (!c.out.empty()) {
    std::string& s = c.out.front();
    ssize_t n = ::send(fd, s.data() + c.out_off, s.size() - c.out_off, MSG_NOSIGNAL);
    if (n > 0) {
      c.out_off += (size_t)n;
      { std::lock_guard<std::mutex> g(mu_); stats_.bytes_out += (uint64_t)n; }
      if (c.out_off == s.size()) {
        c.out.pop_front();
        c.out_off = 0;
      }
      continue;
    }
    if (n < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) {
      if (!c.want_write) {
        c.want_write = true;
        update_interest(c);
      }
      return;
    }
    if (n < 0 && errno == EINTR) continue;
    close_conn(fd, true);
    return;
  }
  if (c.want_write) {
    c.want_write = false;
    update_interest(c);
  }
  if (c.close_after && c.cur == 0) {
    close_conn(fd, false);
    return;
  }
  if (c.cur == 0 && !c.in.empty()) try_parse(c);          // a pipelined request was waiting
}

void HttpFrontend::apply(Cmd& cmd) {
  auto it = req_fd_.find(cmd.id);
  if (it == req_fd_.end()) return;
  auto ci = conns_.find(it->second);
  if (ci == conns_.end()) return;
  Conn& c = ci->second;
  if (cmd.op == 0 || cmd.op == 1) {
    if (!c.keep_alive) {          // the head was formatted with keep-alive: rewrite the framing line for a closing connection
      size_t p = cmd.data.find("Connection: keep-alive\r\n");
      if (p != std::string::npos) cmd.data.replace(p, 24, "Connection: close\r\n");
    }
  }
  if (cmd.op == 4) {          // 101 Switching Protocols
    unsigned char dg[20];
    sha1(c.ws_key + "258EAFA5-E914-47DA-95CA-C5AB0DC85B11", dg);
    c.out.push_back("HTTP/1.1 101 Switching Protocols\r\nUpgrade: websocket\r\nConnection: Upgrade\r\nSec-WebSocket-Accept: " + b64(dg, 20) + "\r\n" +
                    cmd.data + "\r\n");
    c.ws = true;
    flush(c);
    if (conns_.find(it->second) != conns_.end() && !c.in.empty()) parse_ws(c);
    return;
  }
  if (cmd.op == 6) {          // our close: send it, forget the request, drop the connection once written
    c.out.push_back(std::move(cmd.data));
    finish_response(c);
    c.close_after = true;
    flush(c);
    return;
  }
  c.out.push_back(std::move(cmd.data));
  if (cmd.op == 1) c.streaming = true;
  if (cmd.op == 0 || cmd.op == 3) finish_response(c);
  flush(c);
}

void HttpFrontend::loop() {
  epoll_event evs[256];
  while (running_) {
    int n = epoll_wait(efd_, evs, 256, 200);
    for (int i = 0; i < n; ++i) {
      int fd = evs[i].data.fd;
      if (fd == lfd_) {
        on_accept();
        continue;
      }
      if (fd == wfd_) {
        uint64_t v;
        (void)!::read(wfd_, &v, 8);
        continue;      // commands are drained below
      }
      auto it = conns_.find(fd);
      if (it == conns_.end()) continue;
      if (evs[i].events & (EPOLLERR | EPOLLHUP)) {
        close_conn(fd, true);
        continue;
      }
      if (evs[i].events & EPOLLIN) {
        on_readable(it->second);
        it = conns_.find(fd);
        if (it == conns_.end()) continue;
      }
      if (evs[i].events & EPOLLRDHUP) {      // half-close: nothing more will arrive; an unanswered request is abandoned
        close_conn(fd, true);
        continue;
      }
      if (evs[i].events & EPOLLOUT) flush(it->second);
    }
    std::deque<Cmd> todo;
    {
      std::lock_guard<std::mutex> g(mu_);
      todo.swap(cmds_);
    }
    for (auto& c : todo) apply(c);
    if (!accepting_) {          // draining: no new connections, idle keep-alive connections are closed as they fall idle
      if (lfd_ >= 0) {
        epoll_ctl(efd_, EPOLL_CTL_DEL, lfd_, nullptr);
        ::close(lfd_);
        lfd_ = -1;
      }
      std::vector<int> idle;
      for (auto& kv : conns_)
        if (kv.second.cur == 0 && kv.second.out.empty()) idle.push_back(kv.first);
      for (int fd : idle) close_conn(fd, false);
    }
  }
  std::vector<int> all;
  for (auto& kv : conns_) all.push_back(kv.first);
  for (int fd : all) close_conn(fd, false);
  if (lfd_ >= 0) {
    ::close(lfd_);
    lfd_ = -1;
  }
}

}  // namespace sglb_rt
