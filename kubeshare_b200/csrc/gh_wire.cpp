// gh_wire.cpp -- the Gemini token protocol (fixed 80-byte requests, 40-byte responses) and the TCP
// transport to the unmodified gem-pmgr / gem-schd.
//
// Layout (reference comm.h:28-31, comm.cpp:26-120; verified against the reference's object code in
// tests/golden/ref_golden.json): native endian, packed, unaligned.
//   request : [u64 name_len][name][0x00][i32 req_id][i32 type] + payload, zero padded to 80
//             REQ_QUOTA: [f64 overuse_ms][f64 burst_ms]   REQ_MEM_UPDATE: [u64 bytes][i32 is_alloc]
//   response: [i32 req_id] + payload, zero padded to 40
//             REQ_QUOTA: [f64 quota_ms]  REQ_MEM_LIMIT: [u64 used][u64 total]  REQ_MEM_UPDATE: [i32 verdict]
// Unlike the reference (comm.cpp:42-60 writes past the 80-byte buffer for long pod names) an
// over-long name is rejected.
#include <arpa/inet.h>
#include <errno.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>

#include "gh_internal.h"

namespace {
struct Cursor {
  uint8_t* p;
  size_t off;
  template <typename T>
  void put(T v) {
    memcpy(p + off, &v, sizeof(T));
    off += sizeof(T);
  }
};
struct Reader {
  const uint8_t* p;
  size_t off;
  template <typename T>
  T get() {
    T v;
    memcpy(&v, p + off, sizeof(T));
    off += sizeof(T);
    return v;
  }
};
}  // namespace

GH_EXPORT int gemhook_wire_pack_request(const gemhook_request* req, uint8_t* out) {
  size_t nlen = strnlen(req->name, sizeof(req->name));
  size_t payload = req->type == GEMHOOK_REQ_QUOTA ? 16 : (req->type == GEMHOOK_REQ_MEM_UPDATE ? 12 : 0);
  if (8 + nlen + 1 + 8 + payload > GEMHOOK_REQ_LEN) {
    gh_set_error("pod name of %zu bytes does not fit the 80-byte request", nlen);
    return -1;
  }
  memset(out, 0, GEMHOOK_REQ_LEN);
  Cursor c{out, 0};
  c.put<uint64_t>(nlen);
  memcpy(out + c.off, req->name, nlen);
  c.off += nlen + 1;  // terminator already zero
  c.put<int32_t>(req->req_id);
  c.put<int32_t>(req->type);
  if (req->type == GEMHOOK_REQ_QUOTA) {
    c.put<double>(req->overuse_ms);
    c.put<double>(req->burst_ms);
  } else if (req->type == GEMHOOK_REQ_MEM_UPDATE) {
    c.put<uint64_t>(req->bytes);
    c.put<int32_t>(req->is_alloc);
  }
  return (int)c.off;
}

GH_EXPORT int gemhook_wire_unpack_request(const uint8_t* in, gemhook_request* req) {
  memset(req, 0, sizeof(*req));
  Reader r{in, 0};
  uint64_t nlen = r.get<uint64_t>();
  if (nlen > GEMHOOK_REQ_LEN - 17 || nlen >= sizeof(req->name)) {
    gh_set_error("malformed request: name_len %llu", (unsigned long long)nlen);
    return -1;
  }
  memcpy(req->name, in + 8, nlen);
  r.off += nlen + 1;
  req->req_id = r.get<int32_t>();
  req->type = r.get<int32_t>();
  if (req->type == GEMHOOK_REQ_QUOTA) {
    if (r.off + 16 > GEMHOOK_REQ_LEN) return -1;
    req->overuse_ms = r.get<double>();
    req->burst_ms = r.get<double>();
  } else if (req->type == GEMHOOK_REQ_MEM_UPDATE) {
    if (r.off + 12 > GEMHOOK_REQ_LEN) return -1;
    req->bytes = r.get<uint64_t>();
    req->is_alloc = r.get<int32_t>();
  }
  return (int)r.off;
}

GH_EXPORT int gemhook_wire_pack_response(int32_t type, const gemhook_response* rsp, uint8_t* out) {
  memset(out, 0, GEMHOOK_RSP_LEN);
  Cursor c{out, 0};
  c.put<int32_t>(rsp->req_id);
  if (type == GEMHOOK_REQ_QUOTA) {
    c.put<double>(rsp->quota_ms);
  } else if (type == GEMHOOK_REQ_MEM_LIMIT) {
    c.put<uint64_t>(rsp->mem_used);
    c.put<uint64_t>(rsp->mem_total);
  } else if (type == GEMHOOK_REQ_MEM_UPDATE) {
    c.put<int32_t>(rsp->verdict);
  }
  return (int)c.off;
}

GH_EXPORT int gemhook_wire_unpack_response(int32_t type, const uint8_t* in, gemhook_response* rsp) {
  memset(rsp, 0, sizeof(*rsp));
  Reader r{in, 0};
  rsp->req_id = r.get<int32_t>();
  if (type == GEMHOOK_REQ_QUOTA) {
    rsp->quota_ms = r.get<double>();
  } else if (type == GEMHOOK_REQ_MEM_LIMIT) {
    rsp->mem_used = r.get<uint64_t>();
    rsp->mem_total = r.get<uint64_t>();
  } else if (type == GEMHOOK_REQ_MEM_UPDATE) {
    rsp->verdict = r.get<int32_t>();
  }
  return (int)r.off;
}

// ---- TCP transport (one blocking RPC at a time, like reference hook.cpp:300-328) -----------------------
static pthread_mutex_t rpc_mu = PTHREAD_MUTEX_INITIALIZER;
static int rpc_fd = -1;
static int32_t rpc_next_id = 0;  // per-process counter from 0 (reference comm.cpp:29, 62)

static int full_send(int fd, const uint8_t* b, size_t n) {
  while (n) {
    ssize_t k = send(fd, b, n, MSG_NOSIGNAL);
    if (k < 0) {
      if (errno == EINTR) continue;
      return -1;
    }
    b += k;
    n -= (size_t)k;
  }
  return 0;
}
static int full_recv(int fd, uint8_t* b, size_t n) {
  while (n) {
    ssize_t k = recv(fd, b, n, 0);
    if (k < 0 && errno == EINTR) continue;
    if (k <= 0) return -1;
    b += k;
    n -= (size_t)k;
  }
  return 0;
}

// connect with the reference's retry policy: 5 attempts, 10 s apart (hook.cpp:167-168, 280-286);
// GEMHOOK_CONNECT_RETRY_S shortens the interval for tests.
static int rpc_connect(void) {
  const char* ip = gh_cfg.scheduler_ip[0] ? gh_cfg.scheduler_ip : "127.0.0.1";
  int retry_s = getenv("GEMHOOK_CONNECT_RETRY_S") ? atoi(getenv("GEMHOOK_CONNECT_RETRY_S")) : 10;
  for (int attempt = 1; attempt <= 5; attempt++) {
    int fd = socket(AF_INET, SOCK_STREAM, 0);
    if (fd < 0) return -1;
    struct sockaddr_in sa;
    memset(&sa, 0, sizeof(sa));
    sa.sin_family = AF_INET;
    sa.sin_addr.s_addr = inet_addr(ip);
    sa.sin_port = htons((uint16_t)gh_cfg.pod_manager_port);
    if (connect(fd, (struct sockaddr*)&sa, sizeof(sa)) == 0) {
      int one = 1;
      setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));  // 80/40-byte ping-pong: no Nagle
      return fd;
    }
    int e = errno;
    close(fd);
    GH_INFO("connect to pod manager %s:%d failed (attempt %d): %s", ip, gh_cfg.pod_manager_port, attempt, strerror(e));
    if (attempt < 5 && retry_s > 0) sleep((unsigned)retry_s);
  }
  return -1;
}

// One request/response exchange. Returns 0 and fills rsp, or -1 (caller decides whether to exit()).
// Retry policy of the reference's communicate() (hook.cpp:300-328 with comm.cpp:124-134): the receive timeout is set
// per call (10 s for the memory requests, none for a token request, hook.cpp:357, 385, 437), and a failed send or
// receive is retried -- send AND receive again, on the same socket -- up to NET_OP_MAX_ATTEMPT = 5 times before the
// exchange is declared failed.
int gh_rpc(gemhook_request* req, gemhook_response* rsp) {
  uint8_t sbuf[GEMHOOK_REQ_LEN], rbuf[GEMHOOK_RSP_LEN];
  pthread_mutex_lock(&rpc_mu);
  int rc = -1;
  if (rpc_fd < 0) rpc_fd = rpc_connect();
  if (rpc_fd >= 0) {
    snprintf(req->name, sizeof(req->name), "%s", gh_cfg.pod_name);
    req->req_id = rpc_next_id++;
    if (gemhook_wire_pack_request(req, sbuf) > 0) {
      struct timeval tv;
      static int mem_timeout_s = getenv("GEMHOOK_RPC_TIMEOUT_S") ? atoi(getenv("GEMHOOK_RPC_TIMEOUT_S")) : 10;  // NET_OP_RETRY_INTV
      tv.tv_sec = req->type == GEMHOOK_REQ_QUOTA ? 0 : mem_timeout_s;
      tv.tv_usec = 0;
      setsockopt(rpc_fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
      for (int attempt = 1; attempt <= 5 && rc != 0; attempt++) {
        if (full_send(rpc_fd, sbuf, sizeof(sbuf)) == 0 && full_recv(rpc_fd, rbuf, sizeof(rbuf)) == 0) {
          gemhook_wire_unpack_response(req->type, rbuf, rsp);
          rc = 0;
        } else {
          GH_INFO("token protocol exchange failed (attempt %d): %s", attempt, strerror(errno));
        }
      }
    }
    if (rc != 0) {
      gh_set_error("token protocol exchange failed: %s", strerror(errno));
      close(rpc_fd);
      rpc_fd = -1;
    }
  } else {
    gh_set_error("cannot reach the pod manager at %s:%d", gh_cfg.scheduler_ip, gh_cfg.pod_manager_port);
  }
  pthread_mutex_unlock(&rpc_mu);
  return rc;
}
