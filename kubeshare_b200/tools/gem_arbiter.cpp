// gem-arbiter -- native replacement of gem-schd + gem-pmgr for hooks that still speak TCP.
//
// It owns (creates) the per-GPU shared credit pool, keeps it in step with the quota file that
// kubeshare-config writes (inotify, like reference scheduler.cpp:219-265), optionally mirrors that file onto
// the hostPath pods mount, and serves the unchanged 80/40-byte token protocol (reference comm.h:28-31) to
// legacy clients -- the reference libgemhook.so.1 or the reference gem-pmgr -- by running their requests through
// the SAME pool that native hooks (GEMHOOK_POOL) arbitrate in.  Old and new hooks on one GPU therefore share one
// ledger and one token.
//
//   gem-arbiter --pool FILE -p DIR -f QUOTAFILE [-P PORT]... [--port-file FILE] [-q BASE] [-m MIN] [-w WINDOW]
//               [--mirror FILE] [--columns limit_request] [-v]
// --port-file is KubeShare's /kubeshare/scheduler/podmanagerport/<GPU-UUID> ("N" then "name port" rows,
// pkg/config/query.go:57, 86-98): a listener is opened for every port listed and closed when the row disappears --
// what launcher.py does by spawning / killing one gem-pmgr per row (reference launcher.py:34-67), without Python.
// Flags -P -q -m -w -f -p mean what they mean to gem-schd (scheduler.cpp:555-604).  Several -P give several
// listeners (KubeShare hands every pod its own POD_MANAGER_PORT, pkg/scheduler/node.go:14).
//
// Answers: REQ_QUOTA -> blocks in gemhook_pool_acquire (pod-level rule + scheduler policy) and replies the quota;
// REQ_MEM_LIMIT -> (used, limit) of the pod (what gem-pmgr answers, pod-manager.cpp:501-504);
// REQ_MEM_UPDATE -> verdict of the atomic reserve / release; bytes are tracked per connection and given back
// when the connection closes (pod-manager.cpp:533-545).
#include <arpa/inet.h>
#include <errno.h>
#include <limits.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <pthread.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/inotify.h>
#include <sys/socket.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "../../include/gemhook.h"

static gemhook_pool* g_pool;
// legacy connections per pool slot: processes of one pod share the pod's token (pod-manager.cpp:316-473), so the
// token is handed back only when the LAST connection of that pod closes
static int g_conns[GEMHOOK_MAX_SLOTS];
static pthread_mutex_t g_conns_mu = PTHREAD_MUTEX_INITIALIZER;
static std::string g_dir = ".", g_file = "resource-config.txt", g_mirror;
static int g_swap = 0, g_verbose = 0;

static std::string quota_path() { return g_dir + (g_dir.empty() || g_dir.back() == '/' ? "" : "/") + g_file; }

static bool read_file(const std::string& path, std::string* out) {
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return false;
  char buf[4096];
  size_t n;
  out->clear();
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out->append(buf, n);
  fclose(f);
  return true;
}

static void load_quota_file(void) {
  std::string text;
  if (!read_file(quota_path(), &text)) {
    fprintf(stderr, "[gem-arbiter] cannot read %s: %s\n", quota_path().c_str(), strerror(errno));
    return;
  }
  int n = gemhook_pool_load_config(g_pool, text.c_str(), g_swap);
  fprintf(stderr, "[gem-arbiter] quota file: %d clients\n", n);
  if (!g_mirror.empty()) {  // pods mount only /kubeshare/library: give native hooks a copy they can read
    std::string tmp = g_mirror + ".tmp";
    FILE* f = fopen(tmp.c_str(), "w");
    if (f) {
      fwrite(text.data(), 1, text.size(), f);
      fclose(f);
      rename(tmp.c_str(), g_mirror.c_str());
    }
  }
}

static void* watch_main(void*) {
  int fd = inotify_init();
  if (fd < 0 || inotify_add_watch(fd, g_dir.c_str(), IN_CLOSE_WRITE | IN_MOVED_TO) < 0) {
    fprintf(stderr, "[gem-arbiter] inotify on %s failed: %s\n", g_dir.c_str(), strerror(errno));
    return nullptr;
  }
  char buf[4096] __attribute__((aligned(8)));
  for (;;) {
    ssize_t len = read(fd, buf, sizeof(buf));
    if (len <= 0) continue;
    for (char* p = buf; p < buf + len;) {
      struct inotify_event* ev = (struct inotify_event*)p;
      if (ev->len && !strcmp(ev->name, g_file.c_str())) load_quota_file();
      p += sizeof(struct inotify_event) + ev->len;
    }
  }
  return nullptr;
}

static int recv_all(int fd, uint8_t* b, size_t n) {
  while (n) {
    ssize_t k = recv(fd, b, n, 0);
    if (k < 0 && errno == EINTR) continue;
    if (k <= 0) return -1;
    b += k;
    n -= (size_t)k;
  }
  return 0;
}
static int send_all(int fd, const uint8_t* b, size_t n) {
  while (n) {
    ssize_t k = send(fd, b, n, MSG_NOSIGNAL);
    if (k < 0 && errno == EINTR) continue;
    if (k <= 0) return -1;
    b += k;
    n -= (size_t)k;
  }
  return 0;
}

// one legacy client connection
static void* conn_main(void* arg) {
  int fd = (int)(intptr_t)arg;
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
  uint8_t in[GEMHOOK_REQ_LEN], out[GEMHOOK_RSP_LEN];
  uint64_t held = 0;  // bytes this connection reserved (allocation_map[sockfd], pod-manager.cpp:92)
  int held_slot = -1, token_slot = -1, counted_slot = -1;
  while (recv_all(fd, in, sizeof(in)) == 0) {
    gemhook_request req;
    if (gemhook_wire_unpack_request(in, &req) < 0) break;
    int slot = gemhook_pool_find(g_pool, req.name);
    if (slot < 0) {  // gem-schd: "Unknown client ... Ignore this request" (scheduler.cpp:411-414) -> no reply
      fprintf(stderr, "[gem-arbiter] unknown client \"%s\": request ignored\n", req.name);
      continue;
    }
    if (counted_slot < 0 && slot < GEMHOOK_MAX_SLOTS) {
      pthread_mutex_lock(&g_conns_mu);
      g_conns[slot]++;
      pthread_mutex_unlock(&g_conns_mu);
      counted_slot = slot;
    }
    gemhook_response rsp;
    memset(&rsp, 0, sizeof(rsp));
    rsp.req_id = req.req_id;
    if (req.type == GEMHOOK_REQ_QUOTA) {
      rsp.quota_ms = gemhook_pool_acquire(g_pool, slot, req.overuse_ms, req.burst_ms);
      token_slot = slot;
      if (g_verbose) fprintf(stderr, "[gem-arbiter] %s overuse %.3f burst %.3f -> quota %.3f\n", req.name, req.overuse_ms, req.burst_ms, rsp.quota_ms);
    } else if (req.type == GEMHOOK_REQ_MEM_LIMIT) {
      gemhook_pool_mem_info(g_pool, slot, &rsp.mem_used, &rsp.mem_total);
    } else if (req.type == GEMHOOK_REQ_MEM_UPDATE) {
      if (req.is_alloc) {
        rsp.verdict = gemhook_pool_mem_reserve(g_pool, slot, req.bytes);
        if (rsp.verdict) {
          held += req.bytes;
          held_slot = slot;
        }
      } else {
        gemhook_pool_mem_release(g_pool, slot, req.bytes);
        held -= req.bytes;
        rsp.verdict = 1;
      }
    } else {
      continue;  // unknown request type: ignored like scheduler.cpp:456-458
    }
    gemhook_wire_pack_response(req.type, &rsp, out);
    if (send_all(fd, out, sizeof(out)) != 0) break;
  }
  if (held && held_slot >= 0) gemhook_pool_mem_release(g_pool, held_slot, held);  // process gone: reclaim
  bool last = true;
  if (counted_slot >= 0) {
    pthread_mutex_lock(&g_conns_mu);
    last = --g_conns[counted_slot] == 0;
    pthread_mutex_unlock(&g_conns_mu);
  }
  // a pod that disappears while holding the token must not stall the others until its deadline -- but only when its
  // last process is gone (a sibling may be running kernels under that token)
  if (token_slot >= 0 && last) gemhook_pool_release(g_pool, token_slot);
  close(fd);
  return nullptr;
}

#include <map>
static std::map<int, int> g_listeners;  // port -> listening fd (dynamic ones from --port-file)
static pthread_mutex_t g_listeners_mu = PTHREAD_MUTEX_INITIALIZER;
static std::string g_port_file;
static std::vector<int> g_static_ports;  // from -P: never closed by the port file

static void* listen_main(void* arg) {
  int port = (int)(intptr_t)arg;
  int ls = socket(AF_INET, SOCK_STREAM, 0);
  int one = 1;
  setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  struct sockaddr_in sa;
  memset(&sa, 0, sizeof(sa));
  sa.sin_family = AF_INET;
  sa.sin_addr.s_addr = INADDR_ANY;
  sa.sin_port = htons((uint16_t)port);
  if (bind(ls, (struct sockaddr*)&sa, sizeof(sa)) != 0 || listen(ls, SOMAXCONN) != 0) {
    fprintf(stderr, "[gem-arbiter] cannot listen on port %d: %s\n", port, strerror(errno));
    exit(1);
  }
  fprintf(stderr, "[gem-arbiter] listening on port %d\n", port);
  pthread_mutex_lock(&g_listeners_mu);
  g_listeners[port] = ls;
  pthread_mutex_unlock(&g_listeners_mu);
  for (;;) {
    int fd = accept(ls, nullptr, nullptr);
    if (fd < 0) {
      if (errno == EBADF || errno == EINVAL) break;  // listener closed: the pod left the port file
      continue;
    }
    pthread_t t;
    pthread_create(&t, nullptr, conn_main, (void*)(intptr_t)fd);
    pthread_detach(t);
  }
  fprintf(stderr, "[gem-arbiter] port %d closed\n", port);
  return nullptr;
}

// podmanagerport/<UUID>: open listeners for new rows, close the ones that disappeared
static void sync_port_file(void) {
  std::string text;
  if (g_port_file.empty() || !read_file(g_port_file, &text)) return;
  std::vector<int> want;
  const char* c = text.c_str();
  char* end = nullptr;
  long n = strtol(c, &end, 10);
  c = end;
  for (long i = 0; i < n; i++) {
    char name[256];
    int port = 0, used = 0;
    if (sscanf(c, " %255s %d%n", name, &port, &used) != 2) break;
    c += used;
    if (port > 0) want.push_back(port);
  }
  std::vector<int> to_open, to_close;
  pthread_mutex_lock(&g_listeners_mu);
  for (int p : want)
    if (!g_listeners.count(p)) to_open.push_back(p);
  for (auto& kv : g_listeners) {
    bool keep = false;
    for (int p : g_static_ports) keep = keep || p == kv.first;
    for (int p : want) keep = keep || p == kv.first;
    if (!keep) to_close.push_back(kv.first);
  }
  for (int p : to_close) {
    shutdown(g_listeners[p], SHUT_RDWR);
    close(g_listeners[p]);
    g_listeners.erase(p);
  }
  pthread_mutex_unlock(&g_listeners_mu);
  for (int p : to_open) {
    pthread_t t;
    pthread_create(&t, nullptr, listen_main, (void*)(intptr_t)p);
    pthread_detach(t);
  }
}

static void* port_watch_main(void*) {
  std::string dir = g_port_file, base = g_port_file;
  size_t slash = g_port_file.find_last_of('/');
  if (slash == std::string::npos) dir = ".";
  else {
    dir = g_port_file.substr(0, slash);
    base = g_port_file.substr(slash + 1);
  }
  int fd = inotify_init();
  if (fd < 0 || inotify_add_watch(fd, dir.c_str(), IN_CLOSE_WRITE | IN_MOVED_TO) < 0) return nullptr;
  char buf[4096] __attribute__((aligned(8)));
  for (;;) {
    ssize_t len = read(fd, buf, sizeof(buf));
    if (len <= 0) continue;
    for (char* p = buf; p < buf + len;) {
      struct inotify_event* ev = (struct inotify_event*)p;
      if (ev->len && base == ev->name) sync_port_file();
      p += sizeof(struct inotify_event) + ev->len;
    }
  }
  return nullptr;
}

int main(int argc, char** argv) {
  std::string pool_path;
  std::vector<int> ports;
  double base_q = 250.0, min_q = 100.0, window = 10000.0;  // gem-schd's binary defaults (scheduler.cpp:90-92)
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
    if (a == "--pool") pool_path = next();
    else if (a == "-p" || a == "--limit_file_dir") g_dir = next();
    else if (a == "-f" || a == "--limit_file") g_file = next();
    else if (a == "-P" || a == "--port") ports.push_back(atoi(next()));
    else if (a == "-q" || a == "--quota") base_q = atof(next());
    else if (a == "-m" || a == "--min_quota") min_q = atof(next());
    else if (a == "-w" || a == "--window") window = atof(next());
    else if (a == "--mirror") g_mirror = next();
    else if (a == "--port-file") g_port_file = next();
    else if (a == "--columns") g_swap = !strcmp(next(), "limit_request");
    else if (a == "-v" || a == "--verbose") { g_verbose = 1; if (i + 1 < argc && argv[i + 1][0] != '-') i++; }
    else if (a == "-h" || a == "--help") {
      puts("usage: gem-arbiter --pool FILE -p DIR -f QUOTAFILE [-P PORT]... [-q BASE] [-m MIN] [-w WINDOW] [--mirror FILE] [--columns limit_request] [-v]");
      return 0;
    }
  }
  if (pool_path.empty()) {
    fprintf(stderr, "gem-arbiter: --pool FILE is required\n");
    return 2;
  }
  if (ports.empty() && g_port_file.empty()) {
    const char* e = getenv("POD_MANAGER_PORT");  // started in gem-pmgr's place: same env (pod-manager.cpp:180-184)
    ports.push_back(e ? atoi(e) : 50051);
  }
  signal(SIGPIPE, SIG_IGN);
  g_pool = gemhook_pool_open(pool_path.c_str(), 1, base_q, min_q, window, 0);
  if (!g_pool) {
    fprintf(stderr, "gem-arbiter: %s\n", gemhook_last_error());
    return 1;
  }
  load_quota_file();
  pthread_t t;
  pthread_create(&t, nullptr, watch_main, nullptr);
  pthread_detach(t);
  g_static_ports = ports;
  for (size_t i = 0; i < ports.size(); i++) {
    pthread_create(&t, nullptr, listen_main, (void*)(intptr_t)ports[i]);
    pthread_detach(t);
  }
  if (!g_port_file.empty()) {
    sync_port_file();
    pthread_create(&t, nullptr, port_watch_main, nullptr);
    pthread_detach(t);
  }
  for (;;) pause();
  return 0;
}
