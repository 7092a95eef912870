// gh_mem.cpp -- gpu_mem cap: integer-exact accounting of REQUESTED bytes.
//
// Reference rule (hook.cpp:590-617, pod-manager.cpp:295-313): an allocation of `bytes` is allowed iff
// bytes <= limit - used (size_t arithmetic) before the driver call; used grows by the requested size
// (no driver rounding); cuMemFree of a known pointer returns its bytes, an unknown pointer is a no-op
// (hook.cpp:570-581); cuMemGetInfo / cuDeviceTotalMem report (limit - used, limit) and limit
// (hook.cpp:857-872).
//
// Differences, none visible in the byte arithmetic: the reservation is ONE atomic step taken BEFORE the
// driver call (CAS on the shared pool counter, or one REQ_MEM_UPDATE to gem-pmgr) and rolled back if the
// driver fails -- the reference asks twice (REQ_MEM_LIMIT then REQ_MEM_UPDATE) and leaks the real
// allocation when the second answer is negative (hook.cpp:606-609).
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "gh_internal.h"

int gh_rpc(gemhook_request* req, gemhook_response* rsp);

struct gh_live;
// accessors implemented at the bottom of gh_hook.cpp would create a cycle; the two fields we need are
// exported through these tiny helpers instead
extern "C" gemhook_pool* gh_live_pool(void);
extern "C" int gh_live_slot(void);
extern "C" int gh_live_enabled(void);

namespace {

// open-addressing pointer -> bytes table (the reference uses std::map under a mutex, hook.cpp:193-195)
struct Table {
  struct Ent {
    uint64_t key;
    uint64_t bytes;
  };
  Ent* e = nullptr;
  size_t cap = 0, used = 0, tomb = 0;
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  static const uint64_t EMPTY = 0, TOMB = ~0ull;

  static size_t hash(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    return (size_t)k;
  }
  void rehash(size_t ncap) {
    Ent* old = e;
    size_t ocap = cap;
    e = (Ent*)calloc(ncap, sizeof(Ent));
    cap = ncap;
    used = tomb = 0;
    for (size_t i = 0; i < ocap; i++)
      if (old[i].key != EMPTY && old[i].key != TOMB) put_nolock(old[i].key, old[i].bytes);
    free(old);
  }
  void put_nolock(uint64_t k, uint64_t b) {
    if ((used + tomb + 1) * 4 > cap * 3) rehash(cap ? cap * 2 : 1024);
    size_t i = hash(k) & (cap - 1);
    while (e[i].key != EMPTY && e[i].key != TOMB && e[i].key != k) i = (i + 1) & (cap - 1);
    if (e[i].key == TOMB) tomb--;
    if (e[i].key != k) used++;
    e[i].key = k;
    e[i].bytes = b;
  }
  bool take_nolock(uint64_t k, uint64_t* b) {
    if (!cap) return false;
    size_t i = hash(k) & (cap - 1);
    while (e[i].key != EMPTY) {
      if (e[i].key == k) {
        *b = e[i].bytes;
        e[i].key = TOMB;
        used--;
        tomb++;
        return true;
      }
      i = (i + 1) & (cap - 1);
    }
    return false;
  }
};

Table g_table;
std::atomic<uint64_t> g_denied{0};
std::atomic<uint64_t> g_local_used{0};  // this process's share (the reference's gpu_mem_used, hook.cpp:195)
uint64_t g_tcp_limit = 0;
bool g_tcp_limit_known = false;

}  // namespace

uint64_t gh_mem_denied(void) { return g_denied.load(); }

static int tcp_mem_limit(uint64_t* used, uint64_t* total) {
  gemhook_request req;
  gemhook_response rsp;
  memset(&req, 0, sizeof(req));
  req.type = GEMHOOK_REQ_MEM_LIMIT;
  if (gh_rpc(&req, &rsp) != 0) {
    if (gh_cfg.exit_on_failure) exit(1);  // hook.cpp:358-361
    return -1;
  }
  *used = rsp.mem_used;
  *total = rsp.mem_total;
  g_tcp_limit = rsp.mem_total;
  g_tcp_limit_known = true;
  return 0;
}

static int tcp_mem_update(uint64_t bytes, int is_alloc) {
  gemhook_request req;
  gemhook_response rsp;
  memset(&req, 0, sizeof(req));
  req.type = GEMHOOK_REQ_MEM_UPDATE;
  req.bytes = bytes;
  req.is_alloc = is_alloc;
  if (gh_rpc(&req, &rsp) != 0) {
    if (gh_cfg.exit_on_failure) exit(1);  // hook.cpp:386-389
    return 1;
  }
  return rsp.verdict;
}

int gh_mem_reserve(uint64_t bytes) {
  if (!gh_live_get() || !gh_live_enabled()) return 1;
  int ok;
  if (gh_cfg.transport == 1) {
    ok = gemhook_pool_mem_reserve(gh_live_pool(), gh_live_slot(), bytes);
  } else {
    // gem-pmgr tests `used + bytes > limit` in size_t and would wrap for absurd sizes
    // (pod-manager.cpp:299); the reference's pre-hook (bytes > limit - used) catches those first.
    uint64_t used = 0, total = 0;
    if (!g_tcp_limit_known) tcp_mem_limit(&used, &total);
    if (g_tcp_limit_known && bytes > g_tcp_limit) ok = 0;
    else ok = tcp_mem_update(bytes, 1);
  }
  if (!ok) {
    g_denied.fetch_add(1);
    GH_INFO("gpu_mem cap: denied allocation of %llu bytes", (unsigned long long)bytes);
  }
  return ok;
}

void gh_mem_unreserve(uint64_t bytes) {
  if (!gh_live_get() || !gh_live_enabled()) return;
  if (gh_cfg.transport == 1) gemhook_pool_mem_release(gh_live_pool(), gh_live_slot(), bytes);
  else tcp_mem_update(bytes, 0);
}

void gh_mem_commit(uint64_t key, uint64_t bytes) {
  pthread_mutex_lock(&g_table.mu);
  g_table.put_nolock(key, bytes);
  pthread_mutex_unlock(&g_table.mu);
  g_local_used.fetch_add(bytes);
}

void gh_mem_free_key(uint64_t key) {
  uint64_t bytes = 0;
  pthread_mutex_lock(&g_table.mu);
  bool known = g_table.take_nolock(key, &bytes);
  pthread_mutex_unlock(&g_table.mu);
  if (!known) return;  // "Freeing unknown memory": ignored (hook.cpp:572-573)
  g_local_used.fetch_sub(bytes);
  gh_mem_unreserve(bytes);
}

void gh_mem_info(uint64_t* free_b, uint64_t* total_b) {
  uint64_t used = 0, total = 0;
  if (gh_live_get() && gh_live_enabled()) {
    if (gh_cfg.transport == 1) gemhook_pool_mem_info(gh_live_pool(), gh_live_slot(), &used, &total);
    else tcp_mem_limit(&used, &total);
  }
  if (free_b) *free_b = used < total ? total - used : 0;  // (a reload may have lowered the limit below what is in use)
  if (total_b) *total_b = total;
}

// what this process knows without asking anybody: the pool counters, or (TCP transport) its own share and the limit
// gem-pmgr last reported
void gh_mem_local(uint64_t* free_b, uint64_t* total_b) {
  uint64_t used = 0, total = 0;
  if (gh_live_enabled()) {
    if (gh_cfg.transport == 1 && gh_live_pool()) gemhook_pool_mem_info(gh_live_pool(), gh_live_slot(), &used, &total);
    else {
      used = g_local_used.load();
      total = g_tcp_limit_known ? g_tcp_limit : 0;
    }
  }
  if (free_b) *free_b = used < total ? total - used : 0;
  if (total_b) *total_b = total;
}
