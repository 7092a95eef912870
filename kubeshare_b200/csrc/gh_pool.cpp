// gh_pool.cpp -- the shared credit pool: one file-backed MAP_SHARED region per GPU that replaces the
// hook -> gem-pmgr -> gem-schd TCP round trips (reference hook.cpp:300-328, 425-446) for co-resident
// clients, and the token policy that runs inside it.
//
// What stays the reference's (bit-exact, see tests/test_pool_policy.py against the oracle):
//   * per-client adaptive quota        scheduler.cpp:160-174  (EMA 0.5, clamp [min_quota, max_frac*window])
//   * ledger of granted tokens          scheduler.cpp:144-153  Record(); :123-142 update_return_time()
//   * windowed usage with overlap split scheduler.cpp:281-367
//   * eligibility + ordering            scheduler.cpp:369-398, schd-priority.cpp:19-26
//   * ONE outstanding token per GPU     scheduler.cpp:461-529
//   * gpu_mem counter, requested bytes  pod-manager.cpp:295-313, hook.cpp:590-601
// What is new: there is no daemon in the decision path.  Whoever needs a decision takes a short
// spin lock in the shared region (held for microseconds), runs the policy over the shared ledger and
// publishes the grant into the winner's slot; a renewal on an uncontended GPU is a few hundred
// nanoseconds of shared-memory traffic and no context switch.  Clients that must wait (throttled, or
// another client holds the token) futex-wait on their own slot word.
//
// The region is plain memory: it can be cuMemHostRegister'ed (gh_hook.cpp does) so the device sees
// the same counters ("shared-pinned").
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <errno.h>
#include <fcntl.h>
#include <stddef.h>
#include <linux/futex.h>
#include <sched.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <vector>

#include "gh_internal.h"

namespace {

const uint64_t POOL_MAGIC = 0x314c4f4f504d4547ULL;  // "GEMPOOL1"
const uint32_t POOL_VERSION = 2;
const uint32_t MAX_ATTACH = 256;
const uint32_t LEDGER_CAP = 8192;
enum { ST_IDLE = 0, ST_WAITING = 1, ST_GRANTED = 2 };

struct alignas(64) Slot {
  // line 0: identity
  char name[64];
  // line 1: configuration + adaptive quota state (ClientInfo)
  double min_frac, max_frac;
  uint64_t mem_limit;
  double quota;   // quota_
  double burst;   // burst_
  uint32_t configured;
  uint32_t _pad0;
  double last_start, last_end;  // latest token of this client in the FULL history (never pruned)
  // line 2: request / grant mailbox
  std::atomic<uint32_t> state;
  uint32_t _pad1;
  double arrived_ms;
  double granted_quota;
  uint64_t req_seq;  // order of arrival (the candidates list is FIFO)
  uint64_t grants;
  double closed_ms;  // sum(end - start) over this client's finished tokens (full history)
  uint64_t _pad2[2];
  // line 3: counters other parties read
  std::atomic<uint64_t> mem_used;
  std::atomic<uint64_t> gpu_ns, launches;
  // pod-level token shared by the processes of one pod (gem-pmgr's role, pod-manager.cpp:97-101)
  double pod_quota;       // pod_quota
  int64_t pod_token_us;   // quota_updated_tp, microseconds since pool start
  double pod_overuse;     // pod_overuse_ms
  uint64_t _pad3[2];
};
static_assert(sizeof(Slot) == 256, "slot layout");

struct Span {
  int32_t slot;
  int32_t _pad;
  double start, end;
};

struct alignas(64) Header {
  uint64_t magic;
  uint32_t version, nslots_max;
  double base_quota, min_quota, window;
  int64_t start_ns;
  std::atomic<uint32_t> ready;
  std::atomic<uint32_t> nslots;
  // arbitration lock (owner pid, for dead-owner recovery)
  alignas(64) std::atomic<uint32_t> lock;
  uint32_t _padl;
  std::atomic<int64_t> lock_ns;
  // token state
  alignas(64) int32_t holder;  // slot holding the outstanding token or -1
  int32_t _padh;
  double deadline_ms;
  uint64_t next_req_seq;
  uint64_t total_grants;
  uint32_t ledger_len;
  uint32_t ledger_dropped;
  std::atomic<uint64_t> quota_stamp;  // mtime/size stamp of the quota file last loaded (gemhook_pool_sync_quota_file)
  uint64_t boot_id;  // CLOCK_MONOTONIC restarts at boot: a pool file that survived a reboot is re-initialised
};

// one entry per attached process; byte 0 of each entry is covered by an OFD lock held by the owner for its
// lifetime -- the kernel drops it when the process dies, in any container, which is how a dead client's
// bytes are found and reclaimed (the reference reclaims on socket close, pod-manager.cpp:533-545)
struct Attach {
  std::atomic<uint32_t> in_use;
  int32_t slot;
  std::atomic<uint64_t> bytes;
  double burst;  // client_burst_map entry (pod-manager.cpp:98)
  uint32_t pid;
  uint32_t _pad;
};
static_assert(sizeof(Attach) == 32, "attach layout");

struct Region {
  Header h;
  Slot slots[GEMHOOK_MAX_SLOTS];
  Span ledger[LEDGER_CAP];
  Attach attach[MAX_ATTACH];
};

uint64_t boot_id_hash() {
  uint64_t h = 1469598103934665603ULL;
  FILE* f = fopen("/proc/sys/kernel/random/boot_id", "r");
  if (f) {
    int c;
    while ((c = fgetc(f)) != EOF) h = (h ^ (uint64_t)(unsigned char)c) * 1099511628211ULL;
    fclose(f);
  }
  return h ? h : 1;
}

long futex(std::atomic<uint32_t>* addr, int op, uint32_t val, const struct timespec* ts) {
  return syscall(SYS_futex, (uint32_t*)addr, op, val, ts, nullptr, 0);
}

struct Stamp {
  int32_t slot;
  double t;  // negative = start
};
struct Ranked {
  double missing, remaining, usage, arrived;
  int slot;
};
// schd-priority.cpp:19-26
bool rank_before(const Ranked& a, const Ranked& b) {
  if (a.missing > 0 && b.missing > 0) return a.missing / (a.missing + a.usage) > b.missing / (b.missing + b.usage);
  if (a.missing > 0 && b.missing < 0) return true;
  if (a.missing < 0 && b.missing > 0) return false;
  return a.usage < b.usage;
}

}  // namespace

struct gemhook_pool {
  Region* r = nullptr;
  int fd = -1;
  bool anonymous = false;
  int attach_idx = -1;  // this handle's attachment (set by gemhook_pool_attach)

  int attach_lock(int idx, bool take) {
    if (fd < 0) return 0;
    struct flock fl;
    memset(&fl, 0, sizeof(fl));
    fl.l_type = take ? F_WRLCK : F_UNLCK;
    fl.l_whence = SEEK_SET;
    fl.l_start = (off_t)(offsetof(Region, attach) + (size_t)idx * sizeof(Attach));
    fl.l_len = 1;
    return fcntl(fd, F_OFD_SETLK, &fl);
  }
  bool attach_owner_alive(int idx) {
    if (fd < 0) return true;
    struct flock fl;
    memset(&fl, 0, sizeof(fl));
    fl.l_type = F_WRLCK;
    fl.l_whence = SEEK_SET;
    fl.l_start = (off_t)(offsetof(Region, attach) + (size_t)idx * sizeof(Attach));
    fl.l_len = 1;
    if (fcntl(fd, F_OFD_GETLK, &fl) != 0) return true;  // cannot tell: assume alive
    return fl.l_type != F_UNLCK;
  }

  void lock() {
    uint32_t me = (uint32_t)getpid();
    int spins = 0;
    for (;;) {
      uint32_t exp = 0;
      if (r->h.lock.compare_exchange_weak(exp, me, std::memory_order_acquire)) break;
      if (++spins > 2000) {
        // owner may have died inside the critical section (it lasts microseconds): steal after a full second --
        // long enough that a merely preempted owner is never robbed, short enough that a kill -9 landing exactly
        // inside the section does not wedge the GPU's clients for good
        int64_t t = r->h.lock_ns.load(std::memory_order_relaxed);
        if (t && gh_now_ns() - t > 1000000000LL) {
          if (r->h.lock.compare_exchange_strong(exp, me, std::memory_order_acquire)) break;
        }
        sched_yield();
        spins = 0;
      }
      __builtin_ia32_pause();
    }
    r->h.lock_ns.store(gh_now_ns(), std::memory_order_relaxed);
  }
  void unlock() {
    r->h.lock_ns.store(0, std::memory_order_relaxed);
    r->h.lock.store(0, std::memory_order_release);
  }

  // scheduler.cpp:281-367 -- prune the ledger and compute per-slot usage inside the window
  void window_usage(double now, double* usage, double& wsize, double& wstart) {
    Header& h = r->h;
    wsize = h.window;
    wstart = now - h.window;
    if (wstart < 0) wsize = now;
    uint32_t k = 0;
    for (uint32_t i = 0; i < h.ledger_len; i++)
      if (!(r->ledger[i].end < wstart)) r->ledger[k++] = r->ledger[i];
    h.ledger_len = k;

    std::vector<Stamp> st;
    st.reserve(2 * k);
    for (uint32_t i = 0; i < k; i++) {
      st.push_back({r->ledger[i].slot, -r->ledger[i].start});
      st.push_back({r->ledger[i].slot, r->ledger[i].end});
      usage[r->ledger[i].slot] = 0;
    }
    std::sort(st.begin(), st.end(), [](Stamp a, Stamp b) { return std::abs(a.t) < std::abs(b.t); });
    std::vector<int32_t> live;
    int live_cnt = 0;
    size_t j = 0;
    for (; j < st.size(); j++) {
      if (std::abs(st[j].t) <= wstart) {
        live_cnt++;
        live.push_back(st[j].slot);
      } else {
        break;
      }
    }
    double cur = wstart;
    for (size_t i = j; i < st.size(); i++) {
      for (size_t q = 0; q < live.size(); q++) usage[live[q]] += (std::abs(st[i].t) - cur) / live_cnt;
      if (st[i].t < 0) {
        live.push_back(st[i].slot);
        live_cnt++;
      } else {
        for (size_t q = 0; q < live.size(); q++)
          if (live[q] == st[i].slot) {
            live.erase(live.begin() + q);
            break;
          }
        live_cnt--;
      }
      cur = std::abs(st[i].t);
    }
  }

  // caller holds the lock
  int schedule_locked(double now, int* slot_out, double* quota_out, double* sleep_out) {
    Header& h = r->h;
    uint32_t n = h.nslots.load(std::memory_order_relaxed);
    if (h.holder >= 0) {
      // scheduler.cpp:501-521: wait until the holder asks again or its quota times out
      bool back = r->slots[h.holder].state.load(std::memory_order_relaxed) == ST_WAITING;
      if (!back && now < h.deadline_ms) {
        if (sleep_out) *sleep_out = h.deadline_ms - now;
        return -2;
      }
      h.holder = -1;
    }
    // candidates in arrival order
    int order[GEMHOOK_MAX_SLOTS];
    int nc = 0;
    for (uint32_t i = 0; i < n; i++)
      if (r->slots[i].state.load(std::memory_order_relaxed) == ST_WAITING) order[nc++] = (int)i;
    if (nc == 0) return -1;
    std::sort(order, order + nc, [&](int a, int b) { return r->slots[a].req_seq < r->slots[b].req_seq; });

    double usage[GEMHOOK_MAX_SLOTS];
    for (uint32_t i = 0; i < n; i++) usage[i] = 0;
    double wsize, wstart;
    window_usage(now, usage, wsize, wstart);

    int pick = -1;
    bool head_seen = false;  // scheduler.cpp:312-320: head of the queue with no recent history goes first
    for (uint32_t i = 0; i < h.ledger_len; i++)
      if (r->ledger[i].slot == order[0]) {
        head_seen = true;
        break;
      }
    if (!head_seen) {
      pick = order[0];
    } else {
      std::vector<Ranked> ok;
      for (int c = 0; c < nc; c++) {
        Slot& s = r->slots[order[c]];
        double limit = s.max_frac * wsize, require = s.min_frac * wsize;
        double missing = require - usage[order[c]], remaining = limit - usage[order[c]];
        if (remaining > 0) ok.push_back({missing, remaining, usage[order[c]], s.arrived_ms, order[c]});
      }
      if (ok.empty()) {  // scheduler.cpp:383-390
        if (sleep_out) *sleep_out = r->ledger[0].end - wstart;
        return 0;
      }
      std::sort(ok.begin(), ok.end(), rank_before);
      pick = ok[0].slot;
    }

    // get_quota (scheduler.cpp:160-174) + Record (scheduler.cpp:144-153)
    Slot& s = r->slots[pick];
    if (s.burst < 1e-9) {
      s.quota = h.base_quota;
    } else {
      s.quota = s.burst * 0.5 + s.quota * (1 - 0.5);
      s.quota = std::max(s.quota, h.min_quota);
      s.quota = std::min(s.quota, s.max_frac * h.window);
    }
    if (h.ledger_len == LEDGER_CAP) {  // cannot happen with sane quotas; keep the newest entries
      memmove(&r->ledger[0], &r->ledger[1], sizeof(Span) * (LEDGER_CAP - 1));
      h.ledger_len--;
      h.ledger_dropped++;
    }
    r->ledger[h.ledger_len++] = Span{pick, 0, now, now + s.quota};
    if (s.grants) s.closed_ms += s.last_end - s.last_start;
    s.last_start = now;
    s.last_end = now + s.quota;
    s.grants++;
    h.total_grants++;
    h.holder = pick;
    h.deadline_ms = now + s.quota;
    s.granted_quota = s.quota;
    s.state.store(ST_GRANTED, std::memory_order_release);
    if (slot_out) *slot_out = pick;
    if (quota_out) *quota_out = s.quota;
    return 1;
  }

  int64_t now_us() const { return (gh_now_ns() - r->h.start_ns) / 1000; }
  double now_ms() const { return (double)now_us() / 1e3; }  // scheduler.cpp:107-109
};

GH_EXPORT gemhook_pool* gemhook_pool_open(const char* path, int create, double base_quota_ms, double min_quota_ms,
                                          double window_ms, int64_t start_ns) {
  gemhook_pool* p = new gemhook_pool();
  size_t bytes = sizeof(Region);
  void* m = MAP_FAILED;
  bool fresh = false;
  if (!path || !*path) {  // private pool (tests, single-process use)
    m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    p->anonymous = true;
    fresh = true;
  } else {
    int fd = open(path, create ? (O_RDWR | O_CREAT) : O_RDWR, 0666);
    if (fd < 0) {
      gh_set_error("cannot open pool file %s: %s", path, strerror(errno));
      delete p;
      return nullptr;
    }
    struct stat st;
    fstat(fd, &st);
    if ((size_t)st.st_size < bytes) {
      if (!create || ftruncate(fd, (off_t)bytes) != 0) {
        gh_set_error("pool file %s is too small (%lld bytes) and may not be created", path, (long long)st.st_size);
        close(fd);
        delete p;
        return nullptr;
      }
    }
    m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    p->fd = fd;
  }
  if (m == MAP_FAILED) {
    gh_set_error("mmap of the credit pool failed: %s", strerror(errno));
    if (p->fd >= 0) close(p->fd);
    delete p;
    return nullptr;
  }
  p->r = (Region*)m;
  Header& h = p->r->h;
  // first opener initialises: claim with CAS on magic (file content starts as zeros)
  uint64_t zero = 0;
  std::atomic<uint64_t>* magic = reinterpret_cast<std::atomic<uint64_t>*>(&h.magic);
  if (fresh || (create && magic->compare_exchange_strong(zero, POOL_MAGIC))) {
    h.magic = POOL_MAGIC;
    h.version = POOL_VERSION;
    h.nslots_max = GEMHOOK_MAX_SLOTS;
    h.base_quota = base_quota_ms;
    h.min_quota = min_quota_ms;
    h.window = window_ms;
    h.start_ns = start_ns ? start_ns : gh_now_ns();
    h.holder = -1;
    h.boot_id = boot_id_hash();
    h.nslots.store(0);
    h.ready.store(1, std::memory_order_release);
  } else {
    for (int i = 0; i < 20000 && !h.ready.load(std::memory_order_acquire); i++) usleep(100);
    if (h.magic != POOL_MAGIC || !h.ready.load() || h.version != POOL_VERSION) {
      gh_set_error("%s is not an initialised gemhook credit pool", path);
      gemhook_pool_close(p);
      return nullptr;
    }
    if (h.boot_id != boot_id_hash()) {
      // the file outlived a reboot (hostPath): its clock origin, token holder, ledger, attachments and byte
      // counters describe processes that no longer exist.  Keep the configuration rows, drop the dynamic state.
      h.lock.store(0);
      p->lock();
      if (h.boot_id != boot_id_hash()) {
        uint32_t ns = h.nslots.load();
        for (uint32_t i = 0; i < ns; i++) {
          Slot& s = p->r->slots[i];
          s.quota = h.base_quota;
          s.burst = 0;
          s.last_start = s.last_end = s.closed_ms = 0;
          s.grants = 0;
          s.state.store(ST_IDLE);
          s.mem_used.store(0);
          s.gpu_ns.store(0);
          s.launches.store(0);
          s.pod_quota = 0;
          s.pod_token_us = 0;
          s.pod_overuse = 0;
        }
        memset((void*)p->r->attach, 0, sizeof(p->r->attach));
        h.ledger_len = 0;
        h.holder = -1;
        h.deadline_ms = 0;
        h.start_ns = gh_now_ns();
        h.boot_id = boot_id_hash();
      }
      p->unlock();
    }
  }
  return p;
}

GH_EXPORT void gemhook_pool_close(gemhook_pool* p) {
  if (!p) return;
  if (p->r) munmap(p->r, sizeof(Region));
  if (p->fd >= 0) close(p->fd);
  delete p;
}

void* gh_pool_region(gemhook_pool* p, size_t* bytes) {
  if (bytes) *bytes = sizeof(Region);
  return p ? (void*)p->r : nullptr;
}

// read_resource_config (scheduler.cpp:183-217): "N" then N rows "name c2 c3 mem"; a re-read replaces the
// client's ClientInfo, i.e. its adaptive quota restarts from the base quota; usage history is kept.
GH_EXPORT int gemhook_pool_load_config(gemhook_pool* p, const char* text, int swap_columns) {
  if (!p || !text) return -1;
  const char* c = text;
  char* end = nullptr;
  long n = strtol(c, &end, 10);
  if (end == c || n < 0) {
    gh_set_error("quota file: bad client count");
    return -1;
  }
  c = end;
  p->lock();
  Header& h = p->r->h;
  int loaded = 0;
  for (long i = 0; i < n; i++) {
    char name[64];
    double c2, c3;
    unsigned long long mem;
    int used = 0;
    if (sscanf(c, " %63s %lf %lf %llu%n", name, &c2, &c3, &mem, &used) != 4) break;
    c += used;
    uint32_t ns = h.nslots.load(std::memory_order_relaxed);
    int idx = -1;
    for (uint32_t s = 0; s < ns; s++)
      if (!strncmp(p->r->slots[s].name, name, sizeof(p->r->slots[s].name))) idx = (int)s;
    if (idx < 0) {
      if (ns >= GEMHOOK_MAX_SLOTS) {
        gh_set_error("quota file lists more than %d clients", GEMHOOK_MAX_SLOTS);
        break;
      }
      idx = (int)ns;
      memset((void*)&p->r->slots[idx], 0, sizeof(Slot));
      snprintf(p->r->slots[idx].name, sizeof(p->r->slots[idx].name), "%s", name);
      h.nslots.store(ns + 1, std::memory_order_release);
    }
    Slot& s = p->r->slots[idx];
    s.min_frac = swap_columns ? c3 : c2;
    s.max_frac = swap_columns ? c2 : c3;
    s.mem_limit = mem;
    s.quota = h.base_quota;
    s.burst = 0.0;
    s.configured = 1;
    loaded++;
  }
  p->unlock();
  return loaded == n ? (int)n : -1;
}

// Keep the pool in step with the quota file kubeshare-config rewrites (what gem-schd does with inotify,
// scheduler.cpp:219-265): cheap stat, reload only when the file's (mtime, size) stamp differs from the one recorded
// in the pool -- so of N co-resident clients only the first to notice reloads.  1 = reloaded, 0 = unchanged, -1 = error.
GH_EXPORT int gemhook_pool_sync_quota_file(gemhook_pool* p, const char* path, int swap_columns) {
  if (!p || !path || !*path) return -1;
  struct stat st;
  if (stat(path, &st) != 0) return -1;
  uint64_t stamp = ((uint64_t)st.st_mtim.tv_sec * 1000000000ULL + (uint64_t)st.st_mtim.tv_nsec) * 31ULL + (uint64_t)st.st_size + 1;
  if (p->r->h.quota_stamp.load(std::memory_order_acquire) == stamp) return 0;
  FILE* f = fopen(path, "r");
  if (!f) return -1;
  char* text = (char*)calloc(1, 1 << 16);
  size_t n = fread(text, 1, (1 << 16) - 1, f);
  fclose(f);
  text[n] = 0;
  int rc = gemhook_pool_load_config(p, text, swap_columns);
  free(text);
  if (rc < 0) return -1;  // half-written file: keep the old stamp, try again at the next renewal
  p->r->h.quota_stamp.store(stamp, std::memory_order_release);
  return 1;
}

GH_EXPORT int gemhook_pool_find(const gemhook_pool* p, const char* name) {
  uint32_t ns = p->r->h.nslots.load(std::memory_order_acquire);
  for (uint32_t s = 0; s < ns; s++)
    if (!strncmp(p->r->slots[s].name, name, sizeof(p->r->slots[s].name))) return (int)s;
  return -1;
}
GH_EXPORT int gemhook_pool_nslots(const gemhook_pool* p) { return (int)p->r->h.nslots.load(std::memory_order_acquire); }

// handle_message(REQ_QUOTA) (scheduler.cpp:417-429): update_return_time + set_burst + enqueue
static void request_locked(gemhook_pool* p, int slot, double now, double overuse, double burst) {
  Header& h = p->r->h;
  Slot& s = p->r->slots[slot];
  for (uint32_t i = h.ledger_len; i-- > 0;) {
    if (p->r->ledger[i].slot == slot) {
      p->r->ledger[i].end = std::min(now, p->r->ledger[i].end + overuse);
      break;
    }
  }
  if (s.grants) s.last_end = std::min(now, s.last_end + overuse);
  s.burst = burst;
  s.arrived_ms = now;
  s.req_seq = ++h.next_req_seq;
  s.state.store(ST_WAITING, std::memory_order_release);
}

GH_EXPORT int gemhook_pool_request(gemhook_pool* p, int slot, double now_ms, double overuse_ms, double burst_ms) {
  if (!p || slot < 0 || slot >= gemhook_pool_nslots(p)) return -1;
  p->lock();
  request_locked(p, slot, now_ms, overuse_ms, burst_ms);
  p->unlock();
  return 0;
}

GH_EXPORT int gemhook_pool_schedule(gemhook_pool* p, double now_ms, int* slot_out, double* quota_out, double* sleep_ms_out) {
  p->lock();
  int rc = p->schedule_locked(now_ms, slot_out, quota_out, sleep_ms_out);
  p->unlock();
  return rc;
}

GH_EXPORT double gemhook_pool_usage(gemhook_pool* p, int slot, double now_ms) {
  double usage[GEMHOOK_MAX_SLOTS] = {0};
  double a, b;
  p->lock();
  p->window_usage(now_ms, usage, a, b);
  p->unlock();
  return usage[slot];
}

// Observers take the arbitration lock too: the ledger and the per-slot token fields are plain (non-atomic) data
// owned by whoever holds it (ThreadSanitizer-clean, tests/test_sanitizers.py).
GH_EXPORT size_t gemhook_pool_history(const gemhook_pool* cp, int* slots, double* starts, double* ends, size_t cap) {
  gemhook_pool* p = const_cast<gemhook_pool*>(cp);
  p->lock();
  size_t n = p->r->h.ledger_len;
  for (size_t i = 0; i < n && i < cap; i++) {
    if (slots) slots[i] = p->r->ledger[i].slot;
    if (starts) starts[i] = p->r->ledger[i].start;
    if (ends) ends[i] = p->r->ledger[i].end;
  }
  p->unlock();
  return n;
}

GH_EXPORT double gemhook_pool_accumulated_ms(const gemhook_pool* cp, int slot) {
  gemhook_pool* p = const_cast<gemhook_pool*>(cp);
  p->lock();
  const Slot& s = p->r->slots[slot];
  double v = s.grants ? s.closed_ms + (s.last_end - s.last_start) : 0.0;
  p->unlock();
  return v;
}

static int pod_launch_locked(gemhook_pool* p, int slot, int attach_idx, int64_t now_us, double overuse, double burst,
                             double* fwd_overuse, double* fwd_burst, double* remain);
static double pod_granted_locked(gemhook_pool* p, int slot, int64_t now_us, double quota);

// Live acquisition: post the request, then arbitrate/wait until OUR slot is granted.
GH_EXPORT double gemhook_pool_acquire(gemhook_pool* p, int slot, double overuse_ms, double burst_ms) {
  return gemhook_pool_acquire_ex(p, slot, overuse_ms, burst_ms, nullptr);
}
GH_EXPORT double gemhook_pool_acquire_ex(gemhook_pool* p, int slot, double overuse_ms, double burst_ms, int* forwarded) {
  Slot& me = p->r->slots[slot];
  if (forwarded) *forwarded = 1;
  p->lock();
  double fo = overuse_ms, fb = burst_ms, remain = 0.0;
  if (!pod_launch_locked(p, slot, p->attach_idx, p->now_us(), overuse_ms, burst_ms, &fo, &fb, &remain)) {
    p->unlock();
    if (forwarded) *forwarded = 0;
    return remain;  // the pod's token still covers this burst (pod-manager.cpp:472)
  }
  request_locked(p, slot, p->now_ms(), fo, fb);
  p->unlock();
  int idle_rounds = 0;
  for (;;) {
    int who = -1;
    double q = 0, sleep_ms = 0;
    p->lock();
    int rc = p->schedule_locked(p->now_ms(), &who, &q, &sleep_ms);
    p->unlock();
    if (rc == 1 && who != slot) futex(&p->r->slots[who].state, FUTEX_WAKE, 1, nullptr);  // wake the winner
    if (me.state.load(std::memory_order_acquire) == ST_GRANTED) {
      double got = me.granted_quota;
      me.state.store(ST_IDLE, std::memory_order_release);
      p->lock();
      got = pod_granted_locked(p, slot, p->now_us(), got);
      p->unlock();
      return got;
    }
    if (++idle_rounds % 8 == 0) gemhook_pool_reap(p);  // a dead holder must not stall everybody until its deadline
    if (me.state.load(std::memory_order_acquire) == ST_IDLE) {
      // our request is gone without us having consumed a grant: another process of the same pod took it (they share
      // the slot's mailbox), or a reaper cleared it.  Start over: the pod's token may already cover us.
      p->lock();
      if (!pod_launch_locked(p, slot, p->attach_idx, p->now_us(), overuse_ms, burst_ms, &fo, &fb, &remain)) {
        p->unlock();
        if (forwarded) *forwarded = 0;
        return remain;
      }
      request_locked(p, slot, p->now_ms(), fo, fb);
      p->unlock();
      continue;
    }
    // someone else holds the token, or everyone is throttled: sleep on our own slot word until the hint
    // expires or a granter wakes us.  Short waits spin (no context switch on a quick hand-over).
    double wait_ms = (rc == 0 || rc == -2) ? sleep_ms : 0.2;
    if (wait_ms < 0.05) {
      for (int i = 0; i < 200 && me.state.load(std::memory_order_acquire) != ST_GRANTED; i++) __builtin_ia32_pause();
      continue;
    }
    if (wait_ms > 50.0) wait_ms = 50.0;  // re-evaluate periodically (config reloads, dead holders)
    struct timespec ts;
    ts.tv_sec = (time_t)(wait_ms / 1e3);
    ts.tv_nsec = (long)((wait_ms - ts.tv_sec * 1e3) * 1e6);
    futex(&me.state, FUTEX_WAIT, ST_WAITING, &ts);
  }
}

// ---- attachments ------------------------------------------------------------------------------------------
GH_EXPORT int gemhook_pool_attach(gemhook_pool* p, int slot) {
  if (!p || slot < 0) return -1;
  for (uint32_t i = 0; i < MAX_ATTACH; i++) {
    uint32_t exp = 0;
    if (p->r->attach[i].in_use.compare_exchange_strong(exp, 1u)) {
      Attach& a = p->r->attach[i];
      a.slot = slot;
      a.bytes.store(0);
      a.burst = 0.0;
      a.pid = (uint32_t)getpid();
      p->attach_lock((int)i, true);
      p->attach_idx = (int)i;
      return (int)i;
    }
  }
  gh_set_error("credit pool: more than %u attached processes", MAX_ATTACH);
  return -1;
}

GH_EXPORT void gemhook_pool_detach(gemhook_pool* p) {
  if (!p || p->attach_idx < 0) return;
  Attach& a = p->r->attach[p->attach_idx];
  uint64_t left = a.bytes.exchange(0);
  if (left) p->r->slots[a.slot].mem_used.fetch_sub(left, std::memory_order_acq_rel);  // exit without freeing
  p->attach_lock(p->attach_idx, false);
  a.in_use.store(0, std::memory_order_release);
  p->attach_idx = -1;
}

// reclaim what dead processes left behind: their bytes, and the token if one of them held it
GH_EXPORT int gemhook_pool_reap(gemhook_pool* p) {
  if (!p) return 0;
  int reaped = 0;
  for (uint32_t i = 0; i < MAX_ATTACH; i++) {
    Attach& a = p->r->attach[i];
    if (!a.in_use.load(std::memory_order_acquire) || (int)i == p->attach_idx) continue;
    if (p->attach_owner_alive((int)i)) continue;
    p->lock();
    if (a.in_use.load() && !p->attach_owner_alive((int)i)) {
      uint64_t left = a.bytes.exchange(0);
      if (left) p->r->slots[a.slot].mem_used.fetch_sub(left, std::memory_order_acq_rel);
      bool others = false;
      for (uint32_t j = 0; j < MAX_ATTACH; j++)
        if (j != i && p->r->attach[j].in_use.load() && p->r->attach[j].slot == a.slot) others = true;
      if (!others) {
        if (p->r->h.holder == a.slot) p->r->h.holder = -1;
        uint32_t st = p->r->slots[a.slot].state.load();
        if (st != ST_IDLE) p->r->slots[a.slot].state.store(ST_IDLE);
      }
      a.in_use.store(0, std::memory_order_release);
      reaped++;
    }
    p->unlock();
  }
  return reaped;
}

// ---- pod-level token (gem-pmgr hook_kernel_launch, pod-manager.cpp:316-473) -----------------------------------
// Processes of one pod share the pod's token: a request is answered locally with the REMAINING pod quota unless
// `elapsed + burst > pod_quota`, in which case it is forwarded to the scheduler with the pod's maximum overuse and
// the maximum burst over its processes.  Returns 1 = forward (fwd_* filled), 0 = answered (*remain_ms).
static int pod_launch_locked(gemhook_pool* p, int slot, int attach_idx, int64_t now_us, double overuse, double burst,
                             double* fwd_overuse, double* fwd_burst, double* remain) {
  Slot& s = p->r->slots[slot];
  s.pod_overuse = std::max(overuse, s.pod_overuse);
  if (attach_idx >= 0) p->r->attach[attach_idx].burst = burst;
  double elapsed = (double)(now_us - s.pod_token_us) / 1e3;
  if (elapsed + burst > s.pod_quota) {
    double mx = attach_idx >= 0 ? 0.0 : burst;
    for (uint32_t i = 0; i < MAX_ATTACH; i++)
      if (p->r->attach[i].in_use.load(std::memory_order_relaxed) && p->r->attach[i].slot == slot)
        mx = std::max(p->r->attach[i].burst, mx);
    if (fwd_overuse) *fwd_overuse = s.pod_overuse;
    if (fwd_burst) *fwd_burst = mx;
    return 1;
  }
  if (remain) *remain = s.pod_quota - elapsed;
  return 0;
}
static double pod_granted_locked(gemhook_pool* p, int slot, int64_t now_us, double quota) {
  Slot& s = p->r->slots[slot];
  s.pod_quota = quota;
  s.pod_token_us = now_us;
  s.pod_overuse = 0.0;
  return s.pod_quota - 0.0;
}
GH_EXPORT int gemhook_pool_pod_launch(gemhook_pool* p, int slot, int64_t now_us, double overuse_ms, double burst_ms,
                                      double* fwd_overuse_ms, double* fwd_burst_ms, double* remain_ms) {
  p->lock();
  int rc = pod_launch_locked(p, slot, p->attach_idx, now_us, overuse_ms, burst_ms, fwd_overuse_ms, fwd_burst_ms, remain_ms);
  p->unlock();
  return rc;
}
GH_EXPORT double gemhook_pool_pod_granted(gemhook_pool* p, int slot, int64_t now_us, double quota_ms) {
  p->lock();
  double r = pod_granted_locked(p, slot, now_us, quota_ms);
  p->unlock();
  return r;
}

// A client that is going away (process exit) hands its token back instead of letting the scheduler wait
// for the quota to time out (the reference can only time out: scheduler.cpp:507-510).  The ledger entry is
// closed exactly as a returning client's would be (update_return_time with zero overuse).
GH_EXPORT void gemhook_pool_release(gemhook_pool* p, int slot) {
  if (!p || slot < 0) return;
  p->lock();
  Header& h = p->r->h;
  double now = p->now_ms();
  int who = -1;
  if (h.holder == slot) {
    for (uint32_t i = h.ledger_len; i-- > 0;)
      if (p->r->ledger[i].slot == slot) {
        p->r->ledger[i].end = std::min(now, p->r->ledger[i].end);
        break;
      }
    Slot& s = p->r->slots[slot];
    if (s.grants) s.last_end = std::min(now, s.last_end);
    s.pod_quota = 0.0;  // the pod-level token is gone with it: the next request must be forwarded
    h.holder = -1;
    double q, sl;
    p->schedule_locked(now, &who, &q, &sl);
  }
  if (p->r->slots[slot].state.load() == ST_WAITING) p->r->slots[slot].state.store(ST_IDLE);
  p->unlock();
  if (who >= 0) futex(&p->r->slots[who].state, FUTEX_WAKE, 1, nullptr);
}

// The outstanding token is declared timed out (what gem-schd concludes when its timedwait on the holder
// returns ETIMEDOUT, scheduler.cpp:507-510) without touching the ledger: trace replays use it to decouple
// decisions from wall time, a node agent can use it to revoke the token of a client it knows is dead.
GH_EXPORT void gemhook_pool_expire_token(gemhook_pool* p) {
  if (!p) return;
  p->lock();
  p->r->h.holder = -1;
  p->unlock();
}

// 1 if some OTHER client is waiting for the token right now (lock-free peek, used by the yield-on-idle option)
GH_EXPORT int gemhook_pool_others_waiting(const gemhook_pool* p, int slot) {
  uint32_t n = p->r->h.nslots.load(std::memory_order_acquire);
  for (uint32_t i = 0; i < n; i++)
    if ((int)i != slot && p->r->slots[i].state.load(std::memory_order_relaxed) == ST_WAITING) return 1;
  return 0;
}

// ---- gpu_mem cap: integer exact, requested bytes (hook.cpp:590-617, pod-manager.cpp:295-313) -----------
GH_EXPORT int gemhook_pool_mem_reserve(gemhook_pool* p, int slot, uint64_t bytes) {
  Slot& s = p->r->slots[slot];
  uint64_t used = s.mem_used.load(std::memory_order_relaxed);
  for (;;) {
    // reference pre-hook: remain = limit - used (size_t arithmetic); deny iff bytes > remain
    uint64_t remain = s.mem_limit - used;
    if (bytes > remain) return 0;
    if (s.mem_used.compare_exchange_weak(used, used + bytes, std::memory_order_acq_rel)) {
      if (p->attach_idx >= 0 && p->r->attach[p->attach_idx].slot == slot)
        p->r->attach[p->attach_idx].bytes.fetch_add(bytes, std::memory_order_relaxed);
      return 1;
    }
  }
}
GH_EXPORT void gemhook_pool_mem_release(gemhook_pool* p, int slot, uint64_t bytes) {
  p->r->slots[slot].mem_used.fetch_sub(bytes, std::memory_order_acq_rel);
  if (p->attach_idx >= 0 && p->r->attach[p->attach_idx].slot == slot)
    p->r->attach[p->attach_idx].bytes.fetch_sub(bytes, std::memory_order_relaxed);
}
GH_EXPORT void gemhook_pool_mem_info(const gemhook_pool* p, int slot, uint64_t* used, uint64_t* limit) {
  if (used) *used = p->r->slots[slot].mem_used.load(std::memory_order_acquire);
  if (limit) *limit = p->r->slots[slot].mem_limit;
}
GH_EXPORT int gemhook_pool_slot_info(const gemhook_pool* cp, int slot, gemhook_slot_info* out) {
  gemhook_pool* p = const_cast<gemhook_pool*>(cp);
  if (!p || !out || slot < 0 || slot >= (int)p->r->h.nslots.load(std::memory_order_acquire)) return -1;
  p->lock();
  const Slot& s = p->r->slots[slot];
  memset(out, 0, sizeof(*out));
  snprintf(out->name, sizeof(out->name), "%s", s.name);
  out->min_frac = s.min_frac;
  out->max_frac = s.max_frac;
  out->mem_limit = s.mem_limit;
  out->mem_used = s.mem_used.load(std::memory_order_relaxed);
  out->gpu_ns = s.gpu_ns.load(std::memory_order_relaxed);
  out->launches = s.launches.load(std::memory_order_relaxed);
  out->tokens = s.grants;
  out->quota_ms = s.quota;
  out->accumulated_ms = s.grants ? s.closed_ms + (s.last_end - s.last_start) : 0.0;
  out->holds_token = p->r->h.holder == slot ? 1 : 0;
  out->waiting = s.state.load(std::memory_order_relaxed) == ST_WAITING ? 1 : 0;
  p->unlock();
  return 0;
}

void gh_pool_add_usage(gemhook_pool* p, int slot, uint64_t gpu_ns, uint64_t launches) {
  p->r->slots[slot].gpu_ns.fetch_add(gpu_ns, std::memory_order_relaxed);
  p->r->slots[slot].launches.fetch_add(launches, std::memory_order_relaxed);
}

// hook.cpp:638-680: bytes charged for arrays.  CUarray_format: U8 0x01, U16 0x02, U32 0x03, S8 0x08,
// S16 0x09, S32 0x0a, HALF 0x10, FLOAT 0x20; any other format is outside the reference's switch
// (undefined there) and is charged 0 bytes here.
GH_EXPORT uint64_t gemhook_array_bytes(uint64_t w, uint64_t h, uint64_t d, uint32_t channels, uint32_t format, int is3d) {
  uint64_t fs;
  switch (format) {
    case 0x01: case 0x08: fs = 1; break;
    case 0x02: case 0x09: case 0x10: fs = 2; break;
    case 0x03: case 0x0a: case 0x20: fs = 4; break;
    default: fs = 0; break;
  }
  return (is3d ? w * h * d * channels : w * h * channels) * fs;
}

// Opt-in rule for mipmapped arrays (GEMHOOK_ACCOUNT_MANAGED=1; the reference charges nothing, hook.cpp:682-694): the
// array rule above applied to every level, extents halving (floor, at least 1) from one level to the next.
GH_EXPORT uint64_t gemhook_mipmap_bytes(uint64_t w, uint64_t h, uint64_t d, uint32_t channels, uint32_t format, uint32_t levels) {
  uint64_t total = 0;
  for (uint32_t l = 0; l < levels; l++) {
    uint64_t lw = w >> l ? w >> l : 1, lh = h ? (h >> l ? h >> l : 1) : 0, ld = d ? (d >> l ? d >> l : 1) : 0;
    total += gemhook_array_bytes(lw, lh ? lh : 1, ld ? ld : 1, channels, format, 1);
  }
  return total;
}
