// gh_pool.cpp -- the shared credit pool: one file-backed MAP_SHARED region per GPU that replaces the
// hook -> gem-pmgr -> gem-schd TCP round trips (reference hook.cpp:300-328, 425-446) for co-resident
// clients, and the token policy that runs inside it.
//
// What stays the reference's (bit-exact, see tests/test_pool_policy.py against the oracle and the goldens):
//   * per-client adaptive quota        scheduler.cpp:160-174  (EMA 0.5, clamp [min_quota, max_frac*window])
//   * ledger of granted tokens          scheduler.cpp:144-153  Record(); :123-142 update_return_time()
//   * windowed usage with overlap split scheduler.cpp:281-367
//   * eligibility + ordering            scheduler.cpp:369-398, schd-priority.cpp:19-26
//   * ONE outstanding token per GPU     scheduler.cpp:461-529
//   * pod-level token                   pod-manager.cpp:316-473
//   * gpu_mem counter, requested bytes  pod-manager.cpp:295-313, hook.cpp:590-601
// (window_usage() and rank_before() follow scheduler.cpp:274-367 and schd-priority.cpp:19-26 step by step -- the
//  negative-timestamp trick, the overlap split, the evaluation order: bit-exact ledgers leave no freedom there.)
//
// What is new: there is no daemon, NO LOCK and no context switch in the decision path.
//
// LOCK-FREE BY CONSTRUCTION.  All policy state (token holder, deadline, per-client quota/mailbox/pod token, ledger)
// lives in a versioned STATE BLOCK.  The region holds NBLK such blocks; the 64-bit word `cur` = (sequence << 8 |
// block index) names the current one.  Every mutation is a transaction:
//     claim a free block (one CAS)  ->  copy the current block into it, validate `cur` did not move  ->  run the
//     SERIAL policy code on the private copy  ->  publish with ONE compare-and-swap on `cur`  ->  free the old block.
// A failed CAS means somebody else's transaction went through (system-wide progress: lock-free); the loser
// re-copies and retries.  A process that is SIGKILLed anywhere -- before, between or after those steps -- never
// blocks anybody: at worst it leaks a claimed block, which is recognised (owner's liveness byte released, or its
// publication attempt provably resolved) and recycled.  Readers (exporters, gem-poolctl, the hot gpu_mem path)
// never write anything: they copy / peek and validate against `cur`, so a Prometheus scrape cannot delay a token
// hand-over.  Because transactions run the unchanged serial code on a consistent snapshot, the ledger arithmetic
// stays bit-exact with the reference.
// Clients that must wait (throttled, or another client holds the token) futex-wait on their slot's wake word;
// short waits spin.  The region is plain memory: its first pages (header, per-slot counters) are cuMemHostRegister'ed
// by the hook so the device sees the same counters ("shared-pinned").
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <errno.h>
#include <fcntl.h>
#include <limits.h>
#include <linux/futex.h>
#include <math.h>
#include <sched.h>
#include <stddef.h>
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

#ifdef GEMHOOK_FAULT_INJECTION
// tests/native/pool_kill.cpp: die (SIGKILL) at the n-th pass through a named point of the transaction protocol
#include <signal.h>
extern "C" int gemhook_fault_point;  // 0 = off
extern "C" long gemhook_fault_countdown;
#define GH_FAULT(pt)                                                                   \
  do {                                                                                 \
    if (gemhook_fault_point == (pt) && --gemhook_fault_countdown <= 0) raise(SIGKILL); \
  } while (0)
#else
#define GH_FAULT(pt) do { } while (0)
#endif
enum { FP_AFTER_CLAIM = 1, FP_MID_POLICY = 2, FP_BEFORE_CAS = 3, FP_AFTER_CAS = 4 };

#define GH_NO_TSAN __attribute__((no_sanitize("thread")))

namespace {

const uint64_t POOL_MAGIC = 0x324c4f4f504d4547ULL;  // "GEMPOOL2"
const uint32_t POOL_VERSION = 3;
const uint32_t MAX_ATTACH = 256;
const uint32_t LEDGER_CAP = 8192;
const uint32_t NBLK = 16;
const uint32_t NO_OWNER = 0xffffu;
enum { ST_IDLE = 0, ST_WAITING = 1, ST_GRANTED = 2 };

// ---- versioned policy state -----------------------------------------------------------------------------------
struct PSlot {
  char name[64];
  double min_frac, max_frac;
  uint64_t mem_limit;
  double quota;   // quota_  (ClientInfo, scheduler.h:26-60)
  double burst;   // burst_
  uint32_t configured;
  uint32_t listed;              // present in the quota file as last loaded
  double last_start, last_end;  // latest token of this client in the FULL history (never pruned)
  // request / grant mailbox
  uint32_t state;
  int32_t poster;  // attachment that posted the request in flight (-1: none / not attached)
  double arrived_ms;
  double granted_quota;
  uint64_t req_seq;  // order of arrival (the candidates list is FIFO)
  uint64_t grants;
  double closed_ms;  // sum(end - start) over this client's finished tokens (full history)
  // pod-level token shared by the processes of one pod (gem-pmgr's role, pod-manager.cpp:97-101)
  double pod_quota;      // pod_quota
  int64_t pod_token_us;  // quota_updated_tp, microseconds since pool start
  double pod_overuse;    // pod_overuse_ms
};
static_assert(sizeof(PSlot) == 200, "policy slot layout");

struct Span {
  int32_t slot;
  int32_t _pad;
  double start, end;
};

struct State {
  int32_t holder;  // slot holding the outstanding token or -1
  uint32_t nslots;
  double deadline_ms;
  uint64_t next_req_seq;
  uint64_t total_grants;
  uint32_t ledger_len;
  uint32_t ledger_dropped;
  uint64_t quota_stamp;  // (mtime, size) stamp of the quota file this configuration was loaded from; 0 = loaded from text
  uint64_t _pad;
  PSlot slots[GEMHOOK_MAX_SLOTS];
  Span ledger[LEDGER_CAP];
};

// ---- unversioned shared words ---------------------------------------------------------------------------------
struct alignas(64) SlotShared {  // hot counters other parties (and the device) read; one cache line per slot
  std::atomic<uint64_t> mem_used;
  std::atomic<uint64_t> mem_limit_mirror;  // copy of the slot's mem_limit for the device-side mirror (informational)
  std::atomic<uint64_t> gpu_ns, launches;
  std::atomic<uint32_t> wake;  // futex word: bumped whenever something happens that this slot's waiters care about
  uint32_t _pad[7];
};
static_assert(sizeof(SlotShared) == 64, "shared slot layout");

struct alignas(64) Header {
  uint64_t magic;
  uint32_t version, nslots_max;
  double base_quota, min_quota, window;
  std::atomic<int64_t> start_ns;
  std::atomic<uint32_t> ready;
  uint32_t _pad0;
  std::atomic<uint64_t> boot_id;  // CLOCK_MONOTONIC restarts at boot: a pool file that survived a reboot is re-initialised
  alignas(64) std::atomic<uint64_t> cur;  // (sequence << 8) | index of the current state block
  alignas(64) std::atomic<uint64_t> quota_stamp;  // mtime/size stamp of the quota file last loaded
  std::atomic<uint64_t> commits, conflicts, recycled;
};

// one entry per attached handle; byte 0 of each entry is covered by an OFD lock held by the owner for its
// lifetime -- the kernel drops it when the process dies, in any container, which is how a dead client's
// bytes, token and claimed state blocks are found and reclaimed (the reference reclaims on socket close,
// pod-manager.cpp:533-545)
struct Attach {
  std::atomic<uint32_t> in_use;
  std::atomic<int32_t> slot;  // -1: observer (a tool or a test that only needs a liveness byte for its transactions)
  std::atomic<uint64_t> bytes;
  std::atomic<uint64_t> burst_bits;  // client_burst_map entry (pod-manager.cpp:98), a double
  uint32_t pid;
  uint32_t _pad;
};
static_assert(sizeof(Attach) == 32, "attach layout");

// block control word: 0 = FREE, else (1 << 63) | owner attachment << 46 | publication sequence (46 bits)
inline uint64_t ctl_make(uint32_t owner, uint64_t pub) { return (1ull << 63) | ((uint64_t)(owner & 0xffffu) << 46) | (pub & ((1ull << 46) - 1)); }
inline uint32_t ctl_owner(uint64_t w) { return (uint32_t)((w >> 46) & 0xffffu); }
inline uint64_t ctl_pub(uint64_t w) { return w & ((1ull << 46) - 1); }

struct Region {
  Header h;
  SlotShared shared[GEMHOOK_MAX_SLOTS];
  Attach attach[MAX_ATTACH];
  alignas(64) std::atomic<uint64_t> bctl[NBLK];
  alignas(4096) State blocks[NBLK];
};

uint64_t boot_id_hash() {
  uint64_t h = 1469598103934665603ULL;
  FILE* f = fopen("/proc/sys/kernel/random/boot_id", "r");
  if (f) {
    int c;
    while ((c = fgetc(f)) != EOF) h = (h ^ (uint64_t)(unsigned char)c) * 1099511628211ULL;
    fclose(f);
  }
  return h > 1 ? h : 2;
}

long futex(std::atomic<uint32_t>* addr, int op, uint32_t val, const struct timespec* ts) {
  return syscall(SYS_futex, (uint32_t*)addr, op, val, ts, nullptr, 0);
}

inline double bits_to_double(uint64_t b) {
  double d;
  memcpy(&d, &b, sizeof(d));
  return d;
}
inline uint64_t double_to_bits(double d) {
  uint64_t b;
  memcpy(&b, &d, sizeof(b));
  return b;
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

// ---- the serial policy, on ONE state block (a transaction's private copy) --------------------------------------
struct Policy {
  State& s;
  const Header& h;
  bool dirty = false;
  Policy(State& st, const Header& hd) : s(st), h(hd) {}

  // scheduler.cpp:281-367 -- prune the ledger and compute per-slot usage inside the window
  void window_usage(double now, double* usage, double& wsize, double& wstart) {
    wsize = h.window;
    wstart = now - h.window;
    if (wstart < 0) wsize = now;
    uint32_t k = 0;
    for (uint32_t i = 0; i < s.ledger_len; i++)
      if (!(s.ledger[i].end < wstart)) s.ledger[k++] = s.ledger[i];
    if (k != s.ledger_len) dirty = true;
    s.ledger_len = k;

    std::vector<Stamp> st;
    st.reserve(2 * k);
    for (uint32_t i = 0; i < k; i++) {
      st.push_back({s.ledger[i].slot, -s.ledger[i].start});
      st.push_back({s.ledger[i].slot, s.ledger[i].end});
      usage[s.ledger[i].slot] = 0;
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

  // schedule_daemon_func + select_candidate (scheduler.cpp:461-529, 274-399): one decision
  int schedule(double now, int* slot_out, double* quota_out, double* sleep_out) {
    uint32_t n = s.nslots;
    if (s.holder >= 0) {
      // scheduler.cpp:501-521: wait until the holder asks again or its quota times out
      bool back = s.slots[s.holder].state == ST_WAITING;
      if (!back && now < s.deadline_ms) {
        if (sleep_out) *sleep_out = s.deadline_ms - now;
        return -2;
      }
      s.holder = -1;
      dirty = true;
    }
    // candidates in arrival order
    int order[GEMHOOK_MAX_SLOTS];
    int nc = 0;
    for (uint32_t i = 0; i < n; i++)
      if (s.slots[i].state == ST_WAITING) order[nc++] = (int)i;
    if (nc == 0) return -1;
    std::sort(order, order + nc, [&](int a, int b) { return s.slots[a].req_seq < s.slots[b].req_seq; });

    double usage[GEMHOOK_MAX_SLOTS];
    for (uint32_t i = 0; i < n; i++) usage[i] = 0;
    double wsize, wstart;
    window_usage(now, usage, wsize, wstart);

    int pick = -1;
    bool head_seen = false;  // scheduler.cpp:312-320: head of the queue with no recent history goes first
    for (uint32_t i = 0; i < s.ledger_len; i++)
      if (s.ledger[i].slot == order[0]) {
        head_seen = true;
        break;
      }
    if (!head_seen) {
      pick = order[0];
    } else {
      std::vector<Ranked> ok;
      for (int c = 0; c < nc; c++) {
        PSlot& cs = s.slots[order[c]];
        double limit = cs.max_frac * wsize, require = cs.min_frac * wsize;
        double missing = require - usage[order[c]], remaining = limit - usage[order[c]];
        if (remaining > 0) ok.push_back({missing, remaining, usage[order[c]], cs.arrived_ms, order[c]});
      }
      if (ok.empty()) {  // scheduler.cpp:383-390
        if (sleep_out) *sleep_out = s.ledger[0].end - wstart;
        return 0;
      }
      std::sort(ok.begin(), ok.end(), rank_before);
      pick = ok[0].slot;
    }

    // get_quota (scheduler.cpp:160-174) + Record (scheduler.cpp:144-153)
    PSlot& ps = s.slots[pick];
    if (ps.burst < 1e-9) {
      ps.quota = h.base_quota;
    } else {
      ps.quota = ps.burst * 0.5 + ps.quota * (1 - 0.5);
      ps.quota = std::max(ps.quota, h.min_quota);
      ps.quota = std::min(ps.quota, ps.max_frac * h.window);
    }
    if (s.ledger_len == LEDGER_CAP) {  // cannot happen with sane quotas; keep the newest entries
      memmove(&s.ledger[0], &s.ledger[1], sizeof(Span) * (LEDGER_CAP - 1));
      s.ledger_len--;
      s.ledger_dropped++;
    }
    s.ledger[s.ledger_len++] = Span{pick, 0, now, now + ps.quota};
    if (ps.grants) ps.closed_ms += ps.last_end - ps.last_start;
    ps.last_start = now;
    ps.last_end = now + ps.quota;
    ps.grants++;
    s.total_grants++;
    s.holder = pick;
    s.deadline_ms = now + ps.quota;
    ps.granted_quota = ps.quota;
    ps.state = ST_GRANTED;
    dirty = true;
    if (slot_out) *slot_out = pick;
    if (quota_out) *quota_out = ps.quota;
    return 1;
  }

  // handle_message(REQ_QUOTA) (scheduler.cpp:417-429): update_return_time + set_burst + enqueue
  void request(int slot, double now, double overuse, double burst, int poster) {
    PSlot& ps = s.slots[slot];
    for (uint32_t i = s.ledger_len; i-- > 0;) {
      if (s.ledger[i].slot == slot) {
        s.ledger[i].end = std::min(now, s.ledger[i].end + overuse);
        break;
      }
    }
    if (ps.grants) ps.last_end = std::min(now, ps.last_end + overuse);
    ps.burst = burst;
    ps.arrived_ms = now;
    ps.req_seq = ++s.next_req_seq;
    ps.state = ST_WAITING;
    ps.poster = poster;
    dirty = true;
  }

  // A client that is going away hands its token back: the ledger entry is closed exactly as a returning client's
  // would be (update_return_time with zero overuse).  Returns true if the token was this slot's.
  bool give_back(int slot, double now) {
    bool held = s.holder == slot;
    if (held) {
      for (uint32_t i = s.ledger_len; i-- > 0;)
        if (s.ledger[i].slot == slot) {
          s.ledger[i].end = std::min(now, s.ledger[i].end);
          break;
        }
      PSlot& ps = s.slots[slot];
      if (ps.grants) ps.last_end = std::min(now, ps.last_end);
      ps.pod_quota = 0.0;  // the pod-level token is gone with it: the next request must be forwarded
      s.holder = -1;
      dirty = true;
    }
    if (s.slots[slot].state == ST_WAITING) {
      s.slots[slot].state = ST_IDLE;
      s.slots[slot].poster = -1;
      dirty = true;
    }
    return held;
  }
};

}  // namespace

struct gemhook_pool {
  Region* r = nullptr;
  int fd = -1;
  bool anonymous = false;
  int attach_idx = -1;     // this handle's client attachment (set by gemhook_pool_attach)
  int liveness_idx = -1;   // attachment whose OFD byte vouches for this handle's transactions (client or observer)
  std::atomic<uint32_t> rr{0};

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
    if (idx == liveness_idx || idx == attach_idx) return true;
    struct flock fl;
    memset(&fl, 0, sizeof(fl));
    fl.l_type = F_WRLCK;
    fl.l_whence = SEEK_SET;
    fl.l_start = (off_t)(offsetof(Region, attach) + (size_t)idx * sizeof(Attach));
    fl.l_len = 1;
    if (fcntl(fd, F_OFD_GETLK, &fl) != 0) return true;  // cannot tell: assume alive
    return fl.l_type != F_UNLCK;
  }
  // The OFD byte is taken BEFORE the entry is marked in use: an entry that reads in_use == 1 is therefore always
  // vouched for by a held lock (or its owner is dead), and a reaper can never mistake an entry that is still being
  // set up for a dead one -- it did, once: two live clients then shared one attachment and overwrote each other's burst.
  int take_attachment(int slot) {
    for (uint32_t i = 0; i < MAX_ATTACH; i++) {
      if (r->attach[i].in_use.load(std::memory_order_acquire) != 0) continue;
      if (attach_lock((int)i, true) != 0) continue;  // somebody else is setting this entry up (or a stale lock holder lives)
      uint32_t exp = 0;
      if (r->attach[i].in_use.compare_exchange_strong(exp, 3u)) {  // 3 = being initialised: not yet visible to reapers
        Attach& a = r->attach[i];
        a.slot.store(slot);
        a.bytes.store(0);
        a.burst_bits.store(0);
        a.pid = (uint32_t)getpid();
        a.in_use.store(1u, std::memory_order_release);
        return (int)i;
      }
      attach_lock((int)i, false);
    }
    return -1;
  }
  uint32_t owner_tag() const { return liveness_idx >= 0 ? (uint32_t)liveness_idx : NO_OWNER; }

  int64_t now_us() const { return (gh_now_ns() - r->h.start_ns.load(std::memory_order_relaxed)) / 1000; }
  double now_ms() const { return (double)now_us() / 1e3; }  // scheduler.cpp:107-109

  // ---- state blocks ---------------------------------------------------------------------------------------
  // copy the live part of a state block (header fields, nslots slots, ledger_len spans); the source may be
  // overwritten underneath us (then `cur` has moved and the caller discards the copy), so lengths are clamped
  // (word-wise relaxed atomic loads, not memcpy: this is the reader side of a sequence-validated copy -- the loads
  //  may race with a writer that has recycled the source block, which the C++ memory model only tolerates for atomics;
  //  ThreadSanitizer intercepts memcpy even inside a no_sanitize function)
  GH_NO_TSAN static void raw_copy(void* dst, const void* src, size_t bytes) {
    uint64_t* d = (uint64_t*)dst;
    const uint64_t* s = (const uint64_t*)src;
    for (size_t i = 0, n = bytes / 8; i < n; i++) d[i] = __atomic_load_n(s + i, __ATOMIC_RELAXED);
  }
  GH_NO_TSAN static void copy_state(State* dst, const State* src) {
    static_assert(offsetof(State, slots) % 8 == 0 && sizeof(PSlot) % 8 == 0 && sizeof(Span) % 8 == 0, "word copy");
    raw_copy(dst, src, offsetof(State, slots));
    uint32_t ns = dst->nslots, ll = dst->ledger_len;
    if (ns > GEMHOOK_MAX_SLOTS) ns = GEMHOOK_MAX_SLOTS;
    if (ll > LEDGER_CAP) ll = LEDGER_CAP;
    dst->nslots = ns;
    dst->ledger_len = ll;
    raw_copy(dst->slots, src->slots, sizeof(PSlot) * ns);
    raw_copy(dst->ledger, src->ledger, sizeof(Span) * ll);
  }

  // recycle claimed blocks that can no longer become current: the owner is dead, or its publication attempt is
  // resolved.  Never touches the current block.  Why this is safe:
  //  * dead owner: liveness is checked BEFORE `cur` is read; a dead process cannot publish afterwards, so if the
  //    block is not current now it never will be;
  //  * resolved attempt: the control word says "to be published as sequence P" and `cur` is already at sequence
  //    >= P on another block: that publication CAS has failed or will fail.  A live owner notices through its own
  //    control-word CAS (Txn::commit) that the block is gone and takes a fresh one.
  int recycle_blocks() {
    int freed = 0;
    for (uint32_t b = 0; b < NBLK; b++) {
      uint64_t w = r->bctl[b].load(std::memory_order_acquire);
      if (!w) continue;
      uint32_t owner = ctl_owner(w);
      bool dead = owner != NO_OWNER && owner < MAX_ATTACH && !attach_owner_alive((int)owner);
      uint64_t c = r->h.cur.load(std::memory_order_acquire);
      if ((c & 0xff) == b) continue;
      uint64_t pub = ctl_pub(w);
      bool resolved = pub != 0 && (c >> 8) >= pub;
      if ((dead || resolved) && r->bctl[b].compare_exchange_strong(w, 0, std::memory_order_acq_rel)) {
        freed++;
        r->h.recycled.fetch_add(1, std::memory_order_relaxed);
      }
    }
    return freed;
  }

  int claim_block() {
    uint32_t start = rr.fetch_add(1, std::memory_order_relaxed);
    for (;;) {
      for (uint32_t k = 0; k < NBLK; k++) {
        uint32_t b = (start + k) % NBLK;
        uint64_t exp = 0;
        if (r->bctl[b].load(std::memory_order_relaxed) == 0 &&
            r->bctl[b].compare_exchange_strong(exp, ctl_make(owner_tag(), 0), std::memory_order_acq_rel))
          return (int)b;
      }
      if (recycle_blocks() == 0) sched_yield();  // every block is in somebody's live transaction: they finish in microseconds
    }
  }
};

namespace {

// One transaction.  begin() hands out a private, validated copy of the current state; commit() publishes it with a
// single CAS on `cur` (false = lost the race: call begin() again).
struct Txn {
  gemhook_pool* p;
  int blk = -1;
  uint64_t seen = 0;
  State* s = nullptr;

  explicit Txn(gemhook_pool* pool) : p(pool) {}
  ~Txn() { abort(); }
  Txn(const Txn&) = delete;
  Txn& operator=(const Txn&) = delete;

  State& begin() {
    Region* r = p->r;
    if (blk < 0) {
      blk = p->claim_block();
      GH_FAULT(FP_AFTER_CLAIM);
    }
    s = &r->blocks[blk];
    for (;;) {
      seen = r->h.cur.load(std::memory_order_acquire);
      gemhook_pool::copy_state(s, &r->blocks[seen & 0xff]);
      std::atomic_thread_fence(std::memory_order_acquire);
      if (r->h.cur.load(std::memory_order_relaxed) == seen) return *s;
    }
  }

  bool commit() {
    Region* r = p->r;
    uint64_t seq = (seen >> 8) + 1;
    uint64_t mine = ctl_make(p->owner_tag(), 0), pubw = ctl_make(p->owner_tag(), seq);
    if (!r->bctl[blk].compare_exchange_strong(mine, pubw, std::memory_order_acq_rel)) {
      blk = -1;  // our block was recycled from under a stalled transaction: start over with a fresh one
      return false;
    }
    GH_FAULT(FP_BEFORE_CAS);
    uint64_t exp = seen;
    if (r->h.cur.compare_exchange_strong(exp, (seq << 8) | (uint64_t)blk, std::memory_order_acq_rel)) {
      GH_FAULT(FP_AFTER_CAS);
      uint32_t old = (uint32_t)(seen & 0xff);
      uint64_t w = r->bctl[old].load(std::memory_order_relaxed);
      if (w) r->bctl[old].compare_exchange_strong(w, 0, std::memory_order_acq_rel);  // the superseded block is free again
      r->h.commits.fetch_add(1, std::memory_order_relaxed);
      blk = -1;  // now the current block: not ours to touch any more
      return true;
    }
    r->h.conflicts.fetch_add(1, std::memory_order_relaxed);
    uint64_t back = pubw;
    if (!r->bctl[blk].compare_exchange_strong(back, ctl_make(p->owner_tag(), 0), std::memory_order_acq_rel)) blk = -1;
    return false;
  }

  void abort() {
    if (blk >= 0) {
      uint64_t mine = ctl_make(p->owner_tag(), 0);
      p->r->bctl[blk].compare_exchange_strong(mine, 0, std::memory_order_acq_rel);
      blk = -1;
    }
  }
};

// run `fn(Policy&)` as a transaction until it commits (or until it reports nothing to commit); returns fn's result
template <class F>
auto transact(gemhook_pool* p, F fn) -> decltype(fn(*(Policy*)nullptr)) {
  Txn t(p);
  for (;;) {
    Policy pol(t.begin(), p->r->h);
    auto rc = fn(pol);
    GH_FAULT(FP_MID_POLICY);
    if (!pol.dirty) return rc;
    if (t.commit()) return rc;
  }
}

// consistent read-only view: the same machinery, never committed
template <class F>
auto observe(const gemhook_pool* cp, F fn) -> decltype(fn(*(const State*)nullptr)) {
  gemhook_pool* p = const_cast<gemhook_pool*>(cp);
  Txn t(p);
  const State& s = t.begin();
  return fn(s);
}

// hot-path peeks at a few words of the current block without claiming anything (validated against `cur`)
GH_NO_TSAN uint64_t peek_mem_limit(const gemhook_pool* p, int slot) {
  Region* r = p->r;
  for (;;) {
    uint64_t c = r->h.cur.load(std::memory_order_acquire);
    uint64_t v = r->blocks[c & 0xff].slots[slot].mem_limit;
    std::atomic_thread_fence(std::memory_order_acquire);
    if (r->h.cur.load(std::memory_order_relaxed) == c) return v;
  }
}
struct MailView {
  uint32_t state;
  int32_t poster;
};
GH_NO_TSAN MailView peek_mailbox(const gemhook_pool* p, int slot) {
  Region* r = p->r;
  for (;;) {
    uint64_t c = r->h.cur.load(std::memory_order_acquire);
    const PSlot& ps = r->blocks[c & 0xff].slots[slot];
    MailView v{ps.state, ps.poster};
    std::atomic_thread_fence(std::memory_order_acquire);
    if (r->h.cur.load(std::memory_order_relaxed) == c) return v;
  }
}
GH_NO_TSAN int peek_others_waiting(const gemhook_pool* p, int slot, uint32_t* nslots_out) {
  Region* r = p->r;
  for (;;) {
    uint64_t c = r->h.cur.load(std::memory_order_acquire);
    const State& s = r->blocks[c & 0xff];
    uint32_t n = s.nslots <= GEMHOOK_MAX_SLOTS ? s.nslots : (uint32_t)GEMHOOK_MAX_SLOTS;
    int any = 0;
    for (uint32_t i = 0; i < n; i++)
      if ((int)i != slot && s.slots[i].state == ST_WAITING) any = 1;
    std::atomic_thread_fence(std::memory_order_acquire);
    if (r->h.cur.load(std::memory_order_relaxed) == c) {
      if (nslots_out) *nslots_out = n;
      return any;
    }
  }
}

void wake_slot(gemhook_pool* p, int slot) {
  if (slot < 0 || slot >= GEMHOOK_MAX_SLOTS) return;
  p->r->shared[slot].wake.fetch_add(1, std::memory_order_release);
  futex(&p->r->shared[slot].wake, FUTEX_WAKE, INT_MAX, nullptr);
}

void reset_dynamic_state(gemhook_pool* p) {
  // the file outlived a reboot (hostPath): its clock origin, token holder, ledger, attachments, byte counters and
  // block claims describe processes that no longer exist.  Keep the configuration rows, drop the dynamic state.
  Region* r = p->r;
  Header& h = r->h;
  uint64_t c = h.cur.load();
  State& s = r->blocks[c & 0xff];
  for (uint32_t i = 0; i < s.nslots && i < GEMHOOK_MAX_SLOTS; i++) {
    PSlot& ps = s.slots[i];
    ps.quota = h.base_quota;
    ps.burst = 0;
    ps.last_start = ps.last_end = ps.closed_ms = 0;
    ps.grants = 0;
    ps.state = ST_IDLE;
    ps.poster = -1;
    ps.pod_quota = 0;
    ps.pod_token_us = 0;
    ps.pod_overuse = 0;
    r->shared[i].mem_used.store(0);
    r->shared[i].gpu_ns.store(0);
    r->shared[i].launches.store(0);
  }
  s.ledger_len = 0;
  s.holder = -1;
  s.deadline_ms = 0;
  memset((void*)r->attach, 0, sizeof(r->attach));
  for (uint32_t b = 0; b < NBLK; b++) r->bctl[b].store(b == (c & 0xff) ? ctl_make(NO_OWNER, c >> 8) : 0);
  h.start_ns.store(gh_now_ns());
}

}  // namespace

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
    h.start_ns.store(start_ns ? start_ns : gh_now_ns());
    h.boot_id.store(boot_id_hash());
    State& s0 = p->r->blocks[0];
    s0.holder = -1;
    s0.nslots = 0;
    p->r->bctl[0].store(ctl_make(NO_OWNER, 1));
    h.cur.store((1ull << 8) | 0);
    h.ready.store(1, std::memory_order_release);
  } else {
    for (int i = 0; i < 20000 && !h.ready.load(std::memory_order_acquire); i++) usleep(100);
    if (h.magic != POOL_MAGIC || !h.ready.load() || h.version != POOL_VERSION) {
      gh_set_error("%s is not an initialised gemhook credit pool (layout version %u)", path, POOL_VERSION);
      gemhook_pool_close(p);
      return nullptr;
    }
    uint64_t mine = boot_id_hash();
    uint64_t seen = h.boot_id.load();
    if (seen != mine) {
      // one opener resets, the others wait for it (boot id 1 = reset in progress)
      if (seen != 1 && h.boot_id.compare_exchange_strong(seen, 1)) {
        reset_dynamic_state(p);
        h.boot_id.store(mine, std::memory_order_release);
      } else {
        for (int i = 0; i < 20000 && h.boot_id.load(std::memory_order_acquire) != mine; i++) usleep(100);
      }
    }
  }
  // a liveness byte for this handle's transactions (clients turn it into their own attachment in attach())
  if (p->fd >= 0) {
    p->liveness_idx = p->take_attachment(-1);
    if (p->liveness_idx < 0) {  // table full of entries whose owners are gone (crashed tools): clear them out, try again
      gemhook_pool_reap(p);
      p->liveness_idx = p->take_attachment(-1);
    }
  }
  return p;
}

GH_EXPORT void gemhook_pool_close(gemhook_pool* p) {
  if (!p) return;
  if (p->r) {
    if (p->liveness_idx >= 0) {
      if (p->attach_idx >= 0) gemhook_pool_detach(p);
      p->attach_lock(p->liveness_idx, false);
      p->r->attach[p->liveness_idx].in_use.store(0, std::memory_order_release);
    }
    munmap(p->r, sizeof(Region));
  }
  if (p->fd >= 0) close(p->fd);
  delete p;
}

// the part of the region the device should see: header + per-slot counters (page aligned)
void* gh_pool_region(gemhook_pool* p, size_t* bytes) {
  if (bytes) *bytes = (offsetof(Region, attach) + 4095) & ~(size_t)4095;
  return p ? (void*)p->r : nullptr;
}
// byte offset of slot 0's {mem_used, mem_limit} pair inside that part and the stride per slot (device-side mirror)
void gh_pool_shared_layout(size_t* offset, size_t* stride) {
  if (offset) *offset = offsetof(Region, shared);
  if (stride) *stride = sizeof(SlotShared);
}

// read_resource_config (scheduler.cpp:183-217): "N" then N rows "name c2 c3 mem"; a re-read replaces the
// client's ClientInfo, i.e. its adaptive quota restarts from the base quota; usage history is kept.  The reference
// never forgets a client; neither do we until the table is full -- then a slot whose pod has left the file, holds
// nothing and has nobody attached is given to the newcomer.  A row whose name does not fit is skipped (the reference
// overflows char[HOST_NAME_MAX] there), the others are still loaded.
static int load_config_stamped(gemhook_pool* p, const char* text, int swap_columns, uint64_t stamp);
GH_EXPORT int gemhook_pool_load_config(gemhook_pool* p, const char* text, int swap_columns) {
  return load_config_stamped(p, text, swap_columns, 0);
}
// stamp != 0: apply only if the pool's configuration does not carry this stamp yet -- of N clients that notice the same
// rewrite of the quota file together, exactly one reloads (a reload resets every client's adaptive quota, so a second,
// late one would be visible in the quota sequence).  Returns -2 when somebody else had already loaded this version.
static int load_config_stamped(gemhook_pool* p, const char* text, int swap_columns, uint64_t stamp) {
  if (!p || !text) return -1;
  const char* c = text;
  char* end = nullptr;
  long n = strtol(c, &end, 10);
  if (end == c || n < 0) {
    gh_set_error("quota file: bad client count");
    return -1;
  }
  c = end;
  struct Row {
    char name[64];
    double c2, c3;
    unsigned long long mem;
  };
  std::vector<Row> rows;
  long parsed = 0, skipped = 0;
  for (long i = 0; i < n; i++) {
    char name[256];
    Row rw;
    int used = 0;
    if (sscanf(c, " %255s %lf %lf %llu%n", name, &rw.c2, &rw.c3, &rw.mem, &used) != 4) break;
    c += used;
    parsed++;
    if (strlen(name) >= sizeof(rw.name)) {
      gh_set_error("quota file: pod name \"%.40s...\" is longer than %zu bytes; row skipped", name, sizeof(rw.name) - 1);
      skipped++;
      continue;
    }
    snprintf(rw.name, sizeof(rw.name), "%s", name);
    rows.push_back(rw);
  }
  bool full = false;
  uint64_t limits[GEMHOOK_MAX_SLOTS];
  uint32_t nlim = 0;
  bool reused[GEMHOOK_MAX_SLOTS];
  bool stale = false;
  int loaded = transact(p, [&](Policy& pol) {
    State& s = pol.s;
    int cnt = 0;
    full = false;
    stale = stamp != 0 && s.quota_stamp == stamp;
    if (stale) return 0;
    for (uint32_t i = 0; i < GEMHOOK_MAX_SLOTS; i++) reused[i] = false;
    for (uint32_t i = 0; i < s.nslots; i++) s.slots[i].listed = 0;
    // rows of clients we already know keep their slot; new names are placed afterwards
    std::vector<int> where(rows.size(), -1);
    for (size_t k = 0; k < rows.size(); k++)
      for (uint32_t i = 0; i < s.nslots; i++)
        if (!strncmp(s.slots[i].name, rows[k].name, sizeof(s.slots[i].name))) where[k] = (int)i;
    for (size_t k = 0; k < rows.size(); k++)
      if (where[k] >= 0) s.slots[where[k]].listed = 1;
    for (size_t k = 0; k < rows.size(); k++) {
      int idx = where[k];
      if (idx < 0) {
        for (size_t q = 0; q < k; q++)  // the same new name twice in one file: one slot
          if (where[q] >= 0 && !strcmp(rows[q].name, rows[k].name)) idx = where[q];
      }
      if (idx < 0 && s.nslots < GEMHOOK_MAX_SLOTS) {
        idx = (int)s.nslots++;
        memset((void*)&s.slots[idx], 0, sizeof(PSlot));
      } else if (idx < 0) {
        for (uint32_t i = 0; i < s.nslots && idx < 0; i++) {
          PSlot& ps = s.slots[i];
          if (ps.listed || ps.state != ST_IDLE || s.holder == (int)i) continue;
          if (p->r->shared[i].mem_used.load(std::memory_order_relaxed) != 0) continue;
          bool attached = false;
          for (uint32_t a = 0; a < MAX_ATTACH; a++)
            if (p->r->attach[a].in_use.load(std::memory_order_relaxed) && p->r->attach[a].slot.load(std::memory_order_relaxed) == (int)i)
              attached = true;
          if (attached) continue;
          idx = (int)i;  // the pod left the file and holds nothing: its slot goes to the newcomer
          uint32_t kk = 0;
          for (uint32_t e = 0; e < s.ledger_len; e++)
            if (s.ledger[e].slot != idx) s.ledger[kk++] = s.ledger[e];
          s.ledger_len = kk;
          memset((void*)&s.slots[idx], 0, sizeof(PSlot));
          reused[idx] = true;
        }
        if (idx < 0) {
          full = true;
          continue;
        }
      }
      where[k] = idx;
      PSlot& ps = s.slots[idx];
      if (!ps.name[0]) {
        snprintf(ps.name, sizeof(ps.name), "%s", rows[k].name);
        ps.poster = -1;
      }
      ps.min_frac = swap_columns ? rows[k].c3 : rows[k].c2;
      ps.max_frac = swap_columns ? rows[k].c2 : rows[k].c3;
      ps.mem_limit = rows[k].mem;
      ps.quota = pol.h.base_quota;
      ps.burst = 0.0;
      ps.configured = 1;
      ps.listed = 1;
      cnt++;
    }
    nlim = s.nslots;
    for (uint32_t i = 0; i < s.nslots; i++) limits[i] = s.slots[i].mem_limit;
    // a half-written file (fewer rows than announced) keeps the old stamp: it will be read again
    if (parsed == n && !full) s.quota_stamp = stamp;
    pol.dirty = true;
    return cnt;
  });
  if (stale) return -2;
  for (uint32_t i = 0; i < nlim; i++) {
    if (reused[i]) {
      p->r->shared[i].gpu_ns.store(0, std::memory_order_relaxed);
      p->r->shared[i].launches.store(0, std::memory_order_relaxed);
    }
    p->r->shared[i].mem_limit_mirror.store(limits[i], std::memory_order_relaxed);
  }
  if (full) gh_set_error("quota file lists more than %d clients and no departed client's slot is free", GEMHOOK_MAX_SLOTS);
  return (!full && parsed == n && loaded + skipped == n) ? (int)n : -1;
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
  int rc = load_config_stamped(p, text, swap_columns, stamp);
  free(text);
  if (rc == -2) {  // another client loaded this very version while we were reading the file
    p->r->h.quota_stamp.store(stamp, std::memory_order_release);
    return 0;
  }
  if (rc < 0) return -1;  // half-written file: keep the old stamp, try again at the next renewal
  p->r->h.quota_stamp.store(stamp, std::memory_order_release);
  return 1;
}

GH_EXPORT int gemhook_pool_find(const gemhook_pool* p, const char* name) {
  return observe(p, [&](const State& s) {
    for (uint32_t i = 0; i < s.nslots; i++)
      if (!strncmp(s.slots[i].name, name, sizeof(s.slots[i].name))) return (int)i;
    return -1;
  });
}
GH_EXPORT int gemhook_pool_nslots(const gemhook_pool* p) {
  uint32_t n = 0;
  peek_others_waiting(p, -1, &n);
  return (int)n;
}

GH_EXPORT int gemhook_pool_request(gemhook_pool* p, int slot, double now_ms, double overuse_ms, double burst_ms) {
  if (!p || slot < 0 || slot >= gemhook_pool_nslots(p)) return -1;
  transact(p, [&](Policy& pol) {
    pol.request(slot, now_ms, overuse_ms, burst_ms, -1);
    return 0;
  });
  return 0;
}

GH_EXPORT int gemhook_pool_schedule(gemhook_pool* p, double now_ms, int* slot_out, double* quota_out, double* sleep_ms_out) {
  int who = -1;
  int rc = transact(p, [&](Policy& pol) {
    who = -1;
    return pol.schedule(now_ms, &who, quota_out, sleep_ms_out);
  });
  if (rc == 1) {
    if (slot_out) *slot_out = who;
    wake_slot(p, who);
  }
  return rc;
}

GH_EXPORT double gemhook_pool_usage(gemhook_pool* p, int slot, double now_ms) {
  // on a private copy that is never published: asking for a number must not change the ledger
  Txn t(p);
  State& s = t.begin();
  double usage[GEMHOOK_MAX_SLOTS] = {0};
  double a, b;
  Policy pol(s, p->r->h);
  pol.window_usage(now_ms, usage, a, b);
  return (slot >= 0 && slot < GEMHOOK_MAX_SLOTS) ? usage[slot] : 0.0;
}

GH_EXPORT size_t gemhook_pool_history(const gemhook_pool* cp, int* slots, double* starts, double* ends, size_t cap) {
  return observe(cp, [&](const State& s) {
    size_t n = s.ledger_len;
    for (size_t i = 0; i < n && i < cap; i++) {
      if (slots) slots[i] = s.ledger[i].slot;
      if (starts) starts[i] = s.ledger[i].start;
      if (ends) ends[i] = s.ledger[i].end;
    }
    return n;
  });
}

GH_EXPORT double gemhook_pool_now_ms(const gemhook_pool* p) { return p ? p->now_ms() : 0.0; }

GH_EXPORT double gemhook_pool_accumulated_ms(const gemhook_pool* cp, int slot) {
  if (slot < 0 || slot >= GEMHOOK_MAX_SLOTS) return 0.0;
  return observe(cp, [&](const State& s) {
    const PSlot& ps = s.slots[slot];
    return ps.grants ? ps.closed_ms + (ps.last_end - ps.last_start) : 0.0;
  });
}

// ---- pod-level token (gem-pmgr hook_kernel_launch, pod-manager.cpp:316-473) -----------------------------------
// Processes of one pod share the pod's token: a request is answered locally with the REMAINING pod quota unless
// `elapsed + burst > pod_quota`, in which case it is forwarded to the scheduler with the pod's maximum overuse and
// the maximum burst over its processes.  Returns 1 = forward (fwd_* filled), 0 = answered (*remain_ms).
static int pod_launch(gemhook_pool* p, Policy& pol, int slot, int attach_idx, int64_t now_us, double overuse, double burst,
                      double* fwd_overuse, double* fwd_burst, double* remain) {
  PSlot& s = pol.s.slots[slot];
  double mo = std::max(overuse, s.pod_overuse);
  if (mo != s.pod_overuse) {
    s.pod_overuse = mo;
    pol.dirty = true;
  }
  if (attach_idx >= 0) p->r->attach[attach_idx].burst_bits.store(double_to_bits(burst), std::memory_order_relaxed);
  double elapsed = (double)(now_us - s.pod_token_us) / 1e3;
  if (elapsed + burst > s.pod_quota) {
    double mx = attach_idx >= 0 ? 0.0 : burst;
    for (uint32_t i = 0; i < MAX_ATTACH; i++)
      if (p->r->attach[i].in_use.load(std::memory_order_relaxed) && p->r->attach[i].slot.load(std::memory_order_relaxed) == slot)
        mx = std::max(bits_to_double(p->r->attach[i].burst_bits.load(std::memory_order_relaxed)), mx);
    if (fwd_overuse) *fwd_overuse = s.pod_overuse;
    if (fwd_burst) *fwd_burst = mx;
    return 1;
  }
  if (remain) *remain = s.pod_quota - elapsed;
  return 0;
}
static double pod_granted(Policy& pol, int slot, int64_t now_us, double quota) {
  PSlot& s = pol.s.slots[slot];
  s.pod_quota = quota;
  s.pod_token_us = now_us;
  s.pod_overuse = 0.0;
  pol.dirty = true;
  return s.pod_quota - 0.0;
}
GH_EXPORT int gemhook_pool_pod_launch(gemhook_pool* p, int slot, int64_t now_us, double overuse_ms, double burst_ms,
                                      double* fwd_overuse_ms, double* fwd_burst_ms, double* remain_ms) {
  return transact(p, [&](Policy& pol) {
    return pod_launch(p, pol, slot, p->attach_idx, now_us, overuse_ms, burst_ms, fwd_overuse_ms, fwd_burst_ms, remain_ms);
  });
}
GH_EXPORT double gemhook_pool_pod_granted(gemhook_pool* p, int slot, int64_t now_us, double quota_ms) {
  return transact(p, [&](Policy& pol) { return pod_granted(pol, slot, now_us, quota_ms); });
}

// ---- live acquisition ---------------------------------------------------------------------------------------------
GH_EXPORT double gemhook_pool_acquire(gemhook_pool* p, int slot, double overuse_ms, double burst_ms) {
  return gemhook_pool_acquire_ex(p, slot, overuse_ms, burst_ms, nullptr);
}

// Post the request, then arbitrate / wait until OUR request is granted.  Processes of one pod share the slot's
// mailbox: only one request per pod is in flight (what gem-pmgr's sleeping_count hand-shake serialises,
// pod-manager.cpp:316-473); a sibling that arrives meanwhile waits for that grant and is then normally answered from
// the fresh pod token.
GH_EXPORT double gemhook_pool_acquire_ex(gemhook_pool* p, int slot, double overuse_ms, double burst_ms, int* forwarded) {
  if (forwarded) *forwarded = 1;
  // identity of a request in flight: the poster's attachment; an unattached handle (gem-arbiter serving TCP clients,
  // tests) tags its requests with -2 - slot so that a sibling thread is still told apart from "nobody"
  const int me = p->attach_idx >= 0 ? p->attach_idx : -2;
  SlotShared& sh = p->r->shared[slot];
  enum { ANSWERED, POSTED, BUSY_SIBLING };
  int idle_rounds = 0;
  for (;;) {
    // ---- phase 1: pod rule, post, first decision -- one transaction
    double remain = 0.0;
    int granted_to = -1;
    int st = transact(p, [&](Policy& pol) {
      granted_to = -1;
      PSlot& ps = pol.s.slots[slot];
      if (ps.state != ST_IDLE) return (int)BUSY_SIBLING;  // a request of this pod is already in flight
      double fo = overuse_ms, fb = burst_ms;
      if (!pod_launch(p, pol, slot, p->attach_idx, p->now_us(), overuse_ms, burst_ms, &fo, &fb, &remain)) return (int)ANSWERED;
      double now = p->now_ms();
      pol.request(slot, now, fo, fb, me);
      pol.schedule(now, &granted_to, nullptr, nullptr);
      return (int)POSTED;
    });
    if (granted_to >= 0 && granted_to != slot) wake_slot(p, granted_to);
    if (st == ANSWERED) {
      if (forwarded) *forwarded = 0;
      return remain;  // the pod's token still covers this burst (pod-manager.cpp:472)
    }
    const bool mine = st == POSTED;
    // ---- phase 2: wait for the grant of the request in flight
    for (;;) {
      uint32_t w0 = sh.wake.load(std::memory_order_acquire);
      MailView v = peek_mailbox(p, slot);
      if (v.state == ST_GRANTED && mine && v.poster == me) {
        bool took = false;
        double got = transact(p, [&](Policy& pol) {
          PSlot& ps = pol.s.slots[slot];
          took = false;
          if (ps.state != ST_GRANTED || ps.poster != me) return 0.0;  // reaped meanwhile
          took = true;
          double q = ps.granted_quota;
          ps.state = ST_IDLE;
          ps.poster = -1;
          return pod_granted(pol, slot, p->now_us(), q);
        });
        if (took) {
          wake_slot(p, slot);  // siblings waiting for this grant re-evaluate the pod rule
          return got;
        }
        break;  // start over
      }
      if (v.state == ST_IDLE) break;  // the request in flight is gone (consumed by its poster, or reaped): start over
      if (!mine && v.poster >= 0 && v.poster != p->attach_idx && !p->attach_owner_alive(v.poster)) gemhook_pool_reap(p);
      // drive the policy: whoever waits runs the decisions (there is no daemon)
      int who = -1;
      double sleep_ms = 0;
      int rc = transact(p, [&](Policy& pol) {
        who = -1;
        return pol.schedule(p->now_ms(), &who, nullptr, &sleep_ms);
      });
      if (rc == 1) {
        wake_slot(p, who);
        if (who == slot) continue;
      }
      if (++idle_rounds % 8 == 0) gemhook_pool_reap(p);  // a dead holder must not stall everybody until its deadline
      // someone else holds the token, or everyone is throttled: wait on our slot's wake word until the hint expires
      // or a granter wakes us.  Short waits spin (no context switch on a quick hand-over).
      double wait_ms = (rc == 0 || rc == -2) ? sleep_ms : 0.2;
      if (wait_ms < 0.05) {
        for (int i = 0; i < 200 && sh.wake.load(std::memory_order_acquire) == w0; i++) __builtin_ia32_pause();
        continue;
      }
      if (wait_ms > 50.0) wait_ms = 50.0;  // re-evaluate periodically (config reloads, dead holders)
      struct timespec ts;
      ts.tv_sec = (time_t)(wait_ms / 1e3);
      ts.tv_nsec = (long)((wait_ms - ts.tv_sec * 1e3) * 1e6);
      futex(&sh.wake, FUTEX_WAIT, w0, &ts);
    }
  }
}

// ---- attachments ------------------------------------------------------------------------------------------
GH_EXPORT int gemhook_pool_attach(gemhook_pool* p, int slot) {
  if (!p || slot < 0) return -1;
  if (p->attach_idx >= 0) return p->attach_idx;
  int idx;
  if (p->liveness_idx >= 0) {  // the handle's observer entry becomes the client attachment: its OFD byte is already held
    idx = p->liveness_idx;
    p->r->attach[idx].bytes.store(0);
    p->r->attach[idx].burst_bits.store(0);
    p->r->attach[idx].slot.store(slot, std::memory_order_release);
  } else {
    idx = p->take_attachment(slot);
    if (idx < 0) {
      gh_set_error("credit pool: more than %u attached processes", MAX_ATTACH);
      return -1;
    }
    p->liveness_idx = idx;
  }
  p->attach_idx = idx;
  return idx;
}

GH_EXPORT void gemhook_pool_detach(gemhook_pool* p) {
  if (!p || p->attach_idx < 0) return;
  Attach& a = p->r->attach[p->attach_idx];
  uint64_t left = a.bytes.exchange(0);
  int slot = a.slot.load();
  if (left && slot >= 0) p->r->shared[slot].mem_used.fetch_sub(left, std::memory_order_acq_rel);  // exit without freeing
  a.burst_bits.store(0);
  a.slot.store(-1, std::memory_order_release);  // back to a plain observer entry: the handle may still run transactions
  p->attach_idx = -1;
}

static bool slot_has_other_live_attachment(gemhook_pool* p, int slot, int except) {
  for (uint32_t j = 0; j < MAX_ATTACH; j++) {
    if ((int)j == except) continue;
    Attach& o = p->r->attach[j];
    if (o.in_use.load(std::memory_order_acquire) == 1 && o.slot.load(std::memory_order_relaxed) == slot && p->attach_owner_alive((int)j)) return true;
  }
  return false;
}

// reclaim what dead processes left behind: their bytes, the request they posted, the token if their pod held it and
// nobody of the pod is left, and the state blocks they had claimed
GH_EXPORT int gemhook_pool_reap(gemhook_pool* p) {
  if (!p) return 0;
  int reaped = 0;
  for (uint32_t i = 0; i < MAX_ATTACH; i++) {
    Attach& a = p->r->attach[i];
    if (a.in_use.load(std::memory_order_acquire) != 1 || (int)i == p->attach_idx || (int)i == p->liveness_idx) continue;
    if (p->attach_owner_alive((int)i)) continue;
    uint32_t one = 1;
    if (!a.in_use.compare_exchange_strong(one, 2u)) continue;  // somebody else is reaping this entry
    uint64_t left = a.bytes.exchange(0);
    int slot = a.slot.load();
    if (slot >= 0 && slot < GEMHOOK_MAX_SLOTS) {
      if (left) p->r->shared[slot].mem_used.fetch_sub(left, std::memory_order_acq_rel);
      bool others = slot_has_other_live_attachment(p, slot, (int)i);
      int woken = -1;
      transact(p, [&](Policy& pol) {
        PSlot& ps = pol.s.slots[slot];
        woken = -1;
        if (ps.state != ST_IDLE && ps.poster == (int)i) {  // its request / unconsumed grant dies with it
          ps.state = ST_IDLE;
          ps.poster = -1;
          pol.dirty = true;
        }
        if (!others && pol.s.holder == slot) {
          pol.s.holder = -1;
          ps.pod_quota = 0.0;
          if (ps.state != ST_IDLE) {
            ps.state = ST_IDLE;
            ps.poster = -1;
          }
          pol.dirty = true;
          pol.schedule(p->now_ms(), &woken, nullptr, nullptr);
        }
        return 0;
      });
      wake_slot(p, slot);
      if (woken >= 0) wake_slot(p, woken);
      reaped++;  // (observer entries of dead tools are recycled without being counted)
    }
    a.in_use.store(0, std::memory_order_release);
  }
  p->recycle_blocks();
  return reaped;
}

// A client that is going away (process exit) hands its token back instead of letting the scheduler wait
// for the quota to time out (the reference can only time out: scheduler.cpp:507-510).  While another live process of
// the same pod is attached the pod keeps its token (they share it, pod-manager.cpp:316-473): nothing is released.
GH_EXPORT void gemhook_pool_release(gemhook_pool* p, int slot) {
  if (!p || slot < 0 || slot >= GEMHOOK_MAX_SLOTS) return;
  if (p->attach_idx >= 0 && slot_has_other_live_attachment(p, slot, p->attach_idx)) return;
  int who = -1;
  transact(p, [&](Policy& pol) {
    who = -1;
    double now = p->now_ms();
    if (pol.give_back(slot, now)) pol.schedule(now, &who, nullptr, nullptr);
    return 0;
  });
  if (who >= 0) wake_slot(p, who);
  wake_slot(p, slot);
}

// The outstanding token is declared timed out (what gem-schd concludes when its timedwait on the holder
// returns ETIMEDOUT, scheduler.cpp:507-510) without touching the ledger: trace replays use it to decouple
// decisions from wall time, a node agent can use it to revoke the token of a client it knows is dead.
GH_EXPORT void gemhook_pool_expire_token(gemhook_pool* p) {
  if (!p) return;
  transact(p, [&](Policy& pol) {
    if (pol.s.holder != -1) {
      pol.s.holder = -1;
      pol.dirty = true;
    }
    return 0;
  });
}

// 1 if some OTHER client is waiting for the token right now (a peek, used by the yield-on-idle option)
GH_EXPORT int gemhook_pool_others_waiting(const gemhook_pool* p, int slot) { return peek_others_waiting(p, slot, nullptr); }

// ---- gpu_mem cap: integer exact, requested bytes (hook.cpp:590-617, pod-manager.cpp:295-313) -----------
GH_EXPORT int gemhook_pool_mem_reserve(gemhook_pool* p, int slot, uint64_t bytes) {
  SlotShared& sh = p->r->shared[slot];
  const uint64_t limit = peek_mem_limit(p, slot);
  uint64_t used = sh.mem_used.load(std::memory_order_relaxed);
  for (;;) {
    // reference pre-hook: remain = limit - used; deny iff bytes > remain (hook.cpp:593-598).  A quota-file reload may
    // have lowered the limit below what is already in use: then nothing more fits (gem-pmgr: used + bytes > limit,
    // pod-manager.cpp:299) -- the unsigned subtraction must not wrap into "plenty of room".
    if (used > limit || bytes > limit - used) return 0;
    if (sh.mem_used.compare_exchange_weak(used, used + bytes, std::memory_order_acq_rel)) {
      if (p->attach_idx >= 0 && p->r->attach[p->attach_idx].slot.load(std::memory_order_relaxed) == slot)
        p->r->attach[p->attach_idx].bytes.fetch_add(bytes, std::memory_order_relaxed);
      return 1;
    }
  }
}
GH_EXPORT void gemhook_pool_mem_release(gemhook_pool* p, int slot, uint64_t bytes) {
  p->r->shared[slot].mem_used.fetch_sub(bytes, std::memory_order_acq_rel);
  if (p->attach_idx >= 0 && p->r->attach[p->attach_idx].slot.load(std::memory_order_relaxed) == slot)
    p->r->attach[p->attach_idx].bytes.fetch_sub(bytes, std::memory_order_relaxed);
}
GH_EXPORT void gemhook_pool_mem_info(const gemhook_pool* p, int slot, uint64_t* used, uint64_t* limit) {
  if (used) *used = p->r->shared[slot].mem_used.load(std::memory_order_acquire);
  if (limit) *limit = peek_mem_limit(p, slot);
}
GH_EXPORT int gemhook_pool_slot_info(const gemhook_pool* cp, int slot, gemhook_slot_info* out) {
  if (!cp || !out || slot < 0 || slot >= GEMHOOK_MAX_SLOTS) return -1;
  const SlotShared& sh = cp->r->shared[slot];
  return observe(cp, [&](const State& s) {
    if (slot >= (int)s.nslots) return -1;
    const PSlot& ps = s.slots[slot];
    memset(out, 0, sizeof(*out));
    snprintf(out->name, sizeof(out->name), "%s", ps.name);
    out->min_frac = ps.min_frac;
    out->max_frac = ps.max_frac;
    out->mem_limit = ps.mem_limit;
    out->mem_used = sh.mem_used.load(std::memory_order_relaxed);
    out->gpu_ns = sh.gpu_ns.load(std::memory_order_relaxed);
    out->launches = sh.launches.load(std::memory_order_relaxed);
    out->tokens = ps.grants;
    out->quota_ms = ps.quota;
    out->accumulated_ms = ps.grants ? ps.closed_ms + (ps.last_end - ps.last_start) : 0.0;
    out->holds_token = s.holder == slot ? 1 : 0;
    out->waiting = ps.state == ST_WAITING ? 1 : 0;
    return 0;
  });
}

// transaction statistics of the pool: commits, lost publication races, recycled (leaked and recovered) state blocks
GH_EXPORT void gemhook_pool_counters(const gemhook_pool* p, uint64_t* commits, uint64_t* conflicts, uint64_t* recycled) {
  if (commits) *commits = p->r->h.commits.load(std::memory_order_relaxed);
  if (conflicts) *conflicts = p->r->h.conflicts.load(std::memory_order_relaxed);
  if (recycled) *recycled = p->r->h.recycled.load(std::memory_order_relaxed);
}

GH_EXPORT const void* gemhook_pool_shared_words(const gemhook_pool* p, int slot) {
  return (p && slot >= 0 && slot < GEMHOOK_MAX_SLOTS) ? (const void*)&p->r->shared[slot] : nullptr;
}

void gh_pool_add_usage(gemhook_pool* p, int slot, uint64_t gpu_ns, uint64_t launches) {
  p->r->shared[slot].gpu_ns.fetch_add(gpu_ns, std::memory_order_relaxed);
  p->r->shared[slot].launches.fetch_add(launches, std::memory_order_relaxed);
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
