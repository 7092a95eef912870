// gh_hook.cpp -- the live hook: process singleton, launch slow path, token renewal, overuse tracker,
// segment accounting.  The per-launch FAST path is in gh_interpose.cpp: one load of gh_gate_fast and one plain
// increment of the thread's own counter -- no mutex, no lock prefix, no clock read, no syscall (the reference takes three mutex pairs per
// launch: window.record_stop, expiration_status_mutex, burst.record_start; hook.cpp:515-555).
//
// Reference behaviour kept (hook.cpp): first-use initialisation incl. a discarded first token (:724-771);
// renewal only at a burst edge when `held + predicted_burst >= quota` (:520-521); the GPU is drained and
// overuse measured with ONE event pair per token before a new token is requested (:456-502, 527-538);
// sync calls end the burst and start a window (:334-340).
//
// New: (a) every burst (and optionally every K launches inside it) is bracketed by CUDA events on the
// launching stream; the (slot, launches, elapsed_ns) records are reduced ON THE DEVICE by the sm_100a
// kernel on a dedicated non-blocking accounting stream and published to a mapped pinned page;
// (b) gpu_mem is enforced against the shared pool counter; (c) tokens come from the shared credit pool
// when GEMHOOK_POOL is set (TCP to the unmodified gem-pmgr/gem-schd otherwise).
#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <atomic>

#include "gh_internal.h"

int gh_rpc(gemhook_request* req, gemhook_response* rsp);
int gh_acct_push_async(gemhook_acct* a, const gemhook_record* pinned_records, size_t n);
uint64_t gh_acct_reduce_launches(const gemhook_acct* a);
void gh_pool_add_usage(gemhook_pool* p, int slot, uint64_t gpu_ns, uint64_t launches);
void* gh_pool_region(gemhook_pool* p, size_t* bytes);

void gh_mem_local(uint64_t* free_b, uint64_t* total_b);  // like gh_mem_info, but never an RPC (gh_mem.cpp)

uint32_t gh_gate_open = 0;  // accessed with relaxed __atomic builtins only (plain MOVs on x86, race-free by the book)
gh_hot_line gh_hot = {0u, 0xffffffffu, 0, {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}};
__thread gh_tl_state gh_tl __attribute__((tls_model("initial-exec"), aligned(16))) = {0, 0};

// registry of the live threads' counters + the total of the threads that are gone
static pthread_mutex_t g_thr_mu = PTHREAD_MUTEX_INITIALIZER;
static uint64_t** g_thr_addr = nullptr;
static size_t g_thr_n = 0, g_thr_cap = 0;
static uint64_t g_thr_dead_total = 0;
static pthread_key_t g_thr_key;
static pthread_once_t g_thr_key_once = PTHREAD_ONCE_INIT;

static void thread_gone(void* p) {  // pthread_key destructor: fold the leaving thread's count into the total
  uint64_t* addr = (uint64_t*)p;
  pthread_mutex_lock(&g_thr_mu);
  for (size_t i = 0; i < g_thr_n; i++)
    if (g_thr_addr[i] == addr) {
      g_thr_dead_total += __atomic_load_n(addr, __ATOMIC_RELAXED);
      g_thr_addr[i] = g_thr_addr[--g_thr_n];
      break;
    }
  pthread_mutex_unlock(&g_thr_mu);
}
static void thread_key_init(void) {
  pthread_key_create(&g_thr_key, thread_gone);
  // initial-exec TLS: the variable sits at the same offset from the thread pointer in every thread
  gh_hot.tls_off = (int64_t)((char*)&gh_tl - (char*)__builtin_thread_pointer());
}
void gh_thread_register(void) {
  if (gh_tl.registered) return;
  pthread_once(&g_thr_key_once, thread_key_init);
  pthread_mutex_lock(&g_thr_mu);
  if (g_thr_n == g_thr_cap) {
    size_t nc = g_thr_cap ? g_thr_cap * 2 : 16;
    uint64_t** na = (uint64_t**)realloc(g_thr_addr, nc * sizeof(uint64_t*));
    if (na) {
      g_thr_addr = na;
      g_thr_cap = nc;
    }
  }
  if (g_thr_n < g_thr_cap) {
    g_thr_addr[g_thr_n++] = &gh_tl.count;
    pthread_setspecific(g_thr_key, &gh_tl.count);
    gh_tl.registered = 1;
  }
  pthread_mutex_unlock(&g_thr_mu);
}
uint64_t gh_total_launches(void) {
  pthread_mutex_lock(&g_thr_mu);
  uint64_t n = g_thr_dead_total;
  for (size_t i = 0; i < g_thr_n; i++) n += __atomic_load_n(g_thr_addr[i], __ATOMIC_RELAXED);
  pthread_mutex_unlock(&g_thr_mu);
  return n;
}
void gh_gate_set(uint32_t open) {
  __atomic_store_n(&gh_gate_open, open, __ATOMIC_RELAXED);
  // with CU_HOOK_DEBUG every launch goes through the counting slow wrapper, so the fast word stays 0
  __atomic_store_n(&gh_gate_fast, (open && !__atomic_load_n(&gh_hook_debug, __ATOMIC_RELAXED)) ? 1u : 0u, __ATOMIC_RELAXED);
}

namespace {
const int SEG_EVENTS = 64;
}

struct gh_live {
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;  // guards gate + segments (expiration_status_mutex's role)
  pthread_cond_t renew_cv = PTHREAD_COND_INITIALIZER;
  bool renewing = false;
  int64_t last_sync_ns = 0;     // when the last synchronising call returned
  double last_window_ms = 0.0;  // idle gap between that return and the next launch (most recent one)
  bool yielded = false;  // the running token was handed back at a sync (GEMHOOK_YIELD_ON_IDLE): its drain reports no overuse
  gemhook_gate* gate = nullptr;
  bool enabled = false;
  std::atomic<bool> cuda_ready{false};

  // transport
  gemhook_pool* pool = nullptr;
  int slot = 0;

  // CUDA objects
  CUcontext ctx = nullptr;
  CUevent ev_token = nullptr, ev_drain = nullptr;
  gemhook_acct* acct = nullptr;

  // overuse tracker (hook.cpp:456-502)
  pthread_t trk_tid;
  pthread_mutex_t trk_mu = PTHREAD_MUTEX_INITIALIZER;
  pthread_cond_t trk_start_cv = PTHREAD_COND_INITIALIZER, trk_done_cv = PTHREAD_COND_INITIALIZER;
  pthread_cond_t trk_intr_cv;  // CLOCK_MONOTONIC
  bool trk_armed = false, trk_done = true, trk_intr = false;
  int64_t trk_deadline_ns = 0;

  // segments: events bracketing launch runs on the launching stream
  CUevent seg_ev[SEG_EVENTS];
  int seg_head = 0;      // index of the event that opened the current segment (-1: none open)
  bool seg_open = false;
  CUstream seg_stream = nullptr;
  uint64_t seg_first_launch = 0;
  struct Pending {
    int ev_begin, ev_end;
    uint32_t launches;
    uint64_t idle_ns;  // host-measured idle gaps inside a merged segment, subtracted from the event span
  } pending[SEG_EVENTS];
  // segment merging: a segment is closed at a sync only once it is at least seg_min_ns old, so sync-heavy
  // applications do not pay two event records per tiny burst; the idle gaps it then spans are measured on
  // the host clock (sync return -> next launch) and subtracted
  int64_t seg_begin_host_ns = 0, seg_sync_return_ns = 0;
  uint64_t seg_idle_ns = 0;
  bool seg_spans_sync = false;
  int npending = 0;
  bool seg_end_recorded = false;

  // staging of records (pinned, double buffered)
  gemhook_record* stage[2] = {nullptr, nullptr};
  CUevent stage_done[2] = {nullptr, nullptr};  // recorded on the accounting stream after each buffer's copy
  bool stage_inflight[2] = {false, false};
  uint32_t stage_cap = 0, stage_n = 0;
  int stage_cur = 0;

  // stats
  std::atomic<uint64_t> slow_path{0}, token_requests{0}, host_syncs{0}, segments{0}, yields{0};
  uint64_t gpu_ns_host = 0;  // host-side running sum of the same records (cross-check of the device totals)
  std::atomic<uint64_t> token_wait_ns{0};
  double accumulated_token_ms = 0;
  int64_t last_token_ns = 0;
  int64_t token_returned_ns = 0;  // set when the token was handed back early (yield): the ledger entry ended there
  double last_quota_ms = 0;
  bool token_ev_valid = false;  // ev_token was recorded for the running token (not while a stream capture was active)

  // live publication of the device-reduced usage into the pool slot (what exporters scrape)
  uint64_t pub_epoch = 0, pub_ns = 0, pub_launches = 0;

  // GEMHOOK_TOKEN_TRACE: (t, overuse, burst) -> quota of every token request, kept in memory, written at exit
  struct TokenTrace {
    double t_ms, overuse_ms, burst_ms, quota_ms;
    int forwarded;
  };
  TokenTrace* trace = nullptr;
  uint32_t trace_n = 0, trace_cap = 0;
};

static gh_live* g_live = nullptr;
static pthread_once_t live_once = PTHREAD_ONCE_INIT;

static void fatal_or_disable(gh_live* L, const char* what) {
  gh_set_error("%s", what);
  fprintf(stderr, "[gemhook %d] %s\n", (int)getpid(), what);
  if (gh_cfg.exit_on_failure) exit(1);  // reference: exit() of the application (hook.cpp:236, 285, 360, 388, 439)
  L->enabled = false;
}

// A stream that is being captured into a CUDA graph must not see our events: they would become event-record
// nodes of the application's graph, be re-recorded at every replay, and an elapsed-time query on them fails the
// capture.  The legacy stream cannot be captured itself, but touching it while a BLOCKING stream captures is an
// implicit dependency that invalidates the application's capture -- cuStreamIsCapturing(legacy) reports exactly
// that case with an error and does not invalidate anything (driver API docs).  The reference never looks
// (hook.cpp:482-485, 543 record on the legacy stream unconditionally).
// On the application's OWN launch path the legacy stream needs no check: a launch on the legacy stream while a blocking
// stream captures has already invalidated that capture by itself (launch_path = true skips the driver call there).
static bool stream_capturing(CUstream s, bool launch_path = false) {
  if (!gh_real.cuStreamIsCapturing || (launch_path && !s)) return false;
  CUstreamCaptureStatus st = CU_STREAM_CAPTURE_STATUS_NONE;
  CUresult r = GH_CALL(cuStreamIsCapturing, s, &st);
  return r != CUDA_SUCCESS || st != CU_STREAM_CAPTURE_STATUS_NONE;
}

// device totals -> pool slot, as deltas (several processes of one pod add into the same slot).  The totals page is
// read without any CUDA call; it reflects every reduce kernel that has completed.  Caller holds L->mu.
static void publish_usage_locked(gh_live* L, bool force = false) {
  if (!L->acct || !L->pool) return;
  // every reduce launch publishes exactly one epoch: when the last epoch seen equals the number of reduce launches issued,
  // nothing can have changed (one relaxed load instead of a walk over the page at every synchronising call)
  if (!force && L->pub_epoch == gh_acct_reduce_launches(L->acct)) return;
  uint64_t tot[GEMHOOK_MAX_SLOTS * 3], ep = 0;
  if (gemhook_acct_read_totals(L->acct, tot, &ep) != 0 || ep == L->pub_epoch) return;  // (kernel still running: look again)
  uint64_t ns = tot[L->slot * 3], la = tot[L->slot * 3 + 1];
  gh_pool_add_usage(L->pool, L->slot, ns - L->pub_ns, la - L->pub_launches);
  L->pub_epoch = ep;
  L->pub_ns = ns;
  L->pub_launches = la;
}

// ---- token transport ----------------------------------------------------------------------------------
static double token_from_scheduler(gh_live* L, double overuse_ms, double next_burst_ms) {
  L->token_requests.fetch_add(1, std::memory_order_relaxed);
  int64_t t0 = gh_now_ns();
  double q = 0.0;
  int forwarded = -1;  // unknown over TCP: gem-pmgr decides (pod-manager.cpp:316-473)
  if (gh_cfg.transport == 1) {
    q = gemhook_pool_acquire_ex(L->pool, L->slot, overuse_ms, next_burst_ms, &forwarded);
  } else {
    gemhook_request req;
    gemhook_response rsp;
    memset(&req, 0, sizeof(req));
    req.type = GEMHOOK_REQ_QUOTA;
    req.overuse_ms = overuse_ms;
    req.burst_ms = next_burst_ms;
    if (gh_rpc(&req, &rsp) != 0) {
      fatal_or_disable(L, "failed to get a token from the scheduler");
      return 1e12;
    }
    q = rsp.quota_ms;
  }
  L->token_wait_ns.fetch_add((uint64_t)(gh_now_ns() - t0), std::memory_order_relaxed);
  GH_DEBUG("token: overuse %.3f ms, next burst %.3f ms -> quota %.3f ms", overuse_ms, next_burst_ms, q);
  if (L->trace && L->trace_n < L->trace_cap)
    L->trace[L->trace_n++] = {(double)t0 / 1e6, overuse_ms, next_burst_ms, q, forwarded};
  return q;
}

// ---- overuse tracker ----------------------------------------------------------------------------------
static void host_sync_locked(gh_live* L, int64_t now);
static void gate_edge_locked(gh_live* L, int64_t now);
static void resolve_pending_locked(gh_live* L, bool may_block, int skip_last = 0);
static void sync_pre(bool force);

static void* tracker_main(void* arg) {
  gh_live* L = (gh_live*)arg;
  GH_CALL(cuCtxSetCurrent, L->ctx);
  if (gh_real.cuThreadExchangeStreamCaptureMode) {
    // this thread's driver calls must never be judged against (or invalidate) a capture the application runs in
    // global mode on another thread
    CUstreamCaptureMode m = CU_STREAM_CAPTURE_MODE_RELAXED;
    GH_CALL(cuThreadExchangeStreamCaptureMode, &m);
  }
  for (;;) {
    pthread_mutex_lock(&L->trk_mu);
    while (!L->trk_armed) pthread_cond_wait(&L->trk_start_cv, &L->trk_mu);
    L->trk_armed = false;
    // sleep until the token expires or a renewal wants the result earlier (hook.cpp:466-479)
    struct timespec ts;
    ts.tv_sec = L->trk_deadline_ns / 1000000000LL;
    ts.tv_nsec = L->trk_deadline_ns % 1000000000LL;
    while (!L->trk_intr) {
      if (pthread_cond_timedwait(&L->trk_intr_cv, &L->trk_mu, &ts) == ETIMEDOUT) break;
    }
    L->trk_intr = false;
    pthread_mutex_unlock(&L->trk_mu);

    // drain everything issued so far: an event on the legacy default stream waits for all blocking
    // streams (hook.cpp:449-453, 482-485); the events are reused, not leaked per token.
    float elapsed_ms = 0.f;
    bool measured = false;
    if (!gh_cfg.dry_run) {
      // (The accounting segment is NOT closed from this thread: an end marker recorded here can overtake a launch
      //  that already passed the gate but has not reached the driver yet -- that launch would then belong to no
      //  segment.  Measured on the box: one 1.3 ms graph replay lost per token.  The launching thread closes the
      //  segment itself, in program order, when it comes for the new token: gh_launch_slow.)
      // while the application captures a graph the legacy stream is off limits: fall back to the host clock
      if (L->token_ev_valid && !stream_capturing((CUstream)0) &&
          GH_CALL(cuEventRecord, L->ev_drain, (CUstream)0) == CUDA_SUCCESS &&
          GH_CALL(cuEventSynchronize, L->ev_drain) == CUDA_SUCCESS &&
          GH_CALL(cuEventElapsedTime, &elapsed_ms, L->ev_token, L->ev_drain) == CUDA_SUCCESS)
        measured = true;
    }
    pthread_mutex_lock(&L->mu);
    int64_t now = gh_now_ns();
    gate_edge_locked(L, now);  // burst ends here; next launch re-evaluates the token (the segment stays as it is)
    if (L->yielded) elapsed_ms = 0.f;  // token given back early: nothing was overused
    else if (!measured) elapsed_ms = (float)((double)(now - L->last_token_ns) / 1e6);
    L->yielded = false;
    gemhook_gate_tracker_fire(L->gate, now, elapsed_ms);
    if (L->cuda_ready && !gh_cfg.dry_run) resolve_pending_locked(L, false);
    pthread_mutex_unlock(&L->mu);

    pthread_mutex_lock(&L->trk_mu);
    L->trk_done = true;
    pthread_cond_broadcast(&L->trk_done_cv);
    pthread_mutex_unlock(&L->trk_mu);
  }
  return nullptr;
}

static void wait_tracker(gh_live* L) {
  pthread_mutex_lock(&L->trk_mu);
  if (!L->trk_done) {
    L->trk_intr = true;  // ask for the drain now (hook.cpp:527-533)
    pthread_cond_signal(&L->trk_intr_cv);
    while (!L->trk_done) pthread_cond_wait(&L->trk_done_cv, &L->trk_mu);
  }
  pthread_mutex_unlock(&L->trk_mu);
}

// ---- segments -----------------------------------------------------------------------------------------
static void stage_record(gh_live* L, uint32_t launches, uint64_t ns) {
  if (!L->stage[0]) return;
  if (L->stage_n == L->stage_cap) return;  // flushed on the sync path; cannot overflow in practice
  gemhook_record& r = L->stage[L->stage_cur][L->stage_n++];
  r.slot = (uint32_t)L->slot;
  r.launches = launches;
  r.elapsed_ns = ns;
  L->gpu_ns_host += ns;
  L->segments.fetch_add(1, std::memory_order_relaxed);
}

static void flush_stage_locked(gh_live* L, bool force) {
  if (!L->acct || L->stage_n == 0) return;
  if (!force && L->stage_n < gh_cfg.flush_records) return;
  int b = L->stage_cur;
  {  // the pod's gpu_mem counter rides along: the launch mirrors it into device memory and the totals page
    uint64_t fr = 0, tot = 0;
    gh_mem_local(&fr, &tot);
    gemhook_acct_set_mem(L->acct, (uint32_t)L->slot, tot - fr, tot);
  }
  if (gh_acct_push_async(L->acct, L->stage[b], L->stage_n) == 0 && L->stage_done[b]) {
    CUstream as = (CUstream)(uintptr_t)gemhook_acct_stream(L->acct);
    L->stage_inflight[b] = GH_CALL(cuEventRecord, L->stage_done[b], as) == CUDA_SUCCESS;
  }
  L->stage_cur = b ^ 1;
  L->stage_n = 0;
  // the buffer we switch to was handed to the copy engine one flush ago: it must be done before we overwrite it
  // (it always is in practice -- a flush happens every GEMHOOK_FLUSH_RECORDS segments -- so this never blocks)
  if (L->stage_inflight[b ^ 1]) {
    GH_CALL(cuEventSynchronize, L->stage_done[b ^ 1]);
    L->stage_inflight[b ^ 1] = false;
  }
  if (force) {
    gemhook_acct_sync(L->acct);
    L->stage_inflight[0] = L->stage_inflight[1] = false;
  }
}

// Returns false when the launches that follow are NOT covered by an accounting segment (their stream is being
// captured, or the marker could not be recorded): the caller then keeps the gate closed, so the next launch comes
// through here again instead of running unaccounted on the fast path.
static bool seg_begin_locked(gh_live* L, CUstream stream, int64_t now) {
  if (gh_cfg.dry_run || !L->cuda_ready) return true;
  // events are a ring: never re-record one that an unresolved segment still refers to
  if (L->npending > SEG_EVENTS / 2) resolve_pending_locked(L, true);
  if (L->seg_open && L->seg_spans_sync) {  // the running segment continues across the sync we just passed
    if (L->seg_sync_return_ns && now > L->seg_sync_return_ns) L->seg_idle_ns += (uint64_t)(now - L->seg_sync_return_ns);
    L->seg_spans_sync = false;
    if (!stream_capturing(stream, true)) L->seg_stream = stream;  // (its end marker must go to a stream we may record on)
    return true;
  }
  if (L->seg_open && !L->seg_end_recorded) return true;  // still running (the gate was closed for another reason): continue it
  if (stream_capturing(stream, true)) return false;      // launches into a capturing stream do not run now: nothing to time
  L->seg_head = (L->seg_head + 1) % SEG_EVENTS;
  if (GH_CALL(cuEventRecord, L->seg_ev[L->seg_head], stream) != CUDA_SUCCESS) return false;
  L->seg_open = true;
  L->seg_end_recorded = false;
  L->seg_spans_sync = false;
  L->seg_stream = stream;
  L->seg_begin_host_ns = now;
  L->seg_idle_ns = 0;
  L->seg_first_launch = gh_total_launches();
  return true;
}

// The launching thread is about to wait for a new token: whatever it has launched so far ends the running segment
// (the wait itself must not be accounted).  Recorded by the launching thread, so no launch can slip past the marker.
static void seg_close_for_renewal_locked(gh_live* L) {
  if (gh_cfg.dry_run || !L->cuda_ready || !L->seg_open || L->seg_end_recorded) return;
  int64_t now = gh_now_ns();
  if (L->seg_spans_sync && L->seg_sync_return_ns && now > L->seg_sync_return_ns) L->seg_idle_ns += (uint64_t)(now - L->seg_sync_return_ns);
  L->seg_spans_sync = false;
  uint64_t n = gh_total_launches();
  int nxt = (L->seg_head + 1) % SEG_EVENTS;
  if (n != L->seg_first_launch && L->npending < SEG_EVENTS - 2 && !stream_capturing(L->seg_stream) &&
      GH_CALL(cuEventRecord, L->seg_ev[nxt], L->seg_stream) == CUDA_SUCCESS) {
    L->pending[L->npending++] = {L->seg_head, nxt, (uint32_t)(n - L->seg_first_launch), L->seg_idle_ns};
    L->seg_head = nxt;
  }
  L->seg_open = false;
}

// every K launches inside a burst (GEMHOOK_SEG_LAUNCHES=K): close the running segment at this point
void gh_segment_tick(CUstream stream) {
  gh_live* L = g_live;
  if (!L || !L->cuda_ready || gh_cfg.dry_run) return;
  pthread_mutex_lock(&L->mu);
  if (L->seg_open && L->npending < SEG_EVENTS / 2 - 1 && !stream_capturing(stream, true)) {
    int nxt = (L->seg_head + 1) % SEG_EVENTS;
    if (GH_CALL(cuEventRecord, L->seg_ev[nxt], stream) == CUDA_SUCCESS) {
      uint64_t n = gh_total_launches();
      L->pending[L->npending++] = {L->seg_head, nxt, (uint32_t)(n - L->seg_first_launch), L->seg_idle_ns};
      L->seg_head = nxt;
      L->seg_first_launch = n;
      L->seg_stream = stream;
      L->seg_idle_ns = 0;
      L->seg_begin_host_ns = gh_now_ns();
    }
  }
  pthread_mutex_unlock(&L->mu);
}

// before a synchronising driver call: mark the end of the running segment on its stream, so that the
// event carries the completion time of the burst's last kernel rather than the host's return time
static void sync_pre(bool force) {
  gh_live* L = g_live;
  if (!L || !L->cuda_ready || gh_cfg.dry_run) return;
  pthread_mutex_lock(&L->mu);
  if (!L->seg_open && L->npending == 0 && L->stage_n < gh_cfg.flush_records) {
    pthread_mutex_unlock(&L->mu);
    return;
  }
  int fresh = 0;
  bool old_enough = force || (gh_now_ns() - L->seg_begin_host_ns >= (int64_t)gh_cfg.seg_min_us * 1000);
  if (L->seg_open && !L->seg_end_recorded && !old_enough) {
    L->seg_spans_sync = true;  // keep it open across this sync
  } else if (L->seg_open && !L->seg_end_recorded && stream_capturing(L->seg_stream)) {
    L->seg_open = false;  // its stream went into capture mode meanwhile: the segment cannot be closed there, drop it
    L->seg_spans_sync = false;
  } else if (L->seg_open && !L->seg_end_recorded && L->npending < SEG_EVENTS - 2) {
    if (L->seg_spans_sync && L->seg_sync_return_ns) {  // forced close while idle after a merged sync
      int64_t t = gh_now_ns();
      if (t > L->seg_sync_return_ns) L->seg_idle_ns += (uint64_t)(t - L->seg_sync_return_ns);
      L->seg_spans_sync = false;
    }
    int nxt = (L->seg_head + 1) % SEG_EVENTS;
    if (GH_CALL(cuEventRecord, L->seg_ev[nxt], L->seg_stream) == CUDA_SUCCESS) {
      uint64_t n = gh_total_launches();
      L->pending[L->npending++] = {L->seg_head, nxt, (uint32_t)(n - L->seg_first_launch), L->seg_idle_ns};
      L->seg_head = nxt;
      L->seg_end_recorded = true;
      fresh = 1;
    }
  }
  // Deferred bookkeeping, done HERE on purpose: the application thread reaches a synchronising call ahead
  // of the GPU (its launches are still queued), so resolving the PREVIOUS bursts' event pairs and pushing
  // records to the device overlaps with the GPU draining the queue.  Doing it after the sync returned
  // would add host time while the GPU sits idle (measured: cuEventElapsedTime 2.7 us, i.e. ~0.13 % of a
  // 1024-launch burst).
  resolve_pending_locked(L, false, fresh);
  flush_stage_locked(L, false);
  publish_usage_locked(L);  // totals of the reduce kernels that have completed so far -> pool slot (live export)
  pthread_mutex_unlock(&L->mu);
}
void gh_host_sync_pre(void) { sync_pre(false); }

// cuStreamDestroy pre-hook: the stream is still valid here, so a segment open on it gets its end marker now
// (afterwards L->seg_stream would dangle -- the tracker thread records on it).
void gh_stream_destroyed(CUstream stream) {
  gh_live* L = g_live;
  if (!L || !L->cuda_ready || gh_cfg.dry_run || !stream) return;
  pthread_mutex_lock(&L->mu);
  if (L->seg_open && !L->seg_end_recorded && L->seg_stream == stream) {
    int nxt = (L->seg_head + 1) % SEG_EVENTS;
    if (L->npending < SEG_EVENTS - 2 && !stream_capturing(stream) &&
        GH_CALL(cuEventRecord, L->seg_ev[nxt], stream) == CUDA_SUCCESS) {
      if (L->seg_spans_sync && L->seg_sync_return_ns) {
        int64_t t = gh_now_ns();
        if (t > L->seg_sync_return_ns) L->seg_idle_ns += (uint64_t)(t - L->seg_sync_return_ns);
      }
      uint64_t n = gh_total_launches();
      L->pending[L->npending++] = {L->seg_head, nxt, (uint32_t)(n - L->seg_first_launch), L->seg_idle_ns};
      L->seg_head = nxt;
    }
    L->seg_open = false;
    L->seg_spans_sync = false;
    gh_gate_set(0u);  // the next launch opens a fresh segment on its own stream
  }
  if (L->seg_stream == stream) L->seg_stream = nullptr;
  pthread_mutex_unlock(&L->mu);
}

// all pending events are complete after a host sync: turn them into records
static void resolve_pending_locked(gh_live* L, bool may_block, int skip_last) {
  int kept = 0;
  int upto = L->npending - skip_last;
  for (int i = 0; i < L->npending; i++) {
    gh_live::Pending& p = L->pending[i];
    if (i >= upto) {  // just recorded: certainly not complete yet, leave it for the next call
      L->pending[kept++] = p;
      continue;
    }
    float ms = 0.f;
    CUresult r = GH_CALL(cuEventElapsedTime, &ms, L->seg_ev[p.ev_begin], L->seg_ev[p.ev_end]);
    if (r == CUDA_ERROR_NOT_READY && may_block) {
      GH_CALL(cuEventSynchronize, L->seg_ev[p.ev_end]);
      r = GH_CALL(cuEventElapsedTime, &ms, L->seg_ev[p.ev_begin], L->seg_ev[p.ev_end]);
    }
    if (r == CUDA_SUCCESS) {
      double ns = (double)ms * 1e6 - (double)p.idle_ns;
      stage_record(L, p.launches, ns > 0 ? (uint64_t)(ns + 0.5) : 0);
    } else if (r == CUDA_ERROR_NOT_READY) {
      L->pending[kept++] = p;  // a sync on another stream only: try again later
    }
  }
  L->npending = kept;
}

static void gate_edge_locked(gh_live* L, int64_t now) {
  gemhook_gate_host_sync(L->gate, now);
  gh_gate_set(0u);
}

static void host_sync_locked(gh_live* L, int64_t now) {
  gate_edge_locked(L, now);
  if (L->seg_open && L->seg_spans_sync && !L->seg_end_recorded) L->seg_sync_return_ns = now;  // merged: stays open
  else L->seg_open = false;
}

void gh_host_sync_post(void) {
  gh_live* L = gh_live_get();
  if (!L || !L->enabled) return;
  pthread_mutex_lock(&L->mu);
  L->host_syncs.store(L->host_syncs.load(std::memory_order_relaxed) + 1, std::memory_order_relaxed);  // (under mu: no lock prefix)
  L->last_sync_ns = gh_now_ns();
  host_sync_locked(L, L->last_sync_ns);  // nothing else here: the GPU is idle until the next launch arrives
  bool yield = false;
  if (gh_cfg.yield_on_idle && L->pool && !L->renewing && gemhook_gate_quota_ms(L->gate) > 0 &&
      gemhook_pool_others_waiting(L->pool, L->slot) && L->last_window_ms >= gh_cfg.yield_min_idle_ms) {
    // Work-conserving option (off by default: the reference keeps an idle token until it expires,
    // scheduler.cpp:501-521).  The GPU is drained right now, somebody else wants it, and we do not know when our
    // next burst comes -- but its last idle gap (sync -> next launch) was long enough to pay for a hand-over (a
    // launch storm that relaunches right after every sync keeps its token): hand it back; our next launch asks
    // again like any returning client.
    gemhook_gate_expire(L->gate);
    L->token_returned_ns = L->last_sync_ns;
    L->yielded = true;
    yield = true;
    L->yields.fetch_add(1, std::memory_order_relaxed);
  }
  pthread_mutex_unlock(&L->mu);
  if (yield) {
    gemhook_pool_release(L->pool, L->slot);
    pthread_mutex_lock(&L->trk_mu);  // let the tracker finish the token it was timing (no GPU work pending)
    if (!L->trk_done) {
      L->trk_intr = true;
      pthread_cond_signal(&L->trk_intr_cv);
    }
    pthread_mutex_unlock(&L->trk_mu);
  }
}

// ---- initialisation -----------------------------------------------------------------------------------
void gh_register_exit_hook(void);
static int read_quota_file_into_pool(gh_live* L) {
  if (!gh_cfg.quota_file[0]) return 0;
  return gemhook_pool_sync_quota_file(L->pool, gh_cfg.quota_file, gh_cfg.swap_columns);
}

static void live_init(void) {
  gh_config_load();
  gh_live* L = new gh_live();
  L->gate = gemhook_gate_new();
  if (gh_cfg.disabled || gh_driver_init() != 0) {
    L->enabled = false;
    g_live = L;
    gh_gate_set(1u);  // pass-through
    return;
  }
  L->enabled = true;
  if (gh_cfg.seg_launches) {
    uint32_t k = 1;
    while (k < gh_cfg.seg_launches) k <<= 1;
    gh_seg_mask = k - 1;
  }
  if (gh_cfg.transport == 1) {
    L->pool = gemhook_pool_open(gh_cfg.pool_path, 1, gh_cfg.base_quota_ms, gh_cfg.min_quota_ms, gh_cfg.window_ms, 0);
    if (!L->pool) {
      fatal_or_disable(L, gemhook_last_error());
    } else {
      read_quota_file_into_pool(L);
      L->slot = gemhook_pool_find(L->pool, gh_cfg.pod_name);
      if (L->slot >= 0) {
        gemhook_pool_reap(L->pool);  // leftovers of clients that died without cleaning up
        gemhook_pool_attach(L->pool, L->slot);
      }
      if (L->slot < 0) {
        char msg[256];
        snprintf(msg, sizeof(msg), "pod \"%s\" is not in the quota file / credit pool", gh_cfg.pod_name);
        fatal_or_disable(L, msg);
        L->slot = 0;
      }
    }
  } else if (!gh_cfg.scheduler_ip[0]) {
    // reference: exit(-1) when /kubeshare/library/schedulerIP.txt is missing (hook.cpp:234-237)
    fatal_or_disable(L, "scheduler IP file missing (set GEMHOOK_SCHEDULER_IP or /kubeshare/library/schedulerIP.txt)");
  }
  if (gh_cfg.token_trace[0]) {
    L->trace_cap = 1u << 16;
    L->trace = (gh_live::TokenTrace*)calloc(L->trace_cap, sizeof(gh_live::TokenTrace));
    if (!L->trace) L->trace_cap = 0;
  }
  g_live = L;
  if (!L->enabled) gh_gate_set(1u);
  // registered after the driver's own atexit handlers (cuInit ran before the first intercepted call),
  // so it runs BEFORE them and CUDA is still usable for the final flush
  gh_register_exit_hook();
  GH_INFO("gemhook ready: pod \"%s\", transport %s, segment mask %#x", gh_cfg.pod_name,
          gh_cfg.transport ? "pool" : "tcp", gh_seg_mask);
}

gh_live* gh_live_get(void) {
  pthread_once(&live_once, live_init);
  return g_live;
}

// CUDA-side objects need a current context: created at the first launch (hook.cpp:754-764 does the same
// from initialize()).  Caller holds L->mu.
static void cuda_init_locked(gh_live* L) {
  if (L->cuda_ready) return;
  L->cuda_ready = true;
  pthread_condattr_t attr;
  pthread_condattr_init(&attr);
  pthread_condattr_setclock(&attr, CLOCK_MONOTONIC);
  pthread_cond_init(&L->trk_intr_cv, &attr);
  GH_CALL(cuCtxGetCurrent, &L->ctx);
  if (!gh_cfg.dry_run) {
    GH_CALL(cuEventCreate, &L->ev_token, CU_EVENT_DEFAULT);
    // blocking sync: the tracker thread must not spin against the launching thread while it waits for the drain (a
    // client pinned to one CPU would lose a scheduler timeslice per token)
    GH_CALL(cuEventCreate, &L->ev_drain, CU_EVENT_BLOCKING_SYNC);
    for (int i = 0; i < SEG_EVENTS; i++) GH_CALL(cuEventCreate, &L->seg_ev[i], CU_EVENT_DEFAULT);
    uint32_t nslots = L->pool ? (uint32_t)gemhook_pool_nslots(L->pool) : 1;
    if (nslots < 1) nslots = 1;
    L->acct = gemhook_acct_create(nslots, 1u << 14);
    if (!L->acct) {
      // no silent CPU accounting: say so, keep gating alive
      fprintf(stderr, "[gemhook %d] device accounting unavailable: %s\n", (int)getpid(), gemhook_last_error());
    } else {
      L->stage_cap = gh_cfg.flush_records * 4 > 1024 ? gh_cfg.flush_records * 4 : 1024;
      for (int b = 0; b < 2; b++) {
        void* hp = nullptr;
        if (GH_CALL(cuMemHostAlloc, &hp, L->stage_cap * sizeof(gemhook_record), CU_MEMHOSTALLOC_PORTABLE) == CUDA_SUCCESS)
          L->stage[b] = (gemhook_record*)hp;
        GH_CALL(cuEventCreate, &L->stage_done[b], CU_EVENT_DISABLE_TIMING);
      }
      if (L->pool) {
        // "shared-pinned": the credit pool is page-locked and device-mapped, so device code (and peers'
        // device code) can read the same words the host arbitrates on
        size_t bytes = 0;
        void* base = gh_pool_region(L->pool, &bytes);
        CUresult r = GH_CALL(cuMemHostRegister_v2, base, bytes, CU_MEMHOSTREGISTER_PORTABLE | CU_MEMHOSTREGISTER_DEVICEMAP);
        GH_DEBUG("cuMemHostRegister(pool, %zu B) -> %d", bytes, (int)r);
      }
    }
  }
  pthread_create(&L->trk_tid, nullptr, tracker_main, L);
  pthread_detach(L->trk_tid);
  // first token request; its quota is discarded so that the first launch renews immediately (hook.cpp:766-768).
  // It goes through the same one-at-a-time protocol as renewals: a second application thread arriving meanwhile
  // waits on renew_cv instead of posting a competing request for the same slot.
  L->renewing = true;
  pthread_mutex_unlock(&L->mu);
  token_from_scheduler(L, 0.0, 0.0);
  pthread_mutex_lock(&L->mu);
  L->renewing = false;
  pthread_cond_broadcast(&L->renew_cv);
}

// ---- launch slow path ---------------------------------------------------------------------------------
void gh_launch_slow(CUstream stream) {
  gh_live* L = gh_live_get();
  if (!L || !L->enabled) return;
  pthread_mutex_lock(&L->mu);
  L->slow_path.store(L->slow_path.load(std::memory_order_relaxed) + 1, std::memory_order_relaxed);
  while (L->renewing) pthread_cond_wait(&L->renew_cv, &L->mu);
  // The reference initialises inside the FIRST hooked call, whatever it is (pthread_once in every intercept,
  // hook.cpp:781-783); here the CUDA-side part waits for the first launch.  The time it takes (module load, pinned
  // allocations, the first token -- up to a whole quota when a peer holds the GPU) is not application idle time: the
  // predictors see this launch at the moment it was issued, or the span would sit in the window predictor for its whole
  // validity and switch off the burst doubling of estimate_full_burst (hook.cpp:412) for the first seconds of the run.
  int64_t now;
  if (!L->cuda_ready) {
    now = gh_now_ns();
    cuda_init_locked(L);
    while (L->renewing) pthread_cond_wait(&L->renew_cv, &L->mu);
  } else {
    now = gh_now_ns();
  }
  if (L->last_sync_ns) {  // the application's most recent idle gap: decides whether yielding at syncs pays off
    L->last_window_ms = (double)(now - L->last_sync_ns) / 1e6;
    L->last_sync_ns = 0;
  }
  if (L->enabled && gemhook_gate_launch_begin(L->gate, now)) {
    L->renewing = true;
    seg_close_for_renewal_locked(L);
    pthread_mutex_unlock(&L->mu);
    wait_tracker(L);  // GPU drained, overuse known
    pthread_mutex_lock(&L->mu);
    double overuse = 0, next_burst = 0;
    int64_t t_req = gh_now_ns();
    gemhook_gate_renew_request(L->gate, t_req, &overuse, &next_burst);
    if (L->last_token_ns) {
      // what the scheduler's ledger will hold for the token we are returning:
      // end = min(now, start + quota + overuse) (reference scheduler.cpp:123-153)
      int64_t t_end = L->token_returned_ns ? L->token_returned_ns : t_req;
      L->token_returned_ns = 0;
      double held = (double)(t_end - L->last_token_ns) / 1e6, cap = L->last_quota_ms + overuse;
      L->accumulated_token_ms += held < cap ? held : cap;
    }
    pthread_mutex_unlock(&L->mu);
    if (L->pool) read_quota_file_into_pool(L);  // one stat() per token: pick up kubeshare-config's rewrites
    double quota = token_from_scheduler(L, overuse, next_burst);
    pthread_mutex_lock(&L->mu);
    // hook.cpp:543 records on the legacy stream; not while the application captures (see stream_capturing)
    L->token_ev_valid = !gh_cfg.dry_run && !stream_capturing((CUstream)0) &&
                        GH_CALL(cuEventRecord, L->ev_token, (CUstream)0) == CUDA_SUCCESS;
    now = gh_now_ns();
    gemhook_gate_renew_granted(L->gate, now, quota);
    L->last_token_ns = now;
    L->last_quota_ms = quota;
    pthread_mutex_lock(&L->trk_mu);
    L->trk_done = false;
    L->trk_armed = true;
    L->trk_intr = false;
    double q = quota > 0 ? quota : 0;
    if (q > 8.64e7) q = 8.64e7;  // clamp absurd quotas to a day to keep the timespec sane
    L->trk_deadline_ns = now + (int64_t)(q * 1e6);
    pthread_cond_signal(&L->trk_start_cv);
    pthread_mutex_unlock(&L->trk_mu);
    L->renewing = false;
    pthread_cond_broadcast(&L->renew_cv);
  }
  gemhook_gate_launch_end(L->gate, now);  // (`now` was re-read after a renewal; otherwise it is a few tens of ns old)
  // (while a capture is going on the gate stays closed: a capture is short, and the first launch after it must open a
  //  segment -- measured on the box: with the gate left open the graph replays that followed ran unaccounted)
  if (seg_begin_locked(L, stream, now)) gh_gate_set(1u);
  pthread_mutex_unlock(&L->mu);
}

// ---- introspection ------------------------------------------------------------------------------------
GH_EXPORT int gemhook_flush(void) {
  gh_live* L = g_live;
  if (!L || !L->cuda_ready || gh_cfg.dry_run) return 0;
  pthread_mutex_lock(&L->mu);
  if (L->seg_open && !L->seg_end_recorded) {
    pthread_mutex_unlock(&L->mu);
    sync_pre(true);
    pthread_mutex_lock(&L->mu);
    L->seg_open = false;
  }
  resolve_pending_locked(L, true);
  flush_stage_locked(L, true);
  publish_usage_locked(L, true);
  pthread_mutex_unlock(&L->mu);
  return 0;
}

GH_EXPORT int gemhook_get_stats(gemhook_stats* out) {
  if (!out) return -1;
  memset(out, 0, sizeof(*out));
  gh_live* L = g_live;
  if (!L) return -1;
  uint64_t n = gh_total_launches();
  out->launches = n;
  out->slow_path = L->slow_path.load();
  out->fast_path = n - (out->slow_path < n ? out->slow_path : n);
  out->token_requests = L->token_requests.load();
  out->host_syncs = L->host_syncs.load();
  out->segments = L->segments.load();
  out->acct_kernels = L->acct ? gemhook_acct_kernel_launches(L->acct) : 0;
  if (L->acct) {
    uint64_t tot[GEMHOOK_MAX_SLOTS * 3];
    if (gemhook_acct_read_totals(L->acct, tot, nullptr) == 0) out->gpu_ns = tot[L->slot * 3];
  }
  gh_mem_info(&out->mem_used, &out->mem_limit);
  out->mem_used = out->mem_limit - out->mem_used;  // gh_mem_info returns (free, total)
  out->quota_ms = gemhook_gate_quota_ms(L->gate);
  out->overuse_ms = gemhook_gate_overuse_ms(L->gate);
  out->token_wait_ms = (double)L->token_wait_ns.load() / 1e6;
  out->accumulated_token_ms = L->accumulated_token_ms;
  return 0;
}

// GEMHOOK_STATS_FILE=<path with optional %d for the pid>: one JSON object written at process exit, used
// by bench.py and the tests to collect per-client numbers without touching the application.
uint64_t gh_mem_denied(void);
void gh_register_exit_hook(void);
static void write_token_trace(gh_live* L) {
  if (!L->trace || !gh_cfg.token_trace[0]) return;
  char path[600];
  snprintf(path, sizeof(path), gh_cfg.token_trace, (int)getpid());
  FILE* f = fopen(path, "w");
  if (!f) return;
  for (uint32_t i = 0; i < L->trace_n; i++)
    fprintf(f, "{\"pod\": \"%s\", \"t_ms\": %.6f, \"overuse_ms\": %.17g, \"burst_ms\": %.17g, \"quota_ms\": %.17g, \"forwarded\": %d}\n", gh_cfg.pod_name,
            L->trace[i].t_ms, L->trace[i].overuse_ms, L->trace[i].burst_ms, L->trace[i].quota_ms, L->trace[i].forwarded);
  fclose(f);
}

static void write_stats_file(void) {
  const char* pat = getenv("GEMHOOK_STATS_FILE");
  gh_live* L = g_live;
  if (!L) return;
  gemhook_flush();  // the last segments reach the device totals and the pool slot whether or not anybody asked for a file
  if (L->enabled && L->pool) {
    gemhook_pool_release(L->pool, L->slot);  // do not make peers wait for a timeout (no-op while a sibling process runs)
    gemhook_pool_detach(L->pool);            // bytes this process never freed go back to the pod's budget
  }
  write_token_trace(L);
  if (!pat || !*pat) return;
  gemhook_stats s;
  gemhook_get_stats(&s);
  char path[600];
  snprintf(path, sizeof(path), pat, (int)getpid());
  FILE* f = fopen(path, "w");
  if (!f) return;
  fprintf(f,
          "{\"pod\": \"%s\", \"pid\": %d, \"launches\": %llu, \"fast_path\": %llu, \"slow_path\": %llu, "
          "\"token_requests\": %llu, \"host_syncs\": %llu, \"segments\": %llu, \"acct_kernels\": %llu, "
          "\"gpu_ns\": %llu, \"gpu_ns_host\": %llu, \"mem_used\": %llu, \"mem_limit\": %llu, \"allocs_denied\": %llu, "
          "\"quota_ms\": %.6f, \"overuse_ms\": %.6f, \"token_wait_ms\": %.6f, \"accumulated_token_ms\": %.6f, \"yields\": %llu",
          gh_cfg.pod_name, (int)getpid(), (unsigned long long)s.launches, (unsigned long long)s.fast_path,
          (unsigned long long)s.slow_path, (unsigned long long)s.token_requests, (unsigned long long)s.host_syncs,
          (unsigned long long)s.segments, (unsigned long long)s.acct_kernels, (unsigned long long)s.gpu_ns,
          (unsigned long long)L->gpu_ns_host, (unsigned long long)s.mem_used, (unsigned long long)s.mem_limit,
          (unsigned long long)gh_mem_denied(), s.quota_ms, s.overuse_ms, s.token_wait_ms, s.accumulated_token_ms,
          (unsigned long long)L->yields.load());
  if (gh_cfg.hook_debug) {  // CU_HOOK_DEBUG=1: the per-symbol call counters (reference hookInfo::call_count)
    const char* const* names = nullptr;
    const uint64_t* counts = nullptr;
    size_t n = gemhook_call_counts(&names, &counts);
    fprintf(f, ", \"calls\": {");
    for (size_t i = 0; i < n; i++) fprintf(f, "%s\"%s\": %llu", i ? ", " : "", names[i], (unsigned long long)counts[i]);
    fprintf(f, "}");
  }
  fprintf(f, "}\n");
  fclose(f);
}

void gh_register_exit_hook(void) { atexit(write_stats_file); }

// accessors for gh_mem.cpp / gh_interpose.cpp
extern "C" gemhook_pool* gh_live_pool(void) { return g_live ? g_live->pool : nullptr; }
extern "C" int gh_live_slot(void) { return g_live ? g_live->slot : 0; }
extern "C" int gh_live_enabled(void) { return (g_live && g_live->enabled) ? 1 : 0; }
