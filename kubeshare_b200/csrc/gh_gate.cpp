// gh_gate.cpp -- burst/window predictor and the launch-gate state machine, clock injected.
//
// Behavioural contract (reference Gemini/src): predictor.cpp:41-186 (sliding 3000 ms maximum of plain
// and merged period lengths, measured in whole microseconds / 1e3), hook.cpp:402-418 (estimate_full_burst),
// hook.cpp:508-558 (launch pre-hook), hook.cpp:334-340 (host_sync_call), hook.cpp:456-502 (overuse tracker).
// The structure is ours: no mutexes inside (the live hook serialises slow-path calls itself, and the
// fast path only reads one word), a flat growable ring instead of std::deque, and every method takes
// the time as an argument so the same object is driven by CLOCK_MONOTONIC in the hook and by a trace
// replayer in the parity tests.
#include <stdlib.h>
#include <string.h>

#include "gh_internal.h"

namespace {

const int64_t NEVER = INT64_MAX;      // "no period open"
const int64_t LONG_AGO = INT64_MIN;   // "merged period has no end yet"
const int64_t HORIZON_MS = 3000;      // PREDICT_MAX_KEEP, predictor.h:27
const double SCHD_OVERHEAD_MS = 2.0;  // hook.cpp:176

// (a - b) in whole microseconds / 1e3, two's-complement wrap like the reference's chrono arithmetic
inline double span_ms(int64_t a, int64_t b) {
  int64_t d = (int64_t)((uint64_t)a - (uint64_t)b);
  return (double)(d / 1000) / 1e3;
}
inline int64_t span_whole_ms(int64_t a, int64_t b) {
  int64_t d = (int64_t)((uint64_t)a - (uint64_t)b);
  return d / 1000000;
}

// Sliding-window maximum: values kept non-increasing front->back, so the front is the max of the
// window and every add is amortised O(1).
struct MaxWindow {
  struct Item {
    int64_t at;
    double v;
  };
  Item* buf = nullptr;
  uint32_t cap = 0, head = 0, len = 0;

  ~MaxWindow() { free(buf); }
  Item& at_index(uint32_t i) { return buf[(head + i) & (cap - 1)]; }
  void grow() {
    uint32_t ncap = cap ? cap * 2 : 64;
    Item* nb = (Item*)malloc(sizeof(Item) * ncap);
    for (uint32_t i = 0; i < len; i++) nb[i] = at_index(i);
    free(buf);
    buf = nb;
    cap = ncap;
    head = 0;
  }
  void push(double v, int64_t at) {
    while (len && at_index(len - 1).v < v) len--;
    if (len == cap) grow();
    at_index(len) = Item{at, v};
    len++;
  }
  void expire(int64_t now) {
    while (len && span_whole_ms(now, at_index(0).at) > HORIZON_MS) {
      head = (head + 1) & (cap - 1);
      len--;
    }
  }
  double max() { return len ? at_index(0).v : 0.0; }
};

}  // namespace

struct gemhook_predictor {
  double merge_gap_ms;
  int64_t open_at = NEVER;         // current plain period
  int64_t merged_open_at = NEVER;  // current merged period
  int64_t merged_last_end = LONG_AGO;
  MaxWindow plain, merged;

  bool open() const { return open_at != NEVER; }
  bool merged_open() const { return merged_open_at != NEVER; }

  void stop(int64_t now) {
    if (open()) {
      plain.push(span_ms(now, open_at), now);
      merged_last_end = now;
      merged.push(span_ms(merged_last_end, merged_open_at), now);
    }
    open_at = NEVER;
  }
  void start(int64_t now) {
    if (open()) return;
    open_at = now;
    double gap = span_ms(open_at, merged_last_end);
    if (!merged_open() || gap > merge_gap_ms) {
      merged_open_at = open_at;
      merged_last_end = LONG_AGO;
    }
  }
  void interrupt() {
    open_at = NEVER;
    merged_open_at = NEVER;
    merged_last_end = LONG_AGO;
  }
  double predict_plain(int64_t now) {
    plain.expire(now);
    return plain.max();
  }
  double predict_merged(int64_t now) {
    merged.expire(now);
    return merged.max();
  }
};

GH_EXPORT gemhook_predictor* gemhook_predictor_new(double merge_thres_ms) {
  gemhook_predictor* p = new gemhook_predictor();
  p->merge_gap_ms = merge_thres_ms;
  return p;
}
GH_EXPORT void gemhook_predictor_free(gemhook_predictor* p) { delete p; }
GH_EXPORT void gemhook_predictor_record_start(gemhook_predictor* p, int64_t now_ns) { p->start(now_ns); }
GH_EXPORT void gemhook_predictor_record_stop(gemhook_predictor* p, int64_t now_ns) { p->stop(now_ns); }
GH_EXPORT void gemhook_predictor_interrupt(gemhook_predictor* p) { p->interrupt(); }
GH_EXPORT int gemhook_predictor_ongoing_unmerged(const gemhook_predictor* p) { return p->open() ? 1 : 0; }
GH_EXPORT int gemhook_predictor_ongoing_merged(const gemhook_predictor* p) { return p->merged_open() ? 1 : 0; }
GH_EXPORT double gemhook_predictor_predict_unmerged(gemhook_predictor* p, int64_t now_ns) { return p->predict_plain(now_ns); }
GH_EXPORT double gemhook_predictor_predict_merged(gemhook_predictor* p, int64_t now_ns) { return p->predict_merged(now_ns); }

GH_EXPORT double gemhook_estimate_full_burst(double burst_ms, double window_ms) {
  if (burst_ms < 1e-9) return 0.0;  // no valid burst data yet
  return window_ms < SCHD_OVERHEAD_MS ? burst_ms * 2 : burst_ms;
}

// ---- launch gate -------------------------------------------------------------------------------------
struct gemhook_gate {
  gemhook_predictor burst, window;
  double quota_ms = 0.0;
  double overuse_ms = 0.0;
  int64_t token_at = 0;      // request_start; zero like the reference's zero-initialised timespec
  bool tracker_done = true;  // hook.cpp:767
  gemhook_gate() {
    burst.merge_gap_ms = SCHD_OVERHEAD_MS;  // hook.cpp:177
    window.merge_gap_ms = 0.0;              // hook.cpp:178
  }
};

// us_since() of the reference works on timespec fields: seconds difference * 1e6 + nanosecond-field
// difference / 1000 truncated toward zero (hook.cpp:202-206).
static inline int64_t us_since_fields(int64_t begin_ns, int64_t now_ns) {
  const int64_t G = 1000000000LL;
  return (now_ns / G - begin_ns / G) * 1000000LL + (now_ns % G - begin_ns % G) / 1000LL;
}

GH_EXPORT gemhook_gate* gemhook_gate_new(void) { return new gemhook_gate(); }
GH_EXPORT void gemhook_gate_free(gemhook_gate* g) { delete g; }

GH_EXPORT int gemhook_gate_launch_begin(gemhook_gate* g, int64_t now) {
  g->window.stop(now);
  if (g->burst.open()) return 0;  // burst already running: launch freely
  double held_ms = (double)us_since_fields(g->token_at, now) / 1e3;
  return (held_ms + g->burst.predict_plain(now) >= g->quota_ms) ? 1 : 0;
}

GH_EXPORT void gemhook_gate_renew_request(gemhook_gate* g, int64_t now, double* overuse_ms, double* next_burst_ms) {
  double nb = gemhook_estimate_full_burst(g->burst.predict_merged(now), g->window.predict_merged(now));
  g->window.interrupt();  // the window opened by the tracker's sync is not a real idle period
  if (overuse_ms) *overuse_ms = g->overuse_ms;
  if (next_burst_ms) *next_burst_ms = nb;
}

GH_EXPORT void gemhook_gate_renew_granted(gemhook_gate* g, int64_t now, double quota_ms) {
  g->token_at = now;
  g->quota_ms = quota_ms;
  g->tracker_done = false;
}

GH_EXPORT void gemhook_gate_launch_end(gemhook_gate* g, int64_t now) { g->burst.start(now); }

GH_EXPORT void gemhook_gate_host_sync(gemhook_gate* g, int64_t now) {
  g->burst.stop(now);
  g->window.start(now);
}

GH_EXPORT void gemhook_gate_tracker_fire(gemhook_gate* g, int64_t now, float elapsed_ms) {
  gemhook_gate_host_sync(g, now);
  double over = (double)elapsed_ms - g->quota_ms;
  g->overuse_ms = over > 0.0 ? over : 0.0;
  g->tracker_done = true;
}

GH_EXPORT int gemhook_gate_tracker_complete(const gemhook_gate* g) { return g->tracker_done ? 1 : 0; }
GH_EXPORT double gemhook_gate_quota_ms(const gemhook_gate* g) { return g->quota_ms; }
GH_EXPORT double gemhook_gate_overuse_ms(const gemhook_gate* g) { return g->overuse_ms; }
GH_EXPORT int gemhook_gate_is_open(const gemhook_gate* g) { return g->burst.open() ? 1 : 0; }
GH_EXPORT void gemhook_gate_expire(gemhook_gate* g) { g->quota_ms = 0.0; }
GH_EXPORT double gemhook_gate_predicted_window_ms(gemhook_gate* g, int64_t now_ns) { return g->window.predict_merged(now_ns); }
