/*
 * oracle/gemini_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see gemini_oracle.h).
 *
 * CPU restatement of the reference Gemini hot path with an injected clock.  It is written
 * from the reference's behaviour (file:line cited per function, paths relative to
 * /root/reference/Gemini/src), not from its text, and is pinned against outputs of the
 * reference's own object code in tests/golden/ (see tests/test_oracle_golden.py).
 *
 * Clock conventions: hook/pmgr side = int64 nanoseconds of a monotonic clock (the reference
 * reads steady_clock / CLOCK_MONOTONIC and truncates differences to whole microseconds before
 * dividing by 1e3); scheduler side = double milliseconds since scheduler start
 * (scheduler.cpp:107-109).
 */
#include "gemini_oracle.h"

#include <pthread.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <list>
#include <map>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

namespace {

const int64_t TP_MAX = std::numeric_limits<int64_t>::max();  // timepoint_t::max()
const int64_t TP_MIN = std::numeric_limits<int64_t>::min();  // timepoint_t::min()

// duration_cast<microseconds>(a - b).count() / 1e3 with the reference's wrap-around when one
// side is time_point::min() (predictor.cpp:113).
inline double diff_ms_from_us(int64_t a, int64_t b) {
  int64_t d = (int64_t)((uint64_t)a - (uint64_t)b);
  return (double)(d / 1000) / 1e3;
}
inline int64_t diff_whole_ms(int64_t a, int64_t b) {
  int64_t d = (int64_t)((uint64_t)a - (uint64_t)b);
  return d / 1000000;
}

template <typename T>
void put(uint8_t *buf, size_t &pos, T v) {
  memcpy(buf + pos, &v, sizeof(T));
  pos += sizeof(T);
}
template <typename T>
T take(const uint8_t *buf, size_t &pos) {
  T v;
  memcpy(&v, buf + pos, sizeof(T));
  pos += sizeof(T);
  return v;
}

// RecordKeeper (predictor.cpp:41-60): sliding maximum over the last keep_ms milliseconds kept as a
// non-increasing deque.
struct SlidingMax {
  std::deque<std::pair<int64_t, double>> q;
  int64_t keep_ms = 3000;  // PREDICT_MAX_KEEP, predictor.h:27
  void add(double v, int64_t tp) {
    while (!q.empty() && q.back().second < v) q.pop_back();
    q.emplace_back(tp, v);
  }
  void expire(int64_t tp) {
    while (!q.empty() && diff_whole_ms(tp, q.front().first) > keep_ms) q.pop_front();
  }
  double top() const { return q.empty() ? 0.0 : q.front().second; }
};

}  // namespace

/* ======================================================================================
 * wire format
 * ==================================================================================== */

// comm.cpp:26-63 prepare_request: [u64 name_len][name][NUL][i32 id][i32 type] + payload.
size_t orc_wire_request(uint8_t *buf, const char *name, int32_t req_id, int32_t type,
                        double overuse_ms, double burst_ms, uint64_t bytes, int32_t is_alloc) {
  size_t pos = 0;
  uint64_t n = strlen(name);
  put<uint64_t>(buf, pos, n);
  memcpy(buf + pos, name, n);
  pos += n;
  put<char>(buf, pos, '\0');
  put<int32_t>(buf, pos, req_id);
  put<int32_t>(buf, pos, type);
  if (type == ORC_REQ_QUOTA) {
    put<double>(buf, pos, overuse_ms);
    put<double>(buf, pos, burst_ms);
  } else if (type == ORC_REQ_MEM_UPDATE) {
    put<uint64_t>(buf, pos, bytes);
    put<int32_t>(buf, pos, is_alloc);
  }
  return pos;
}

// comm.cpp:66-86 parse_request.
size_t orc_wire_parse_request(const uint8_t *buf, char *name_out, uint64_t *name_len,
                              int32_t *req_id, int32_t *type) {
  size_t pos = 0;
  uint64_t n = take<uint64_t>(buf, pos);
  if (name_out) {
    size_t c = n < 71 ? (size_t)n : 71;
    memcpy(name_out, buf + 8, c);
    name_out[c] = 0;
  }
  pos += n + 1;
  int32_t id = take<int32_t>(buf, pos);
  int32_t ty = take<int32_t>(buf, pos);
  if (name_len) *name_len = n;
  if (req_id) *req_id = id;
  if (type) *type = ty;
  return pos;
}

// comm.cpp:88-109 prepare_response: [i32 id] + payload.
size_t orc_wire_response(uint8_t *buf, int32_t type, int32_t req_id, double quota_ms,
                         uint64_t mem_used, uint64_t mem_total, int32_t verdict) {
  size_t pos = 0;
  put<int32_t>(buf, pos, req_id);
  if (type == ORC_REQ_QUOTA) {
    put<double>(buf, pos, quota_ms);
  } else if (type == ORC_REQ_MEM_UPDATE) {
    put<int32_t>(buf, pos, verdict);
  } else if (type == ORC_REQ_MEM_LIMIT) {
    put<uint64_t>(buf, pos, mem_used);
    put<uint64_t>(buf, pos, mem_total);
  }
  return pos;
}

/* ======================================================================================
 * Predictor
 * ==================================================================================== */

struct orc_pred {
  double merge_thres;
  int64_t begin = TP_MAX;       // period_begin_
  int64_t long_begin = TP_MAX;  // long_period_begin_
  int64_t long_end = TP_MIN;    // long_period_end_
  SlidingMax plain, merged;     // normal_records, long_records
};

orc_pred *orc_pred_new(double merge_thres_ms) {
  orc_pred *p = new orc_pred();
  p->merge_thres = merge_thres_ms;
  return p;
}
void orc_pred_free(orc_pred *p) { delete p; }

int orc_pred_ongoing_unmerged(const orc_pred *p) { return p->begin != TP_MAX; }   // predictor.cpp:77
int orc_pred_ongoing_merged(const orc_pred *p) { return p->long_begin != TP_MAX; }  // :80

// predictor.cpp:83-102
void orc_pred_record_stop(orc_pred *p, int64_t now) {
  if (orc_pred_ongoing_unmerged(p)) {
    double dur = diff_ms_from_us(now, p->begin);
    p->plain.add(dur, now);
    p->long_end = now;
    p->merged.add(diff_ms_from_us(p->long_end, p->long_begin), now);
  }
  p->begin = TP_MAX;
}

// predictor.cpp:105-124
void orc_pred_record_start(orc_pred *p, int64_t now) {
  if (orc_pred_ongoing_unmerged(p)) return;
  p->begin = now;
  double gap = diff_ms_from_us(p->begin, p->long_end);
  if (!orc_pred_ongoing_merged(p) || gap > p->merge_thres) {
    p->long_begin = p->begin;
    p->long_end = TP_MIN;
  }
}

// predictor.cpp:128-138
void orc_pred_interrupt(orc_pred *p) {
  p->begin = TP_MAX;
  p->long_begin = TP_MAX;
  p->long_end = TP_MIN;
}

// predictor.cpp:141-152
double orc_pred_predict_unmerged(orc_pred *p, int64_t now) {
  p->plain.expire(now);
  return p->plain.top();
}
// predictor.cpp:155-165
double orc_pred_predict_merged(orc_pred *p, int64_t now) {
  p->merged.expire(now);
  return p->merged.top();
}

/* ======================================================================================
 * hook launch gate
 * ==================================================================================== */

static const double kSchdOverheadMs = 2.0;  // hook.cpp:176

// hook.cpp:402-418
double orc_estimate_full_burst(double measured_burst, double measured_window) {
  if (measured_burst < 1e-9) return 0.0;
  double full = measured_burst;
  if (measured_window < kSchdOverheadMs) full *= 2;
  return full;
}

// hook.cpp:202-206, with timespec fields rebuilt from the ns clock.
int64_t orc_us_since(int64_t begin_ns, int64_t now_ns) {
  int64_t bs = begin_ns / 1000000000LL, bn = begin_ns % 1000000000LL;
  int64_t ns = now_ns / 1000000000LL, nn = now_ns % 1000000000LL;
  return (ns - bs) * 1000000LL + (nn - bn) / 1000LL;
}

struct orc_hook {
  orc_pred *burst = orc_pred_new(kSchdOverheadMs);  // hook.cpp:177
  orc_pred *window = orc_pred_new(0.0);             // hook.cpp:178
  double quota_time = 0.0;                          // hook.cpp:171
  double overuse = 0.0;                             // hook.cpp:172
  int64_t request_start = 0;                        // hook.cpp:184 (zero-initialised global)
  bool trk_complete = true;                         // hook.cpp:767
};

orc_hook *orc_hook_new(void) { return new orc_hook(); }
void orc_hook_free(orc_hook *h) {
  if (!h) return;
  orc_pred_free(h->burst);
  orc_pred_free(h->window);
  delete h;
}

// hook.cpp:515-521
int orc_hook_launch_begin(orc_hook *h, int64_t now) {
  orc_pred_record_stop(h->window, now);
  if (orc_pred_ongoing_unmerged(h->burst)) return 0;
  double since_ms = (double)orc_us_since(h->request_start, now) / 1e3;
  return (since_ms + orc_pred_predict_unmerged(h->burst, now) >= h->quota_time) ? 1 : 0;
}

// hook.cpp:523-538 (the tracker wait at 527-533 is modelled by the caller)
void orc_hook_renew_request(orc_hook *h, int64_t now, double *overuse_ms, double *next_burst_ms) {
  double nb = orc_estimate_full_burst(orc_pred_predict_merged(h->burst, now),
                                      orc_pred_predict_merged(h->window, now));
  orc_pred_interrupt(h->window);
  if (overuse_ms) *overuse_ms = h->overuse;
  if (next_burst_ms) *next_burst_ms = nb;
}

// hook.cpp:541-552
void orc_hook_renew_granted(orc_hook *h, int64_t now, double new_quota_ms) {
  h->request_start = now;
  h->quota_time = new_quota_ms;
  h->trk_complete = false;
}

void orc_hook_launch_end(orc_hook *h, int64_t now) { orc_pred_record_start(h->burst, now); }  // :554

// hook.cpp:334-340
void orc_hook_host_sync(orc_hook *h, int64_t now) {
  orc_pred_record_stop(h->burst, now);
  orc_pred_record_start(h->window, now);
}

// hook.cpp:482-499
void orc_hook_tracker_fire(orc_hook *h, int64_t now, float elapsed_ms) {
  orc_hook_host_sync(h, now);
  h->overuse = std::max(0.0, (double)elapsed_ms - h->quota_time);
  h->trk_complete = true;
}
int orc_hook_tracker_complete(const orc_hook *h) { return h->trk_complete ? 1 : 0; }
double orc_hook_quota_ms(const orc_hook *h) { return h->quota_time; }
double orc_hook_overuse_ms(const orc_hook *h) { return h->overuse; }

/* ======================================================================================
 * hook-side gpu_mem rules
 * ==================================================================================== */

// hook.cpp:347-367 + 590-601: remain = total - used (size_t arithmetic), reject iff bytes > remain.
int orc_mem_prehook_allows(uint64_t bytesize, uint64_t used, uint64_t total) {
  uint64_t remain = total - used;
  return bytesize > remain ? 0 : 1;
}

// hook.cpp:638-680.  CUarray_format values: U8=0x01 U16=0x02 U32=0x03 S8=0x08 S16=0x09 S32=0x0a
// HALF=0x10 FLOAT=0x20 (cuda.h).  Other formats fall off the reference's switch (undefined); 0 here.
uint64_t orc_array_bytes(uint64_t w, uint64_t h, uint64_t d, uint32_t channels, uint32_t format,
                         int is3d) {
  uint64_t fs = 0;
  switch (format) {
    case 0x01: case 0x08: fs = 1; break;
    case 0x02: case 0x09: case 0x10: fs = 2; break;
    case 0x03: case 0x0a: case 0x20: fs = 4; break;
    default: fs = 0; break;
  }
  uint64_t n = is3d ? w * h * d * channels : w * h * channels;
  return n * fs;
}

/* ======================================================================================
 * gem-pmgr
 * ==================================================================================== */

struct orc_pmgr {
  uint64_t limit = 0, used = 0;            // pod-manager.cpp:91
  std::map<int, uint64_t> per_conn;        // allocation_map, :92
  double pod_overuse = 0.0;                // :97
  std::map<int, double> burst_by_conn;     // client_burst_map, :98
  double pod_quota = 0.0;                  // :100
  int64_t quota_tp = 0;                    // quota_updated_tp, :101, set at :221
};

orc_pmgr *orc_pmgr_new(uint64_t gpu_mem_limit, int64_t start_ns) {
  orc_pmgr *p = new orc_pmgr();
  p->limit = gpu_mem_limit;
  p->quota_tp = start_ns;
  return p;
}
void orc_pmgr_free(orc_pmgr *p) { delete p; }

// pod-manager.cpp:262-268
void orc_pmgr_connect(orc_pmgr *p, int conn) {
  p->per_conn.insert(std::make_pair(conn, (uint64_t)0));
  p->burst_by_conn.insert(std::make_pair(conn, 0.0));
}
// pod-manager.cpp:533-545
void orc_pmgr_disconnect(orc_pmgr *p, int conn) {
  p->used -= p->per_conn[conn];
  p->per_conn.erase(conn);
  p->burst_by_conn.erase(conn);
}
// pod-manager.cpp:295-313
int orc_pmgr_mem_update(orc_pmgr *p, int conn, uint64_t bytes, int is_alloc) {
  if (is_alloc) {
    if (p->used + bytes > p->limit) return 0;
    p->used += bytes;
    p->per_conn[conn] += bytes;
  } else {
    p->used -= bytes;
    p->per_conn[conn] -= bytes;
  }
  return 1;
}
void orc_pmgr_mem_info(const orc_pmgr *p, uint64_t *used, uint64_t *limit) {  // :501-504
  if (used) *used = p->used;
  if (limit) *limit = p->limit;
}
// pod-manager.cpp:316-473 (single in-flight request; the condvar choreography is not modelled)
int orc_pmgr_kernel_launch(orc_pmgr *p, int conn, int64_t now, double overuse_ms, double burst_ms,
                           double *fwd_overuse_ms, double *fwd_burst_ms, double *remain_ms) {
  p->pod_overuse = std::max(overuse_ms, p->pod_overuse);
  p->burst_by_conn[conn] = burst_ms;
  double elapsed = diff_ms_from_us(now, p->quota_tp);
  if (elapsed + burst_ms > p->pod_quota) {
    double mx = 0.0;
    for (auto &kv : p->burst_by_conn) mx = std::max(kv.second, mx);
    if (fwd_overuse_ms) *fwd_overuse_ms = p->pod_overuse;
    if (fwd_burst_ms) *fwd_burst_ms = mx;
    return 1;
  }
  if (remain_ms) *remain_ms = p->pod_quota - elapsed;
  return 0;
}
double orc_pmgr_schd_reply(orc_pmgr *p, int64_t now, double quota_ms) {  // :422-427, 472
  p->pod_quota = quota_ms;
  p->quota_tp = now;
  p->pod_overuse = 0.0;
  return p->pod_quota - 0.0;
}

/* ======================================================================================
 * gem-schd
 * ==================================================================================== */

namespace {
struct Span {
  std::string who;
  double start, end;
};
struct Client {  // ClientInfo, scheduler.h:36-60, scheduler.cpp:111-174
  std::string name;
  double min_frac, max_frac, base_q, min_q, max_q;
  double quota, burst = 0.0;
  uint64_t mem_limit = 0;
};
struct Waiting {  // candidate_t
  std::string name;
  double arrived;
};
struct Ranked {  // valid_candidate_t
  double missing, remaining, usage, arrived;
  std::list<Waiting>::iterator it;
};
bool rank_before(const Ranked &a, const Ranked &b) {  // schd-priority.cpp:19-26
  if (a.missing > 0 && b.missing > 0)
    return a.missing / (a.missing + a.usage) > b.missing / (b.missing + b.usage);
  if (a.missing > 0 && b.missing < 0) return true;
  if (a.missing < 0 && b.missing > 0) return false;
  return a.usage < b.usage;
}
}  // namespace

struct orc_schd {
  double base_q, min_q, window;
  std::map<std::string, Client> clients;
  std::list<Span> ledger;       // history_list (pruned)
  std::list<Span> full_ledger;  // full_history (_DEBUG)
  std::list<Waiting> waiting;   // candidates

  // scheduler.cpp:281-367: usage per client in the window ending at `now`; prunes the ledger.
  void window_usage(double now, std::map<std::string, double> &usage, double &window_size,
                    double &window_start) {
    window_size = window;
    window_start = now - window;
    if (window_start < 0) window_size = now;
    ledger.remove_if([=](const Span &s) { return s.end < window_start; });

    struct Stamp {
      std::string who;
      double t;  // negative = start
    };
    std::vector<Stamp> stamps;
    for (const Span &s : ledger) {
      stamps.push_back({s.who, -s.start});
      stamps.push_back({s.who, s.end});
      usage[s.who] = 0;
    }
    std::sort(stamps.begin(), stamps.end(),
              [](Stamp a, Stamp b) { return std::abs(a.t) < std::abs(b.t); });

    std::vector<std::string> live;
    int live_cnt = 0;
    size_t k = 0;
    for (; k < stamps.size(); k++) {
      if (std::abs(stamps[k].t) <= window_start) {
        live_cnt++;
        live.push_back(stamps[k].who);
      } else {
        break;
      }
    }
    double cur = window_start;
    for (size_t i = k; i < stamps.size(); ++i) {
      for (size_t j = 0; j < live.size(); ++j) usage[live[j]] += (std::abs(stamps[i].t) - cur) / live_cnt;
      if (stamps[i].t < 0) {
        live.push_back(stamps[i].who);
        live_cnt++;
      } else {
        for (size_t j = 0; j < live.size(); ++j) {
          if (live[j] == stamps[i].who) {
            live.erase(live.begin() + j);
            break;
          }
        }
        live_cnt--;
      }
      cur = std::abs(stamps[i].t);
    }
  }
};

orc_schd *orc_schd_new(double base_quota_ms, double min_quota_ms, double window_ms) {
  orc_schd *s = new orc_schd();
  s->base_q = base_quota_ms;
  s->min_q = min_quota_ms;
  s->window = window_ms;
  return s;
}
void orc_schd_free(orc_schd *s) { delete s; }

// scheduler.cpp:203-212: a re-read replaces the ClientInfo (its adaptive quota restarts at BASE).
void orc_schd_set_client(orc_schd *s, const char *name, double min_frac, double max_frac,
                         uint64_t mem_limit) {
  Client c;
  c.name = name;
  c.min_frac = min_frac;
  c.max_frac = max_frac;
  c.base_q = s->base_q;
  c.min_q = s->min_q;
  c.max_q = max_frac * s->window;
  c.quota = s->base_q;
  c.mem_limit = mem_limit;
  s->clients[c.name] = c;
}

// scheduler.cpp:183-217 read_resource_config: `N` then N rows `name min max mem` split on blanks.
int orc_schd_load_config(orc_schd *s, const char *text) {
  std::istringstream in(text);
  int n = 0;
  if (!(in >> n)) return -1;
  for (int i = 0; i < n; i++) {
    std::string name;
    double mn = 0, mx = 0;
    uint64_t mem = 0;
    in >> name >> mn >> mx >> mem;
    orc_schd_set_client(s, name.c_str(), mn, mx, mem);
  }
  return n;
}
int orc_schd_has_client(const orc_schd *s, const char *name) { return s->clients.count(name) ? 1 : 0; }
uint64_t orc_schd_mem_limit(const orc_schd *s, const char *name) {
  auto it = s->clients.find(name);
  return it == s->clients.end() ? 0 : it->second.mem_limit;
}

// scheduler.cpp:402-429 handle_message(REQ_QUOTA) -> update_return_time (:123-142), set_burst, enqueue
int orc_schd_request(orc_schd *s, const char *name, double now, double overuse_ms, double burst_ms) {
  auto it = s->clients.find(name);
  if (it == s->clients.end()) return -1;  // :411-414 unknown client: dropped
  for (auto r = s->ledger.rbegin(); r != s->ledger.rend(); ++r) {
    if (r->who == it->first) {
      r->end = std::min(now, r->end + overuse_ms);
      break;
    }
  }
  for (auto r = s->full_ledger.rbegin(); r != s->full_ledger.rend(); ++r) {
    if (r->who == it->first) {
      r->end = std::min(now, r->end + overuse_ms);
      break;
    }
  }
  it->second.burst = burst_ms;
  s->waiting.push_back({it->first, now});
  return 0;
}

// scheduler.cpp:274-399 select_candidate (one pass of its loop)
int orc_schd_select(orc_schd *s, double now, char *name_out, double *sleep_ms) {
  if (s->waiting.empty()) return -1;
  std::map<std::string, double> usage;
  double wsize, wstart;
  // the reference prunes + collects stamps before the quick exit (:297-307), so prune first
  s->window_usage(now, usage, wsize, wstart);

  const std::string &head = s->waiting.front().name;
  bool seen = false;
  for (const Span &sp : s->ledger)
    if (sp.who == head) {
      seen = true;
      break;
    }
  if (!seen) {  // :312-320
    if (name_out) strcpy(name_out, head.c_str());
    s->waiting.pop_front();
    return 1;
  }

  std::vector<Ranked> ok;
  for (auto it = s->waiting.begin(); it != s->waiting.end(); ++it) {
    const Client &c = s->clients[it->name];
    double limit = c.max_frac * wsize;
    double require = c.min_frac * wsize;
    double missing = require - usage[it->name];
    double remaining = limit - usage[it->name];
    if (remaining > 0) ok.push_back({missing, remaining, usage[it->name], it->arrived, it});
  }
  if (ok.empty()) {  // :383-390
    if (sleep_ms) *sleep_ms = s->ledger.begin()->end - wstart;
    return 0;
  }
  std::sort(ok.begin(), ok.end(), rank_before);
  auto pick = ok.begin()->it;
  if (name_out) strcpy(name_out, pick->name.c_str());
  s->waiting.erase(pick);
  return 1;
}

// scheduler.cpp:160-174 get_quota + :144-153 Record
double orc_schd_grant(orc_schd *s, const char *name, double now) {
  Client &c = s->clients[name];
  if (c.burst < 1e-9) {
    c.quota = c.base_q;
  } else {
    c.quota = c.burst * 0.5 + c.quota * (1 - 0.5);
    c.quota = std::max(c.quota, c.min_q);
    c.quota = std::min(c.quota, c.max_q);
  }
  Span sp{c.name, now, now + c.quota};
  s->ledger.push_back(sp);
  s->full_ledger.push_back(sp);
  return c.quota;
}

double orc_schd_usage(orc_schd *s, const char *name, double now) {
  std::map<std::string, double> usage;
  double a, b;
  s->window_usage(now, usage, a, b);
  auto it = usage.find(name);
  return it == usage.end() ? 0.0 : it->second;
}

size_t orc_schd_history_len(const orc_schd *s) { return s->ledger.size(); }
int orc_schd_history_get(const orc_schd *s, size_t i, char *name_out, double *start_ms,
                         double *end_ms) {
  if (i >= s->ledger.size()) return -1;
  auto it = s->ledger.begin();
  std::advance(it, i);
  if (name_out) strcpy(name_out, it->who.c_str());
  if (start_ms) *start_ms = it->start;
  if (end_ms) *end_ms = it->end;
  return 0;
}
double orc_schd_accumulated_ms(const orc_schd *s, const char *name) {
  double t = 0;
  for (const Span &sp : s->full_ledger)
    if (sp.who == name) t += sp.end - sp.start;
  return t;
}
int orc_schd_priority(double a_missing, double a_usage, double b_missing, double b_usage) {
  Ranked a{a_missing, 0, a_usage, 0, {}}, b{b_missing, 0, b_usage, 0, {}};
  return rank_before(a, b) ? 1 : 0;
}

/* ======================================================================================
 * accounting reduction (CPU statement of the device kernel's contract; the reference has no
 * device-side accounting -- its only measurement is one event pair per token, hook.cpp:482-492)
 * ==================================================================================== */

void orc_acct_reduce(const orc_acct_record *rec, size_t n, uint32_t nslots, uint64_t *total_ns,
                     uint64_t *total_launches, uint64_t *total_records) {
  for (size_t i = 0; i < n; i++) {
    uint32_t s = rec[i].slot;
    if (s >= nslots) continue;
    total_ns[s] += rec[i].elapsed_ns;
    total_launches[s] += rec[i].launches;
    total_records[s] += 1;
  }
}

namespace {
struct MtJob {
  const orc_acct_record *rec;
  size_t n;
  uint32_t nslots;
  std::vector<uint64_t> ns, launches, records;
};
void *mt_worker(void *arg) {
  MtJob *j = (MtJob *)arg;
  orc_acct_reduce(j->rec, j->n, j->nslots, j->ns.data(), j->launches.data(), j->records.data());
  return nullptr;
}
}  // namespace

void orc_acct_reduce_mt(const orc_acct_record *rec, size_t n, uint32_t nslots, uint64_t *total_ns,
                        uint64_t *total_launches, uint64_t *total_records, int threads) {
  if (threads < 1) threads = 1;
  std::vector<MtJob> jobs(threads);
  std::vector<pthread_t> tids(threads);
  size_t chunk = (n + threads - 1) / threads;
  for (int t = 0; t < threads; t++) {
    size_t lo = std::min(n, (size_t)t * chunk), hi = std::min(n, lo + chunk);
    jobs[t].rec = rec + lo;
    jobs[t].n = hi - lo;
    jobs[t].nslots = nslots;
    jobs[t].ns.assign(nslots, 0);
    jobs[t].launches.assign(nslots, 0);
    jobs[t].records.assign(nslots, 0);
    pthread_create(&tids[t], nullptr, mt_worker, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) {
    pthread_join(tids[t], nullptr);
    for (uint32_t s = 0; s < nslots; s++) {
      total_ns[s] += jobs[t].ns[s];
      total_launches[s] += jobs[t].launches[s];
      total_records[s] += jobs[t].records[s];
    }
  }
}
