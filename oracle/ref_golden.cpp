/*
 * oracle/ref_golden.cpp -- TEST INFRASTRUCTURE ONLY: golden-vector generator.
 *
 * Links the reference's OWN object code (comm.o, predictor.o, scheduler.o, pod-manager.o, built by
 * oracle/Makefile from /root/reference/Gemini/src with main() renamed by -Dmain=...) and drives it
 * under a VIRTUAL clock: this executable defines clock_gettime() and pthread_cond_timedwait(), so
 * std::chrono::steady_clock::now(), get_timespec_after() and the scheduler's "sleep until the
 * window moves" all run on `vnow_ns` below.  The emitted JSON is committed as
 * tests/golden/ref_golden.json (see tests/golden/make_golden.py); tests compare the restatement in
 * oracle/gemini_oracle.cpp against it.  Runs only where /root/reference exists.
 *
 * Usage: ref_golden wire <POD_NAME> | predictor <seed> <nops> | schd <seed> <nsteps> <cfgdir> <cfgfile> <base> <min> <win> <mean_gap_ms> <hist_every>
 *        | pmgr <seed> <nops>
 */
#include <errno.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <list>
#include <map>
#include <string>
#include <vector>

#include "comm.h"
#include "predictor.h"
#include "scheduler.h"

/* ---------------- virtual clock ---------------- */
static int64_t vnow_ns = 1000LL * 1000000000LL;  // constant-initialised: valid during static init
static std::vector<int64_t> sleeps_ns;           // wake-up instants recorded by timedwait below

extern "C" int clock_gettime(clockid_t, struct timespec *ts) {
  ts->tv_sec = vnow_ns / 1000000000LL;
  ts->tv_nsec = vnow_ns % 1000000000LL;
  return 0;
}
extern "C" int pthread_cond_timedwait(pthread_cond_t *, pthread_mutex_t *, const struct timespec *abst) {
  int64_t t = (int64_t)abst->tv_sec * 1000000000LL + abst->tv_nsec;
  if (t > vnow_ns) vnow_ns = t;
  else vnow_ns += 1000;  // guarantee progress
  sleeps_ns.push_back(vnow_ns);
  return ETIMEDOUT;
}

/* hDEBUG 4-arg overload missing from a non-_DEBUG debug.o (see ref_compat.cpp, item 2) */
void hDEBUG(const char *, const char *, long, const char *, ...) {}

/* ---------------- reference globals / functions we drive ---------------- */
extern std::list<History> history_list;
extern std::list<candidate_t> candidates;
extern std::map<std::string, ClientInfo *> client_info_map;
extern double QUOTA, MIN_QUOTA, WINDOW_SIZE;
extern char limit_file_name[], limit_file_dir[];
extern pthread_mutex_t candidate_mutex;
void read_resource_config();
candidate_t select_candidate();
void handle_message(int client_sock, char *message);

// gem-pmgr is a second program with clashing global names (log_name, sig_handler); its vectors come
// from the live binary instead (tests/golden/make_golden.py, "pmgr_live").

/* ---------------- helpers ---------------- */
static uint64_t lcg_state;
static uint32_t lcg() {
  lcg_state = lcg_state * 6364136223846793005ULL + 1442695040888963407ULL;
  return (uint32_t)(lcg_state >> 33);
}
static void hex(const char *buf, size_t n) {
  for (size_t i = 0; i < n; i++) printf("%02x", (unsigned char)buf[i]);
}

/* ---------------- wire ---------------- */
static int do_wire(const char *name) {
  setenv("POD_NAME", name, 1);
  char b[REQ_MSG_LEN];
  printf("{\"name\": \"%s\", \"requests\": [", name);
  struct {
    comm_request_t t;
    double a, b;
    size_t bytes;
    int alloc;
  } cases[] = {{REQ_QUOTA, 0.0, 0.0, 0, 0},        {REQ_MEM_LIMIT, 0, 0, 0, 0},
               {REQ_MEM_UPDATE, 0, 0, 4096, 1},    {REQ_QUOTA, 12.625, 301.5, 0, 0},
               {REQ_MEM_UPDATE, 0, 0, 8589934592ULL, 0}, {REQ_QUOTA, 1e-3, 9999.999, 0, 0},
               {REQ_MEM_UPDATE, 0, 0, 18446744073709551615ULL, 1}};
  int n = sizeof(cases) / sizeof(cases[0]);
  for (int i = 0; i < n; i++) {
    memset(b, 0, sizeof(b));
    reqid_t id;
    if (cases[i].t == REQ_QUOTA) id = prepare_request(b, REQ_QUOTA, cases[i].a, cases[i].b);
    else if (cases[i].t == REQ_MEM_UPDATE) id = prepare_request(b, REQ_MEM_UPDATE, cases[i].bytes, cases[i].alloc);
    else id = prepare_request(b, REQ_MEM_LIMIT);
    char *pn; size_t pl; reqid_t pid; comm_request_t pt;
    char *att = parse_request(b, &pn, &pl, &pid, &pt);
    printf("%s{\"type\": %d, \"id\": %d, \"overuse\": %.17g, \"burst\": %.17g, \"bytes\": %zu, \"alloc\": %d, "
           "\"hex\": \"", i ? ", " : "", (int)cases[i].t, id, cases[i].a, cases[i].b, cases[i].bytes, cases[i].alloc);
    hex(b, REQ_MSG_LEN);
    printf("\", \"parsed_name\": \"%s\", \"parsed_len\": %zu, \"parsed_id\": %d, \"parsed_type\": %d, \"payload_off\": %ld}",
           pn, pl, pid, (int)pt, (long)(att - b));
  }
  printf("], \"responses\": [");
  char r[RSP_MSG_LEN];
  memset(r, 0, sizeof(r)); size_t l0 = prepare_response(r, REQ_QUOTA, 7, 127.5);
  printf("{\"type\": 0, \"id\": 7, \"quota\": 127.5, \"len\": %zu, \"hex\": \"", l0); hex(r, RSP_MSG_LEN);
  memset(r, 0, sizeof(r)); size_t l1 = prepare_response(r, REQ_MEM_LIMIT, 8, (size_t)4096, (size_t)8589934592ULL);
  printf("\"}, {\"type\": 1, \"id\": 8, \"used\": 4096, \"total\": 8589934592, \"len\": %zu, \"hex\": \"", l1); hex(r, RSP_MSG_LEN);
  memset(r, 0, sizeof(r)); size_t l2 = prepare_response(r, REQ_MEM_UPDATE, 9, 1);
  printf("\"}, {\"type\": 2, \"id\": 9, \"verdict\": 1, \"len\": %zu, \"hex\": \"", l2); hex(r, RSP_MSG_LEN);
  printf("\"}]}\n");
  return 0;
}

/* ---------------- predictor ---------------- */
static int do_predictor(uint64_t seed, int nops, double thres) {
  lcg_state = seed;
  Predictor p("golden", thres);
  printf("{\"seed\": %lu, \"thres\": %.17g, \"t0_ns\": %ld, \"ops\": [", (unsigned long)seed, thres, (long)vnow_ns);
  for (int i = 0; i < nops; i++) {
    // time steps: mostly sub-ms .. tens of ms, sometimes seconds (to cross the 3000 ms horizon)
    uint32_t r = lcg() % 100;
    int64_t dt;
    if (r < 50) dt = lcg() % 1500000;            // < 1.5 ms
    else if (r < 85) dt = lcg() % 40000000;      // < 40 ms
    else if (r < 97) dt = lcg() % 900000000;     // < 0.9 s
    else dt = 1000000000LL + lcg() % 3000000000u;  // 1..4 s
    vnow_ns += dt;
    uint32_t op = lcg() % 100;
    const char *name;
    if (op < 40) { p.record_start(); name = "start"; }
    else if (op < 80) { p.record_stop(); name = "stop"; }
    else if (op < 86) { p.interrupt(); name = "interrupt"; }
    else { name = "peek"; }
    double pu = p.predict_unmerged(), pm = p.predict_merged();
    printf("%s{\"op\": \"%s\", \"t_ns\": %ld, \"unmerged\": %.17g, \"merged\": %.17g, \"on_u\": %d, \"on_m\": %d}",
           i ? ", " : "", name, (long)vnow_ns, pu, pm, p.ongoing_unmerged() ? 1 : 0, p.ongoing_merged() ? 1 : 0);
  }
  printf("]}\n");
  return 0;
}

/* ---------------- scheduler ---------------- */
static void dump_history() {
  printf("[");
  bool first = true;
  for (auto &h : history_list) {
    printf("%s[\"%s\", %.17g, %.17g]", first ? "" : ", ", h.name.c_str(), h.start, h.end);
    first = false;
  }
  printf("]");
}

static int do_schd(uint64_t seed, int nsteps, const char *cfgdir, const char *cfgfile, double base, double minq, double win,
                   double mean_gap_ms, int hist_every) {
  lcg_state = seed;
  QUOTA = base; MIN_QUOTA = minq; WINDOW_SIZE = win;
  strncpy(limit_file_dir, cfgdir, 4000);
  strncpy(limit_file_name, cfgfile, 4000);
  int64_t start_ns = vnow_ns;  // == PROGRESS_START (static init read the same constant)
  FILE *saved = stderr; (void)saved;
  read_resource_config();
  std::vector<std::string> names;
  for (auto &kv : client_info_map) names.push_back(kv.first);
  printf("{\"seed\": %lu, \"base\": %.17g, \"min\": %.17g, \"window\": %.17g, \"start_ns\": %ld, \"clients\": [",
         (unsigned long)seed, base, minq, win, (long)start_ns);
  for (size_t i = 0; i < names.size(); i++)
    printf("%s[\"%s\", %.17g, %.17g, %zu]", i ? ", " : "", names[i].c_str(), client_info_map[names[i]]->get_min_fraction(),
           client_info_map[names[i]]->get_max_fraction(), client_info_map[names[i]]->gpu_mem_limit);
  printf("], \"steps\": [");
  std::map<std::string, bool> waiting;
  char msg[REQ_MSG_LEN];
  for (int s = 0; s < nsteps; s++) {
    // advance time
    int64_t dt = (int64_t)((lcg() % 2000) / 1000.0 * mean_gap_ms * 1e6);
    vnow_ns += dt;
    printf("%s{\"t_ns\": %ld", s ? ", " : "", (long)vnow_ns);
    // a random non-waiting client posts REQ_QUOTA (sometimes two do)
    int posts = 1 + (lcg() % 4 == 0);
    printf(", \"requests\": [");
    bool firstreq = true;
    for (int q = 0; q < posts; q++) {
      std::string nm = names[lcg() % names.size()];
      if (waiting[nm]) continue;
      double overuse = (lcg() % 3 == 0) ? (lcg() % 5000) / 1000.0 : 0.0;
      double burst = (lcg() % 5 == 0) ? 0.0 : (lcg() % 400000) / 1000.0;
      // build the 80-byte request exactly as a client would: [u64 len][name][0][id][type][overuse][burst]
      memset(msg, 0, sizeof(msg));
      size_t pos = 0, nl = nm.size();
      append_msg_data(msg, pos, nl);
      memcpy(msg + pos, nm.c_str(), nl); pos += nl;
      append_msg_data(msg, pos, '\0');
      append_msg_data(msg, pos, (reqid_t)s);
      append_msg_data(msg, pos, REQ_QUOTA);
      append_msg_data(msg, pos, overuse);
      append_msg_data(msg, pos, burst);
      handle_message(-1, msg);
      waiting[nm] = true;
      printf("%s[\"%s\", %.17g, %.17g]", firstreq ? "" : ", ", nm.c_str(), overuse, burst);
      firstreq = false;
    }
    printf("]");
    // scheduler daemon: one decision if anyone waits (scheduler.cpp:468-479)
    if (!candidates.empty()) {
      sleeps_ns.clear();
      pthread_mutex_lock(&candidate_mutex);
      candidate_t sel = select_candidate();
      double quota = client_info_map[sel.name]->get_quota();
      client_info_map[sel.name]->Record(quota);
      pthread_mutex_unlock(&candidate_mutex);
      waiting[sel.name] = false;
      printf(", \"selected\": \"%s\", \"quota\": %.17g, \"wakeups_ns\": [", sel.name.c_str(), quota);
      for (size_t i = 0; i < sleeps_ns.size(); i++) printf("%s%ld", i ? ", " : "", (long)sleeps_ns[i]);
      printf("], \"t_after_ns\": %ld", (long)vnow_ns);
    } else {
      printf(", \"selected\": null");
    }
    if (s % hist_every == 0 || s == nsteps - 1) {
      printf(", \"history\": ");
      dump_history();
    }
    printf("}");
  }
  printf("]}\n");
  return 0;
}

int main(int argc, char **argv) {
  if (argc >= 3 && !strcmp(argv[1], "wire")) return do_wire(argv[2]);
  if (argc >= 5 && !strcmp(argv[1], "predictor")) return do_predictor(strtoull(argv[2], 0, 0), atoi(argv[3]), atof(argv[4]));
  if (argc >= 11 && !strcmp(argv[1], "schd"))
    return do_schd(strtoull(argv[2], 0, 0), atoi(argv[3]), argv[4], argv[5], atof(argv[6]), atof(argv[7]), atof(argv[8]), atof(argv[9]),
                   atoi(argv[10]));
  fprintf(stderr, "usage: see header\n");
  return 2;
}
