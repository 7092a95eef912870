/*
 * oracle/gemini_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * C API of the CPU restatement of the Gemini hot path (the hook's launch gate, predictor,
 * gpu_mem rules, wire format; gem-pmgr's pod counters; gem-schd's token policy), with an
 * INJECTED clock so launch traces replay deterministically.  Every function cites the
 * reference file:line it restates (paths relative to /root/reference/Gemini/src).
 *
 * Nothing in the product (kubeshare_b200/, libgemhook.so.1) may include, link or call this.
 * Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference.
 *
 * Pinning: tests/golden/ref_golden.json holds outputs of the reference's OWN object code
 * (predictor.o, comm.o, scheduler.o, pod-manager via the live gem-pmgr binary) under a virtual
 * clock, produced by oracle/ref_golden.cpp + tests/golden/make_golden.py; tests/test_oracle_*.py
 * check this restatement against them.
 */
#ifndef GEMINI_ORACLE_H
#define GEMINI_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- wire format (comm.h:28-31, comm.cpp:26-120) ------------------------------------ */
enum { ORC_REQ_QUOTA = 0, ORC_REQ_MEM_LIMIT = 1, ORC_REQ_MEM_UPDATE = 2 };
enum { ORC_REQ_LEN = 80, ORC_RSP_LEN = 40 };

/* Fills buf[80] (caller zeroes it first, as hook.cpp:353,381,431 do). Returns bytes used. */
size_t orc_wire_request(uint8_t *buf, const char *name, int32_t req_id, int32_t type,
                        double overuse_ms, double burst_ms, uint64_t bytes, int32_t is_alloc);
/* Parses a request; returns offset of payload. name_out must hold >= 72 bytes. */
size_t orc_wire_parse_request(const uint8_t *buf, char *name_out, uint64_t *name_len,
                              int32_t *req_id, int32_t *type);
size_t orc_wire_response(uint8_t *buf, int32_t type, int32_t req_id, double quota_ms,
                         uint64_t mem_used, uint64_t mem_total, int32_t verdict);

/* ---- Predictor / RecordKeeper (predictor.h:27-65, predictor.cpp:41-186) -------------- */
typedef struct orc_pred orc_pred;
orc_pred *orc_pred_new(double merge_thres_ms);
void orc_pred_free(orc_pred *);
void orc_pred_record_start(orc_pred *, int64_t now_ns);
void orc_pred_record_stop(orc_pred *, int64_t now_ns);
void orc_pred_interrupt(orc_pred *);
int orc_pred_ongoing_unmerged(const orc_pred *);
int orc_pred_ongoing_merged(const orc_pred *);
double orc_pred_predict_unmerged(orc_pred *, int64_t now_ns);
double orc_pred_predict_merged(orc_pred *, int64_t now_ns);

/* ---- hook launch gate (hook.cpp:402-418, 456-502, 508-558, 334-340) ------------------ */
double orc_estimate_full_burst(double measured_burst, double measured_window);

typedef struct orc_hook orc_hook;
orc_hook *orc_hook_new(void);
void orc_hook_free(orc_hook *);
/* cuLaunchKernel_prehook up to the renewal decision (hook.cpp:515-521).
 * Returns 1 when a token must be renewed (then call renew_request / renew_granted), else 0.
 * In both cases finish the launch with orc_hook_launch_end(). */
int orc_hook_launch_begin(orc_hook *, int64_t now_ns);
/* hook.cpp:523-538: next_burst estimate; the caller models the tracker wait (527-533) by calling
 * orc_hook_tracker_fire() first if the tracker had not completed.  Outputs the REQ_QUOTA payload. */
void orc_hook_renew_request(orc_hook *, int64_t now_ns, double *overuse_ms, double *next_burst_ms);
/* hook.cpp:541-552: token received at now_ns (cuevent_start recorded, request_start stamped). */
void orc_hook_renew_granted(orc_hook *, int64_t now_ns, double new_quota_ms);
/* hook.cpp:554: burst_predictor.record_start(). */
void orc_hook_launch_end(orc_hook *, int64_t now_ns);
/* host_sync_call (hook.cpp:334-340), run by the five sync post-hooks (696-722). */
void orc_hook_host_sync(orc_hook *, int64_t now_ns);
/* wait_cuda_kernels after its drain (hook.cpp:482-499): elapsed_ms = cudaEventElapsedTime
 * (cuevent_start -> drain event) as a float. */
void orc_hook_tracker_fire(orc_hook *, int64_t now_ns, float elapsed_ms);
int orc_hook_tracker_complete(const orc_hook *);
double orc_hook_quota_ms(const orc_hook *);
double orc_hook_overuse_ms(const orc_hook *);
/* us_since(request_start) (hook.cpp:202-206) with the timespec arithmetic kept exact. */
int64_t orc_us_since(int64_t begin_ns, int64_t now_ns);

/* ---- hook-side gpu_mem rules (hook.cpp:570-680, 857-872) ------------------------------ */
/* pre-hook test (hook.cpp:590-601): 1 = allowed, 0 = CUDA_ERROR_OUT_OF_MEMORY. */
int orc_mem_prehook_allows(uint64_t bytesize, uint64_t used, uint64_t total);
/* array byte rules (hook.cpp:638-680); format is the CUarray_format value. */
uint64_t orc_array_bytes(uint64_t w, uint64_t h, uint64_t d, uint32_t channels, uint32_t format,
                         int is3d);

/* ---- gem-pmgr (pod-manager.cpp:295-313, 316-473, 501-504, 533-545) -------------------- */
typedef struct orc_pmgr orc_pmgr;
orc_pmgr *orc_pmgr_new(uint64_t gpu_mem_limit, int64_t start_ns);
void orc_pmgr_free(orc_pmgr *);
void orc_pmgr_connect(orc_pmgr *, int conn);
void orc_pmgr_disconnect(orc_pmgr *, int conn); /* reclaim, pod-manager.cpp:533-545 */
int orc_pmgr_mem_update(orc_pmgr *, int conn, uint64_t bytes, int is_alloc);
void orc_pmgr_mem_info(const orc_pmgr *, uint64_t *used, uint64_t *limit);
/* hook_kernel_launch: returns 1 if the request must be forwarded to gem-schd (outputs the
 * forwarded payload), else 0 and *remain_ms is the reply. */
int orc_pmgr_kernel_launch(orc_pmgr *, int conn, int64_t now_ns, double overuse_ms, double burst_ms,
                           double *fwd_overuse_ms, double *fwd_burst_ms, double *remain_ms);
/* scheduler answered with quota at now_ns; returns the reply to the hook (pod_quota - 0). */
double orc_pmgr_schd_reply(orc_pmgr *, int64_t now_ns, double quota_ms);

/* ---- gem-schd (scheduler.cpp:123-174, 183-217, 274-399, 402-459; schd-priority.cpp) --- */
typedef struct orc_schd orc_schd;
orc_schd *orc_schd_new(double base_quota_ms, double min_quota_ms, double window_ms);
void orc_schd_free(orc_schd *);
/* one quota-file row in gem-schd's column order: name min_frac max_frac mem (scheduler.cpp:205) */
void orc_schd_set_client(orc_schd *, const char *name, double min_frac, double max_frac,
                         uint64_t mem_limit);
/* whole quota file text -> clients (read_resource_config). Returns client count or -1. */
int orc_schd_load_config(orc_schd *, const char *text);
int orc_schd_has_client(const orc_schd *, const char *name);
uint64_t orc_schd_mem_limit(const orc_schd *, const char *name);
/* handle_message(REQ_QUOTA) at now_ms (ms since scheduler start). 0 ok, -1 unknown client. */
int orc_schd_request(orc_schd *, const char *name, double now_ms, double overuse_ms,
                     double burst_ms);
/* select_candidate at now_ms. Returns 1 and writes the selected name (>= 64 bytes) -- the
 * candidate is removed; returns 0 when every candidate is at its limit and writes the
 * relative sleep (history.front().end - window_start, scheduler.cpp:385); -1 if no candidates. */
int orc_schd_select(orc_schd *, double now_ms, char *name_out, double *sleep_ms);
/* get_quota() + Record(quota) at now_ms (scheduler.cpp:475-479). */
double orc_schd_grant(orc_schd *, const char *name, double now_ms);
/* usage of `name` inside the window ending at now_ms, as select_candidate computes it. */
double orc_schd_usage(orc_schd *, const char *name, double now_ms);
size_t orc_schd_history_len(const orc_schd *);
/* ledger row i: returns 0 ok. name_out >= 64 bytes. */
int orc_schd_history_get(const orc_schd *, size_t i, char *name_out, double *start_ms,
                         double *end_ms);
/* sum over the FULL ledger (the _DEBUG full_history, scheduler.cpp:144-156, 693-714) of
 * end-start for one client: "accumulated GPU-ms" as SURVEY.md 8(a) defines it. */
double orc_schd_accumulated_ms(const orc_schd *, const char *name);
/* schd_priority comparator (schd-priority.cpp:19-26) on (missing, usage) pairs. */
int orc_schd_priority(double a_missing, double a_usage, double b_missing, double b_usage);

/* ---- accounting reduction: CPU statement of the device kernel's contract --------------- */
/* record = 16 bytes: u32 slot | u32 launches | u64 elapsed_ns.  slot >= nslots is ignored. */
typedef struct {
  uint32_t slot;
  uint32_t launches;
  uint64_t elapsed_ns;
} orc_acct_record;
void orc_acct_reduce(const orc_acct_record *rec, size_t n, uint32_t nslots, uint64_t *total_ns,
                     uint64_t *total_launches, uint64_t *total_records);
/* same, split over `threads` pthreads (the all-cores CPU baseline). */
void orc_acct_reduce_mt(const orc_acct_record *rec, size_t n, uint32_t nslots, uint64_t *total_ns,
                        uint64_t *total_launches, uint64_t *total_records, int threads);

#ifdef __cplusplus
}
#endif
#endif
