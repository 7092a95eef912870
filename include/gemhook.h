/*
 * include/gemhook.h -- C ABI of libgemhook.so.1 (B200-native replacement of Gemini's LD_PRELOAD hook).
 *
 * Two groups of entry points, both plain C (pointers + sizes, no C++/torch types):
 *
 *  (1) The HOOK SURFACE: the CUDA driver symbols the reference hook interposes.  An application (or
 *      its cudart) binds these exactly as it binds the reference's libgemhook.so.1 -- by symbol
 *      interposition under LD_PRELOAD, through the interposed dlsym(), or through the interposed
 *      cuGetProcAddress / cuGetProcAddress_v2.  Prototypes are the driver API's own (cuda.h); they are
 *      listed here with the reference line each one replaces.
 *
 *  (2) The CONTROL / ACCOUNTING API (gemhook_*): what a node agent, a test, or a Go/cgo binding calls:
 *      wire codec, launch-gate state machine with an injected clock, gpu_mem cap, shared credit pool,
 *      token policy, and the device accounting reduction.  INTEGRATION.md shows the cgo/ctypes stubs.
 *
 * All functions return 0 (or a CUresult of CUDA_SUCCESS) on success unless stated otherwise.
 * Reference paths are relative to /root/reference/Gemini/src.
 */
#ifndef GEMHOOK_H
#define GEMHOOK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GEMHOOK_ABI_VERSION 1

/* ===================================================================================================
 * (1) hook surface -- interposed symbols (reference hook.cpp:109-159 dlsym, :875-980 cuGetProcAddress,
 *     :1005-1062 the 20 wrappers, :857-872 mem-info overrides)
 * ===================================================================================================
 *   void *dlsym(void *handle, const char *symbol);                          hook.cpp:109
 *   CUresult cuGetProcAddress(const char*, void**, int, cuuint64_t);        hook.cpp:875   (legacy 4-arg)
 *   CUresult cuGetProcAddress_v2(const char*, void**, int, cuuint64_t,
 *                                CUdriverProcAddressQueryResult*);          (CUDA >= 12; reference lacks it)
 *   cuLaunchKernel                 hook.cpp:1049   pre: token gate (508-558)
 *   cuLaunchCooperativeKernel      hook.cpp:1056   pre: same gate (560-567)
 *   cuMemAlloc_v2                  hook.cpp:1021   pre/post: gpu_mem cap (590-617)
 *   cuMemAllocManaged              hook.cpp:1023   not accounted (619-627)
 *   cuMemAllocPitch_v2             hook.cpp:1026   cap on pitch*height (629-636)
 *   cuMemFree_v2                   hook.cpp:1030   pre: release bytes (570-581)
 *   cuArrayCreate_v2               hook.cpp:1033   cap on W*H*C*fmt (654-666)
 *   cuArray3DCreate_v2             hook.cpp:1036   cap on W*H*D*C*fmt (668-680)
 *   cuMipmappedArrayCreate         hook.cpp:1039   not accounted (682-694)
 *   cuArrayDestroy                 hook.cpp:1044   release (583)
 *   cuMipmappedArrayDestroy        hook.cpp:1045   release (585-587)
 *   cuMemGetInfo_v2                hook.cpp:865    virtualised: (limit-used, limit)
 *   cuDeviceTotalMem_v2            hook.cpp:857    virtualised: limit
 *   cuCtxSynchronize               hook.cpp:1018   post: host_sync_call (696-699)
 *   cuMemcpyAtoH_v2 / cuMemcpyDtoH_v2 / cuMemcpyHtoA_v2 / cuMemcpyHtoD_v2
 *                                  hook.cpp:1005-1017  post: host_sync_call (701-722)
 *      (the reference exports cuMemcpyDtoH C++-mangled by accident, hook.cpp:925-926; we export it.)
 *   Beyond the reference: cuLaunchKernelEx and cuGraphLaunch pass the token gate;
 *   cuMemAllocAsync / cuMemAllocFromPoolAsync / cuMemFreeAsync and cuMemCreate / cuMemRelease are charged against
 *   gpu_mem; every stream-taking hook has its _ptsz / _ptds twin; cuStreamSynchronize / cuEventSynchronize are
 *   exported and count as burst edges only with GEMHOOK_EXTRA_HOOKS=1; cuStreamDestroy_v2 closes an accounting
 *   segment still open on the dying stream; cuMemAllocHost_v2 / cuMemHostAlloc / cuMemFreeHost are charged only with
 *   GEMHOOK_ACCOUNT_HOST=1, cuMemAllocManaged / cuMipmappedArrayCreate only with GEMHOOK_ACCOUNT_MANAGED=1.
 */
const char *const *gemhook_hooked_symbols(size_t *count); /* names above, NULL-terminated */
/* per-symbol call counters, kept when CU_HOOK_DEBUG=1 (hook.cpp:87-100, 783, 991). Returns the number of counters;
 * names is NULL-terminated. */
size_t gemhook_call_counts(const char *const **names, const uint64_t **counts);

/* ===================================================================================================
 * (2a) wire codec -- comm.h:28-31, comm.cpp:26-120
 * =================================================================================================== */
enum { GEMHOOK_REQ_QUOTA = 0, GEMHOOK_REQ_MEM_LIMIT = 1, GEMHOOK_REQ_MEM_UPDATE = 2 };
enum { GEMHOOK_REQ_LEN = 80, GEMHOOK_RSP_LEN = 40 };

typedef struct gemhook_request {
  char name[72];      /* POD_NAME (<= 47 bytes for REQ_QUOTA, <= 51 for REQ_MEM_UPDATE) */
  int32_t req_id;
  int32_t type;
  double overuse_ms;  /* REQ_QUOTA */
  double burst_ms;    /* REQ_QUOTA */
  uint64_t bytes;     /* REQ_MEM_UPDATE */
  int32_t is_alloc;   /* REQ_MEM_UPDATE */
} gemhook_request;

typedef struct gemhook_response {
  int32_t req_id;
  double quota_ms;    /* REQ_QUOTA */
  uint64_t mem_used;  /* REQ_MEM_LIMIT */
  uint64_t mem_total; /* REQ_MEM_LIMIT */
  int32_t verdict;    /* REQ_MEM_UPDATE */
} gemhook_response;

/* out must hold 80 bytes and is fully written (zero padded). -1 if the name does not fit. */
int gemhook_wire_pack_request(const gemhook_request *req, uint8_t *out);
int gemhook_wire_unpack_request(const uint8_t *in, gemhook_request *req);
/* out must hold 40 bytes. */
int gemhook_wire_pack_response(int32_t type, const gemhook_response *rsp, uint8_t *out);
int gemhook_wire_unpack_response(int32_t type, const uint8_t *in, gemhook_response *rsp);

/* ===================================================================================================
 * (2b) launch gate with an injected clock -- hook.cpp:402-418, 456-502, 508-558, 334-340;
 *      predictor.cpp:41-186.  The live hook drives the same object with CLOCK_MONOTONIC.
 * =================================================================================================== */
typedef struct gemhook_gate gemhook_gate;
gemhook_gate *gemhook_gate_new(void);
void gemhook_gate_free(gemhook_gate *);
/* cuLaunchKernel_prehook up to the decision: 1 = token renewal needed, 0 = launch may proceed. */
int gemhook_gate_launch_begin(gemhook_gate *, int64_t now_ns);
/* renewal payload (overuse, next_burst) -- call after the tracker completed. */
void gemhook_gate_renew_request(gemhook_gate *, int64_t now_ns, double *overuse_ms, double *next_burst_ms);
void gemhook_gate_renew_granted(gemhook_gate *, int64_t now_ns, double quota_ms);
void gemhook_gate_launch_end(gemhook_gate *, int64_t now_ns);
void gemhook_gate_host_sync(gemhook_gate *, int64_t now_ns);
void gemhook_gate_tracker_fire(gemhook_gate *, int64_t now_ns, float elapsed_ms);
int gemhook_gate_tracker_complete(const gemhook_gate *);
double gemhook_gate_quota_ms(const gemhook_gate *);
double gemhook_gate_overuse_ms(const gemhook_gate *);
int gemhook_gate_is_open(const gemhook_gate *); /* fast-path word: burst ongoing */
/* give the current token up: the next launch renews (used by GEMHOOK_YIELD_ON_IDLE). */
void gemhook_gate_expire(gemhook_gate *);
/* largest idle window (sync -> next launch) seen in the last 3 s (window predictor, hook.cpp:178). */
double gemhook_gate_predicted_window_ms(gemhook_gate *, int64_t now_ns);
double gemhook_estimate_full_burst(double measured_burst_ms, double measured_window_ms);

/* stand-alone predictor (predictor.h:42-65) */
typedef struct gemhook_predictor gemhook_predictor;
gemhook_predictor *gemhook_predictor_new(double merge_thres_ms);
void gemhook_predictor_free(gemhook_predictor *);
void gemhook_predictor_record_start(gemhook_predictor *, int64_t now_ns);
void gemhook_predictor_record_stop(gemhook_predictor *, int64_t now_ns);
void gemhook_predictor_interrupt(gemhook_predictor *);
int gemhook_predictor_ongoing_unmerged(const gemhook_predictor *);
int gemhook_predictor_ongoing_merged(const gemhook_predictor *);
double gemhook_predictor_predict_unmerged(gemhook_predictor *, int64_t now_ns);
double gemhook_predictor_predict_merged(gemhook_predictor *, int64_t now_ns);

/* ===================================================================================================
 * (2c) shared credit pool + token policy -- replaces hook->gem-pmgr->gem-schd TCP round trips
 *      (hook.cpp:300-328, 425-446; pod-manager.cpp:295-473; scheduler.cpp:123-174, 274-529).
 *      One file-backed MAP_SHARED region per GPU; every co-resident client maps it (and pins its first pages with
 *      cuMemHostRegister so the device sees the same counters).  LOCK-FREE: every mutation is a copy-on-write
 *      transaction on a versioned state block published with one compare-and-swap; readers validate against the
 *      version word; a client killed at any point blocks nobody (gh_pool.cpp).
 * =================================================================================================== */
typedef struct gemhook_pool gemhook_pool;
/* create=1: create/initialise if absent (node agent or first hook); scheduler parameters as gem-schd's
 * -q -m -w (scheduler.cpp:555-604). start_ns: scheduler epoch on the caller's monotonic clock
 * (0 = now). */
gemhook_pool *gemhook_pool_open(const char *path, int create, double base_quota_ms, double min_quota_ms,
                                double window_ms, int64_t start_ns);
void gemhook_pool_close(gemhook_pool *);
/* Load a quota file text ("N\nname c2 c3 mem\n..."): columns 2/3 are min_frac/max_frac exactly as
 * gem-schd reads them (scheduler.cpp:205); swap_columns=1 reads them as kubeshare-config writes them
 * (limit request, pkg/config/query.go:56).  Returns client count, -1 on parse error. */
int gemhook_pool_load_config(gemhook_pool *, const char *text, int swap_columns);
/* stat the quota file and (re)load it only if it changed since the pool last saw it (gem-schd does this with
 * inotify, scheduler.cpp:219-265). 1 = reloaded, 0 = unchanged, -1 = unreadable / malformed. */
int gemhook_pool_sync_quota_file(gemhook_pool *, const char *path, int swap_columns);
int gemhook_pool_find(const gemhook_pool *, const char *name); /* slot index or -1 */
int gemhook_pool_nslots(const gemhook_pool *);
/* token policy, clock injected (now_ms = ms since pool start_ns, as scheduler.cpp:107-109) */
int gemhook_pool_request(gemhook_pool *, int slot, double now_ms, double overuse_ms, double burst_ms);
/* one scheduling decision: 1 = granted (slot_out, quota_out), 0 = nobody eligible (sleep_ms_out),
 * -1 = nobody waiting, -2 = a token is outstanding (sleep_ms_out = time to its deadline). */
int gemhook_pool_schedule(gemhook_pool *, double now_ms, int *slot_out, double *quota_out, double *sleep_ms_out);
double gemhook_pool_usage(gemhook_pool *, int slot, double now_ms);
size_t gemhook_pool_history(const gemhook_pool *, int *slots, double *starts, double *ends, size_t cap);
double gemhook_pool_accumulated_ms(const gemhook_pool *, int slot);
/* the pool's clock right now: ms since the pool file was created, on CLOCK_MONOTONIC -- the time base of the ledger
 * (gem-schd's ms_since_start, scheduler.cpp:107-109); lets a reader place ledger entries on its own clock. */
double gemhook_pool_now_ms(const gemhook_pool *);
/* blocking convenience used by the live hook: post request, arbitrate, wait for the grant. */
double gemhook_pool_acquire(gemhook_pool *, int slot, double overuse_ms, double burst_ms);
/* same; *forwarded (optional) = 1 if the request went to the scheduler policy (the reply is the client's new adaptive
 * quota), 0 if the pod-level rule answered it with the remaining pod quota (pod-manager.cpp:472). */
double gemhook_pool_acquire_ex(gemhook_pool *, int slot, double overuse_ms, double burst_ms, int *forwarded);
/* hand an outstanding token back early (client exit); the next waiter is scheduled immediately. */
void gemhook_pool_release(gemhook_pool *, int slot);
/* 1 if another client is waiting for the token (lock-free peek). */
int gemhook_pool_others_waiting(const gemhook_pool *, int slot);
/* declare the outstanding token timed out (scheduler.cpp:507-510) without touching the ledger. */
void gemhook_pool_expire_token(gemhook_pool *);

/* per-client view for exporters (kubeshare-aggregator style scraping) */
typedef struct gemhook_slot_info {
  char name[64];
  double min_frac, max_frac;
  uint64_t mem_limit, mem_used;
  uint64_t gpu_ns;        /* SM-time reduced on the device, published by the client's hook */
  uint64_t launches;
  uint64_t tokens;        /* tokens granted so far */
  double quota_ms;        /* current adaptive quota */
  double accumulated_ms;  /* ledger sum(end-start) */
  int32_t holds_token, waiting;
} gemhook_slot_info;
int gemhook_pool_slot_info(const gemhook_pool *, int slot, gemhook_slot_info *out);
/* transaction statistics of the lock-free pool: committed transactions, lost publication races (retried), state blocks
 * recycled after their owner died or stalled. */
void gemhook_pool_counters(const gemhook_pool *, uint64_t *commits, uint64_t *conflicts, uint64_t *recycled);

/* process attachment: liveness is an OFD byte-range lock on the pool file, so bytes (and a held token) of a
 * process that died without cleaning up are reclaimed by gemhook_pool_reap() -- the reference does this on
 * socket close (pod-manager.cpp:533-545). mem_reserve/release through an attached handle are tracked per process. */
int gemhook_pool_attach(gemhook_pool *, int slot); /* attachment index or -1 */
void gemhook_pool_detach(gemhook_pool *);
int gemhook_pool_reap(gemhook_pool *);              /* number of dead attachments reclaimed */
/* pod-level token shared by the processes of one pod (gem-pmgr's hook_kernel_launch, pod-manager.cpp:316-473):
 * 1 = must be forwarded to the scheduler with (fwd_overuse, fwd_burst); 0 = answered locally, *remain_ms. */
int gemhook_pool_pod_launch(gemhook_pool *, int slot, int64_t now_us, double overuse_ms, double burst_ms,
                            double *fwd_overuse_ms, double *fwd_burst_ms, double *remain_ms);
double gemhook_pool_pod_granted(gemhook_pool *, int slot, int64_t now_us, double quota_ms);

/* gpu_mem cap on the pool slot (hook.cpp:590-617, pod-manager.cpp:295-313): uint64, requested bytes. */
int gemhook_pool_mem_reserve(gemhook_pool *, int slot, uint64_t bytes);   /* 1 ok, 0 over the cap */
void gemhook_pool_mem_release(gemhook_pool *, int slot, uint64_t bytes);
void gemhook_pool_mem_info(const gemhook_pool *, int slot, uint64_t *used, uint64_t *limit);
/* host address of the slot's shared counter words {mem_used, mem_limit (mirror), gpu_ns, launches}: the words the hook
 * page-locks and device-maps (cuMemHostRegister), i.e. what device code reads when it looks at the pool. */
const void *gemhook_pool_shared_words(const gemhook_pool *, int slot);

/* byte rules for arrays / pitch (hook.cpp:629-680) */
uint64_t gemhook_array_bytes(uint64_t w, uint64_t h, uint64_t d, uint32_t channels, uint32_t format, int is3d);
/* opt-in (GEMHOOK_ACCOUNT_MANAGED=1) charge for a mipmapped array: the rule above per level, extents halving. */
uint64_t gemhook_mipmap_bytes(uint64_t w, uint64_t h, uint64_t d, uint32_t channels, uint32_t format, uint32_t levels);

/* ===================================================================================================
 * (2d) device accounting -- the sm_100a reduction over 16-byte launch records
 * =================================================================================================== */
typedef struct gemhook_record {
  uint32_t slot;       /* client slot; >= nslots: ignored */
  uint32_t launches;   /* launches covered by this record */
  uint64_t elapsed_ns; /* SM-time of the segment */
} gemhook_record;

#define GEMHOOK_MAX_SLOTS 64

typedef struct gemhook_acct gemhook_acct;
/* Binds to the CUDA context current on the calling thread (creates nothing on the CPU side if no
 * device: returns NULL and gemhook_last_error() says why -- there is no CPU fallback). */
gemhook_acct *gemhook_acct_create(uint32_t nslots, size_t ring_capacity_records);
void gemhook_acct_destroy(gemhook_acct *);
/* HOST buffers end to end: copy n records host->device ring, reduce, copy totals back.
 * totals_out: [nslots][3] = elapsed_ns, launches, records (running totals since create/reset). */
int gemhook_acct_reduce_host(gemhook_acct *, const gemhook_record *records, size_t n, uint64_t *totals_out);
/* records already resident in device memory (device pointer); kernel_ms_out (optional) receives the
 * kernel duration measured with CUDA events on the accounting stream. */
int gemhook_acct_reduce_device(gemhook_acct *, uint64_t d_records, size_t n, float *kernel_ms_out);
/* read the mapped pinned totals page (no CUDA call; seqlock reader). */
int gemhook_acct_read_totals(gemhook_acct *, uint64_t *totals_out, uint64_t *epoch_out);
int gemhook_acct_sync(gemhook_acct *);   /* wait for the accounting stream */
int gemhook_acct_reset(gemhook_acct *);  /* zero running totals (stream-ordered) */
uint64_t gemhook_acct_kernel_launches(const gemhook_acct *); /* our own kernels launched so far */
uint64_t gemhook_acct_stream(const gemhook_acct *);           /* CUstream handle of the accounting stream */
/* grid the kernel uses for n records (blocks), for the bench's roofline arithmetic (1 = the one-warp fast path) */
uint32_t gemhook_acct_grid_for(const gemhook_acct *, size_t n);
/* the launch shape chosen for this slot count: out = {warps per block, blocks in a full wave, dynamic shared memory in
 * bytes, TMA buffers per warp (0: register-staged kernel), bin columns, shared memory of the one-warp kernel}. */
void gemhook_acct_launch_shape(const gemhook_acct *, uint32_t out[6]);
/* gpu_mem mirror (north_star b): the authoritative counter is the CAS word in the shared-pinned pool; (used, limit) of the
 * process's pod are handed to every reduce launch, whose publish step leaves them in device memory and in the totals
 * page.  read_mem: from_device = 0 reads the page (no CUDA call), 1 copies the device-resident words back. */
void gemhook_acct_set_mem(gemhook_acct *, uint32_t slot, uint64_t used, uint64_t limit);
int gemhook_acct_read_mem(gemhook_acct *, int from_device, uint64_t *slot, uint64_t *used, uint64_t *limit, uint64_t *epoch);
/* "shared-pinned" check: read four u64 words of host memory (e.g. gemhook_pool_shared_words) THROUGH THE DEVICE -- the
 * page is cuMemHostRegister'ed for the call unless the caller already did -- and return them. */
int gemhook_acct_peek_host_words(gemhook_acct *, const void *host_words, uint64_t out[4]);

/* ===================================================================================================
 * (2e) live hook introspection (the process that has libgemhook.so.1 preloaded)
 * =================================================================================================== */
typedef struct gemhook_stats {
  uint64_t launches;         /* intercepted cuLaunchKernel + cooperative */
  uint64_t fast_path;        /* launches that took the open-gate path */
  uint64_t slow_path;        /* launches at a burst edge */
  uint64_t token_requests;   /* REQ_QUOTA round trips / pool acquisitions */
  uint64_t host_syncs;       /* sync post-hooks seen */
  uint64_t segments;         /* accounting records produced */
  uint64_t acct_kernels;     /* our reduce kernels launched */
  uint64_t gpu_ns;           /* accumulated SM-time (from the totals page / last reduce) */
  uint64_t mem_used;         /* bytes accounted against gpu_mem */
  uint64_t mem_limit;
  uint64_t allocs_denied;
  double quota_ms;           /* current token */
  double overuse_ms;         /* last measured overuse */
  double token_wait_ms;      /* total time blocked waiting for tokens */
  double accumulated_token_ms; /* sum of (quota + overuse) clipped as the ledger does */
} gemhook_stats;
int gemhook_get_stats(gemhook_stats *out);
int gemhook_flush(void); /* resolve pending segments, run the reduce, publish totals */
const char *gemhook_last_error(void);
const char *gemhook_version(void);

#ifdef __cplusplus
}
#endif
#endif
