// gh_internal.h -- declarations shared by the libgemhook.so.1 translation units (not installed).
#ifndef GH_INTERNAL_H
#define GH_INTERNAL_H

#include <cuda.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>
#include <time.h>

#include "../../include/gemhook.h"

#define GH_EXPORT extern "C" __attribute__((visibility("default")))

// ---- logging / errors -----------------------------------------------------------------------------
// GEMHOOK_LOG=0 silent (default) | 1 errors+info | 2 debug.  Never opens a file per call (the reference's
// hINFO/hERROR open+append+close /kubeshare/log/hook.log every time, reference debug.cpp:48-60).
extern int gh_log_level;
void gh_log(int level, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
void gh_set_error(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
#define GH_INFO(...) do { if (gh_log_level >= 1) gh_log(1, __VA_ARGS__); } while (0)
#define GH_DEBUG(...) do { if (gh_log_level >= 2) gh_log(2, __VA_ARGS__); } while (0)

static inline int64_t gh_now_ns(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (int64_t)ts.tv_sec * 1000000000LL + ts.tv_nsec;
}

// ---- the real driver ------------------------------------------------------------------------------
// Every driver entry point the library itself calls, resolved from the real libcuda.so.1 with the libc
// dlsym (never through our own interposers).
#define GH_REAL_DRIVER_FUNCS(X)                                                                        \
  X(cuInit) X(cuDriverGetVersion) X(cuGetErrorString)                                                  \
  X(cuCtxGetCurrent) X(cuCtxSetCurrent) X(cuCtxGetDevice) X(cuCtxSynchronize)                         \
  X(cuDeviceGetAttribute) X(cuDeviceTotalMem_v2) X(cuMemGetInfo_v2)                                    \
  X(cuModuleLoadData) X(cuModuleGetFunction) X(cuModuleUnload) X(cuFuncSetAttribute)                  \
  X(cuOccupancyMaxActiveBlocksPerMultiprocessor)                                                       \
  X(cuStreamCreate) X(cuStreamDestroy_v2) X(cuStreamSynchronize) X(cuStreamIsCapturing)                  \
  X(cuThreadExchangeStreamCaptureMode)                                                                 \
  X(cuEventCreate) X(cuEventDestroy_v2) X(cuEventRecord) X(cuEventSynchronize) X(cuEventElapsedTime)   \
  X(cuEventQuery)                                                                                      \
  X(cuMemAlloc_v2) X(cuMemFree_v2) X(cuMemAllocManaged) X(cuMemAllocPitch_v2)                          \
  X(cuMemcpyHtoDAsync_v2) X(cuMemcpyDtoHAsync_v2) X(cuMemsetD8Async)                                   \
  X(cuMemHostAlloc) X(cuMemFreeHost) X(cuMemHostGetDevicePointer_v2) X(cuMemAllocHost_v2)                 \
  X(cuMemHostRegister_v2) X(cuMemHostUnregister)                                                       \
  X(cuLaunchKernel) X(cuLaunchCooperativeKernel)                                                       \
  X(cuArrayCreate_v2) X(cuArray3DCreate_v2) X(cuArrayDestroy)                                          \
  X(cuMipmappedArrayCreate) X(cuMipmappedArrayDestroy)                                                 \
  X(cuMemcpyAtoH_v2) X(cuMemcpyDtoH_v2) X(cuMemcpyHtoA_v2) X(cuMemcpyHtoD_v2)

struct gh_driver {
#define X(name) void* name;
  GH_REAL_DRIVER_FUNCS(X)
#undef X
  void* gpa_legacy;           // cuGetProcAddress, legacy 4-arg
  void* gpa_v2;               // cuGetProcAddress_v2, 5-arg
  void* handle;               // dlopen handle of the real libcuda
};
extern gh_driver gh_real;
// Resolve the table (idempotent, thread-safe). Returns 0, or -1 when libcuda cannot be opened.
int gh_driver_init(void);
void* gh_true_dlsym(void* handle, const char* symbol);
#define GH_CALL(name, ...) (((decltype(&name))gh_real.name)(__VA_ARGS__))

// ---- predictor / gate (gh_gate.cpp) ---------------------------------------------------------------
struct gh_predictor;
struct gemhook_predictor;
struct gemhook_gate;

// ---- live hook state (gh_hook.cpp) ----------------------------------------------------------------
struct gh_live;
gh_live* gh_live_get(void);  // lazily initialised process singleton (NULL if disabled)

// launch path (called from the interposers in gh_driver.cpp)
void gh_launch_slow(CUstream stream);
void gh_host_sync_pre(void);
void gh_host_sync_post(void);
extern uint32_t gh_gate_open;              // 1: burst ongoing, token valid, segment open (logical gate; relaxed atomics)
#define GH_HIDDEN __attribute__((visibility("hidden")))  /* referenced rip-relative from the assembly fast path */
// Everything the per-launch fast path reads lives in ONE cache line (the driver's own launch path runs ~2 us between two
// launches and evicts L1: measured on a host-bound box, every extra line the hook touched cost ~5 ns = 0.25 %), plus the
// launching thread's own TLS line.
enum { GH_FN_LAUNCH = 0, GH_FN_COOP, GH_FN_LAUNCH_PTSZ, GH_FN_COOP_PTSZ, GH_FN_LAUNCHEX, GH_FN_LAUNCHEX_PTSZ, GH_FN_COUNT };
struct alignas(64) gh_hot_line {
  uint32_t gate_fast;   // +0   gh_gate_open && !CU_HOOK_DEBUG: the ONE word the fast path tests
  uint32_t seg_mask;    // +4   segment every (mask+1) launches; 0xffffffff = burst edges only
  int64_t tls_off;      // +8   offset of gh_tl from the thread pointer (initial-exec TLS: the same in every thread)
  void* fn[GH_FN_COUNT];  // +16  resolved driver entry points of the six kernel-launch hooks
};
static_assert(sizeof(gh_hot_line) == 64, "one cache line");
extern GH_HIDDEN gh_hot_line gh_hot;
#define gh_gate_fast (gh_hot.gate_fast)
#define gh_seg_mask (gh_hot.seg_mask)
// per-thread launch counter: incremented by its own thread with plain moves (no lock prefix, no shared line); readers
// sum the registered threads' counters (slow path only)
struct gh_tl_state {
  uint64_t count;
  uint64_t registered;  // 0 until the thread's first slow path has put &count into the registry
};
extern GH_HIDDEN __thread gh_tl_state gh_tl __attribute__((tls_model("initial-exec"), aligned(16)));
void gh_thread_register(void);             // idempotent; called on this thread's first slow path
uint64_t gh_total_launches(void);          // sum over all threads that ever launched
void gh_gate_set(uint32_t open);           // set the logical gate (and the fast word derived from it)
extern uint32_t gh_hook_debug;            // CU_HOOK_DEBUG=1: count calls per symbol (relaxed atomics; set once by the config load)
void gh_segment_tick(CUstream stream);
void gh_stream_destroyed(CUstream stream);  // cuStreamDestroy pre-hook: a segment open on that stream is closed first

// gpu_mem cap (gh_mem.cpp)
int gh_mem_reserve(uint64_t bytes);        // 1 ok, 0 denied
void gh_mem_commit(uint64_t key, uint64_t bytes);
void gh_mem_unreserve(uint64_t bytes);
void gh_mem_free_key(uint64_t key);
void gh_mem_info(uint64_t* free_b, uint64_t* total_b);

// config (gh_config.cpp)
struct gh_config {
  char pod_name[72];
  char pool_path[512];
  char quota_file[512];
  char scheduler_ip[64];
  int pod_manager_port;
  int transport;          // 0 = tcp (reference daemons), 1 = shared credit pool
  int swap_columns;
  int dry_run;            // no accounting kernels / events (stub driver)
  int extra_hooks;
  int exit_on_failure;    // reference behaviour: exit() when the scheduler is unreachable
  uint32_t seg_launches;  // K
  uint32_t seg_min_us;    // a segment is closed at a sync only when it is at least this old
  uint32_t flush_records;
  double base_quota_ms, min_quota_ms, window_ms;
  int disabled;
  double yield_min_idle_ms;  // ... only when the window predictor expects at least this much idle time
  int yield_on_idle;      // hand the token back at a host sync when another client is waiting (work-conserving option)
  int account_managed;    // GEMHOOK_ACCOUNT_MANAGED=1: charge cuMemAllocManaged / mipmapped arrays (the reference does not)
  int account_host;       // GEMHOOK_ACCOUNT_HOST=1: charge cuMemAllocHost / cuMemHostAlloc (pinned host memory) too
  int hook_debug;         // CU_HOOK_DEBUG=1 (reference hook.cpp:93-100): per-symbol call counters
  char token_trace[512];  // GEMHOOK_TOKEN_TRACE=<path, %d = pid>: one JSON line per token request, written at exit
};
extern gh_config gh_cfg;
void gh_config_load(void);

#endif
