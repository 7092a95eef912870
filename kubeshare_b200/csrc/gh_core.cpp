// gh_core.cpp -- logging, error string, configuration, and resolution of the REAL driver.
#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "gh_internal.h"

int gh_log_level = 0;
gh_driver gh_real;
gh_config gh_cfg;
uint32_t gh_hook_debug = 0;

static __thread char tls_error[512];
static char g_error[512];

void gh_log(int level, const char* fmt, ...) {
  char buf[600];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  fprintf(stderr, "[gemhook %d %s] %s\n", (int)getpid(), level >= 2 ? "dbg" : "inf", buf);
}

void gh_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tls_error, sizeof(tls_error), fmt, ap);
  va_end(ap);
  memcpy(g_error, tls_error, sizeof(g_error));
  if (gh_log_level >= 1) gh_log(1, "error: %s", tls_error);
}

GH_EXPORT const char* gemhook_last_error(void) { return tls_error[0] ? tls_error : g_error; }
GH_EXPORT const char* gemhook_version(void) { return "gemhook-b200 0.1 (abi 1, sm_100a)"; }

// ---- the libc dlsym, reached without going through our own interposer ---------------------------------
// glibc >= 2.34 no longer exports __libc_dlsym (what the reference used, hook.cpp:66-84); dlvsym with
// the versioned name is the supported way to get the real dlsym from inside a dlsym interposer.
typedef void* (*dlsym_fn)(void*, const char*);
#define GH_NO_SANITIZE __attribute__((no_sanitize("thread", "address", "undefined")))
// (sanitizer runtimes call dlsym() while they initialise: these two must not be instrumented)
GH_NO_SANITIZE static dlsym_fn real_dlsym_ptr(void) {
  static dlsym_fn fn = nullptr;
  if (!fn) {
    fn = (dlsym_fn)dlvsym(RTLD_NEXT, "dlsym", "GLIBC_2.2.5");
    if (!fn) fn = (dlsym_fn)dlvsym(RTLD_NEXT, "dlsym", "GLIBC_2.34");
  }
  return fn;
}
GH_NO_SANITIZE void* gh_true_dlsym(void* handle, const char* symbol) {
  dlsym_fn fn = real_dlsym_ptr();
  return fn ? fn(handle, symbol) : nullptr;
}

static pthread_once_t drv_once = PTHREAD_ONCE_INIT;
static int drv_rc = -1;

static void driver_init_once(void) {
  const char* lvl = getenv("GEMHOOK_LOG");
  if (lvl) gh_log_level = atoi(lvl);
  const char* path = getenv("GEMHOOK_LIBCUDA");
  void* h = nullptr;
  if (path && *path) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    gh_set_error("cannot open the CUDA driver (libcuda.so.1): %s", dlerror());
    return;
  }
  gh_real.handle = h;
#define X(name) gh_real.name = gh_true_dlsym(h, #name);
  GH_REAL_DRIVER_FUNCS(X)
#undef X
  gh_real.gpa_legacy = gh_true_dlsym(h, "cuGetProcAddress");
  gh_real.gpa_v2 = gh_true_dlsym(h, "cuGetProcAddress_v2");
  if (!gh_real.cuLaunchKernel || !gh_real.cuMemAlloc_v2) {
    gh_set_error("the CUDA driver lacks cuLaunchKernel/cuMemAlloc_v2");
    return;
  }
  drv_rc = 0;
}

int gh_driver_init(void) {
  pthread_once(&drv_once, driver_init_once);
  return drv_rc;
}

// ---- configuration -------------------------------------------------------------------------------------
// Same inputs as the reference hook (reference hook.cpp:227-256, comm.cpp:33-40): POD_NAME (fallback
// hostname), POD_MANAGER_PORT (default 50052), /kubeshare/library/schedulerIP.txt, CU_HOOK_DEBUG; plus
// GEMHOOK_* overrides used by the dry run, the tests and the pool transport.
static void env_str(const char* key, char* dst, size_t cap, const char* dflt) {
  const char* v = getenv(key);
  snprintf(dst, cap, "%s", (v && *v) ? v : dflt);
}
static double env_f(const char* key, double dflt) {
  const char* v = getenv(key);
  return (v && *v) ? atof(v) : dflt;
}
static long env_i(const char* key, long dflt) {
  const char* v = getenv(key);
  return (v && *v) ? strtol(v, nullptr, 0) : dflt;
}

void gh_config_load(void) {
  gh_config& c = gh_cfg;
  memset(&c, 0, sizeof(c));
  const char* lvl = getenv("GEMHOOK_LOG");
  if (lvl) gh_log_level = atoi(lvl);
  const char* dbg = getenv("CU_HOOK_DEBUG");  // reference hook.cpp:96
  if (dbg && dbg[0] == '1' && gh_log_level < 2) gh_log_level = 2;
  c.hook_debug = (dbg && dbg[0] == '1') ? 1 : 0;
  __atomic_store_n(&gh_hook_debug, (uint32_t)c.hook_debug, __ATOMIC_RELAXED);

  const char* pn = getenv("POD_NAME");
  if (pn && *pn) snprintf(c.pod_name, sizeof(c.pod_name), "%s", pn);
  else gethostname(c.pod_name, sizeof(c.pod_name) - 1);

  c.pod_manager_port = (int)env_i("POD_MANAGER_PORT", 50052);
  env_str("GEMHOOK_SCHEDULER_IP", c.scheduler_ip, sizeof(c.scheduler_ip), "");
  if (!c.scheduler_ip[0]) {
    char ipfile[512];
    env_str("GEMHOOK_SCHEDULER_IP_FILE", ipfile, sizeof(ipfile), "/kubeshare/library/schedulerIP.txt");
    FILE* f = fopen(ipfile, "r");
    if (f) {
      if (fgets(c.scheduler_ip, sizeof(c.scheduler_ip), f)) c.scheduler_ip[strcspn(c.scheduler_ip, "\r\n")] = 0;
      fclose(f);
    }
  }
  env_str("GEMHOOK_POOL", c.pool_path, sizeof(c.pool_path), "");
  env_str("GEMHOOK_QUOTA_FILE", c.quota_file, sizeof(c.quota_file), "");
  const char* tr = getenv("GEMHOOK_TRANSPORT");
  if (tr && !strcmp(tr, "tcp")) c.transport = 0;
  else if (tr && !strcmp(tr, "pool")) c.transport = 1;
  else c.transport = c.pool_path[0] ? 1 : 0;
  const char* cols = getenv("GEMHOOK_QUOTA_COLUMNS");  // SURVEY.md 8b column-order trap
  c.swap_columns = (cols && !strcmp(cols, "limit_request")) ? 1 : 0;
  c.dry_run = (int)env_i("GEMHOOK_DRY_RUN", 0);
  c.extra_hooks = (int)env_i("GEMHOOK_EXTRA_HOOKS", 0);
  c.exit_on_failure = (int)env_i("GEMHOOK_EXIT_ON_FAILURE", 1);
  c.seg_launches = (uint32_t)env_i("GEMHOOK_SEG_LAUNCHES", 0);  // 0 = burst edges only
  c.seg_min_us = (uint32_t)env_i("GEMHOOK_SEG_MIN_US", 4000);
  c.flush_records = (uint32_t)env_i("GEMHOOK_FLUSH_RECORDS", 64);
  c.base_quota_ms = env_f("GEMHOOK_BASE_QUOTA_MS", 300.0);  // reference launcher.py:77-80
  c.min_quota_ms = env_f("GEMHOOK_MIN_QUOTA_MS", 20.0);
  c.window_ms = env_f("GEMHOOK_WINDOW_MS", 10000.0);
  c.disabled = (int)env_i("GEMHOOK_DISABLE", 0);
  c.yield_on_idle = (int)env_i("GEMHOOK_YIELD_ON_IDLE", 0);
  c.yield_min_idle_ms = env_f("GEMHOOK_YIELD_MIN_IDLE_MS", 0.5);
  c.account_managed = (int)env_i("GEMHOOK_ACCOUNT_MANAGED", 0);
  c.account_host = (int)env_i("GEMHOOK_ACCOUNT_HOST", 0);
  env_str("GEMHOOK_TOKEN_TRACE", c.token_trace, sizeof(c.token_trace), "");
}
