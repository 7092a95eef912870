// tests/native/pool_kill.cpp -- TEST: SIGKILL inside the credit pool's transaction protocol never stalls the survivors.
//
//   pool_kill POOLFILE NCLIENTS SECONDS
//
// The parent creates the pool and forks NCLIENTS worker processes (one slot each) that acquire / use / release tokens
// and reserve memory in a tight loop, plus one monitor process.  Every worker is armed to raise(SIGKILL) at a random
// pass through one of the protocol's fault points (after claiming a state block, in the middle of the policy code,
// between the control-word update and the publication CAS, between the CAS and freeing the superseded block --
// gh_pool.cpp GH_FAULT); the parent additionally SIGKILLs random workers from outside and respawns whatever died.
// The monitor owns its own slot and times every single pool call it makes (request, schedule, slot_info, usage):
// with a lock, a kill inside the critical section stalls everybody until the lock is stolen (round 1: 1 s); with the
// lock-free pool no call may ever wait for another process.  Reported: calls, max / p99.99 wall latency, max thread
// CPU time per call (a spinning waiter burns CPU; a preempted monitor does not), kills by kind, recycled blocks.
// Exit status 0 iff: p99.99 of the wall time per call < 2 ms and its maximum < 100 ms (on this 8-core container 8 busy
// processes share the cores, so a call can be preempted for a scheduler tick or two; round 1's lock made EVERY client
// wait a full second after such a kill), the pool is fully functional afterwards (fresh client gets a token, all dead
// clients' bytes reclaimed).  The thread-CPU maximum is reported too: by construction no call ever waits for another
// process, so what a call costs is its own work plus whatever interrupts the kernel charges to the thread.
#include <fcntl.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <vector>

#include "../../include/gemhook.h"

extern "C" {
int gemhook_fault_point = 0;
long gemhook_fault_countdown = 0;
void gemhook_pool_counters(const gemhook_pool*, uint64_t*, uint64_t*, uint64_t*);
}

static double now_s(clockid_t c) {
  struct timespec ts;
  clock_gettime(c, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static volatile sig_atomic_t g_stop = 0;
static void on_term(int) { g_stop = 1; }

struct Shared {  // anonymous shared page: monitor's results
  double max_wall, max_cpu, p9999;
  long calls, over_1ms;
};

static int worker(const char* path, int slot, unsigned seed) {
  signal(SIGTERM, on_term);
  gemhook_pool* p = gemhook_pool_open(path, 0, 0, 0, 0, 0);
  if (!p) return 3;
  gemhook_pool_reap(p);
  if (gemhook_pool_attach(p, slot) < 0) return 4;
  // arm the suicide: one of the four protocol points, some hundreds of passes from now (or none: killed from outside)
  int kind = rand_r(&seed) % 6;
  if (kind >= 1 && kind <= 4) {
    gemhook_fault_point = kind;
    gemhook_fault_countdown = 50 + rand_r(&seed) % 400;
  }
  while (!g_stop) {
    double q = gemhook_pool_acquire(p, slot, (rand_r(&seed) % 3) * 0.25, (rand_r(&seed) % 20) * 0.5);
    if (q <= 0) return 5;
    uint64_t bytes = 1000 + rand_r(&seed) % 5000;
    if (gemhook_pool_mem_reserve(p, slot, bytes)) {
      usleep(rand_r(&seed) % 300);
      gemhook_pool_mem_release(p, slot, bytes);
    }
    if (rand_r(&seed) % 4 == 0) gemhook_pool_release(p, slot);
  }
  gemhook_pool_release(p, slot);
  gemhook_pool_detach(p);
  gemhook_pool_close(p);
  return 0;
}

static int monitor(const char* path, int slot, double seconds, Shared* out) {
  gemhook_pool* p = gemhook_pool_open(path, 0, 0, 0, 0, 0);
  if (!p) return 3;
  gemhook_pool_attach(p, slot);
  std::vector<float> lat;
  lat.reserve(1 << 22);
  double t_end = now_s(CLOCK_MONOTONIC) + seconds, max_wall = 0, max_cpu = 0;
  long over = 0;
  gemhook_slot_info info;
  int k = 0;
  while (now_s(CLOCK_MONOTONIC) < t_end) {
    double w0 = now_s(CLOCK_MONOTONIC), c0 = now_s(CLOCK_THREAD_CPUTIME_ID);
    switch (k++ % 4) {
      case 0: gemhook_pool_request(p, slot, (w0 - (t_end - seconds)) * 1e3 + 1e6, 0.0, 1.0); break;  // a committing transaction
      case 1: {
        int who;
        double q, sl;
        gemhook_pool_schedule(p, (w0 - (t_end - seconds)) * 1e3 + 1e6, &who, &q, &sl);
        break;
      }
      case 2: gemhook_pool_slot_info(p, slot, &info); break;
      default: gemhook_pool_release(p, slot); break;  // gives the token back if the schedule above granted it to us
    }
    double w = now_s(CLOCK_MONOTONIC) - w0, c = now_s(CLOCK_THREAD_CPUTIME_ID) - c0;
    if (w > max_wall) max_wall = w;
    if (c > max_cpu) max_cpu = c;
    if (w > 1e-3) over++;
    if (lat.size() < lat.capacity()) lat.push_back((float)w);
    if (k % 64 == 0) usleep(50);
  }
  std::sort(lat.begin(), lat.end());
  out->calls = k;
  out->max_wall = max_wall;
  out->max_cpu = max_cpu;
  out->over_1ms = over;
  out->p9999 = lat.empty() ? 0 : lat[(size_t)((lat.size() - 1) * 0.9999)];
  gemhook_pool_release(p, slot);
  gemhook_pool_detach(p);
  gemhook_pool_close(p);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: pool_kill POOLFILE NCLIENTS SECONDS\n");
    return 2;
  }
  const char* path = argv[1];
  int n = atoi(argv[2]);
  double seconds = atof(argv[3]);
  if (n < 1 || n > 32) return 2;
  unlink(path);
  gemhook_pool* p = gemhook_pool_open(path, 1, 2.0, 0.5, 200.0, 0);
  if (!p) {
    fprintf(stderr, "%s\n", gemhook_last_error());
    return 2;
  }
  char cfg[8192];
  int off = snprintf(cfg, sizeof(cfg), "%d\n", n + 1);
  for (int i = 0; i <= n; i++) off += snprintf(cfg + off, sizeof(cfg) - off, "c%d %.4f 1.0 100000\n", i, 1.0 / (n + 1));
  if (gemhook_pool_load_config(p, cfg, 0) != n + 1) return 2;
  Shared* sh = (Shared*)mmap(nullptr, 4096, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  memset(sh, 0, sizeof(*sh));

  std::vector<pid_t> pid(n, 0);
  unsigned seed = 12345;
  auto spawn = [&](int i) {
    unsigned s = rand_r(&seed);
    pid_t c = fork();
    if (c == 0) _exit(worker(path, i, s));
    pid[i] = c;
  };
  for (int i = 0; i < n; i++) spawn(i);
  pid_t mon = fork();
  if (mon == 0) _exit(monitor(path, n, seconds, sh));

  long suicides = 0, murders = 0, errors = 0;
  double t_end = now_s(CLOCK_MONOTONIC) + seconds, next_murder = now_s(CLOCK_MONOTONIC) + 0.05;
  while (now_s(CLOCK_MONOTONIC) < t_end) {
    int st;
    pid_t d = waitpid(-1, &st, WNOHANG);
    if (d > 0) {
      for (int i = 0; i < n; i++)
        if (pid[i] == d) {
          if (WIFSIGNALED(st) && WTERMSIG(st) == SIGKILL) suicides++;
          else errors++;
          spawn(i);
        }
      continue;
    }
    if (now_s(CLOCK_MONOTONIC) > next_murder) {
      int v = rand_r(&seed) % n;
      if (pid[v] > 0) {
        kill(pid[v], SIGKILL);
        murders++;
      }
      next_murder = now_s(CLOCK_MONOTONIC) + 0.02 + (rand_r(&seed) % 30) * 1e-3;
    }
    usleep(500);
  }
  for (int i = 0; i < n; i++)
    if (pid[i] > 0) kill(pid[i], SIGTERM);
  int st;
  while (waitpid(-1, &st, 0) > 0) {
  }
  suicides -= murders < suicides ? murders : suicides;  // both arrive as SIGKILL; what the parent did not send was a suicide

  // afterwards: everything a dead client held is reclaimable and the pool still hands out tokens promptly
  int reaped = gemhook_pool_reap(p);
  uint64_t used_total = 0;
  for (int i = 0; i <= n; i++) {
    uint64_t u, l;
    gemhook_pool_mem_info(p, i, &u, &l);
    used_total += u;
  }
  gemhook_pool_attach(p, 0);
  double t0 = now_s(CLOCK_MONOTONIC);
  double q = gemhook_pool_acquire(p, 0, 0.0, 0.0);
  double fresh_s = now_s(CLOCK_MONOTONIC) - t0;
  gemhook_pool_release(p, 0);
  uint64_t commits = 0, conflicts = 0, recycled = 0;
  gemhook_pool_counters(p, &commits, &conflicts, &recycled);
  bool ok = sh->calls > 1000 && sh->p9999 < 2e-3 && sh->max_wall < 0.1 && used_total == 0 && q > 0 && fresh_s < 0.5 && errors == 0;
  printf("{\"ok\": %s, \"clients\": %d, \"seconds\": %.1f, \"monitor_calls\": %ld, \"max_wall_us\": %.1f, \"p9999_wall_us\": %.1f, "
         "\"max_cpu_us\": %.1f, \"calls_over_1ms\": %ld, \"kills_inside_protocol\": %ld, \"kills_from_outside\": %ld, \"worker_errors\": %ld, "
         "\"reaped_at_end\": %d, \"mem_used_after_reap\": %llu, \"fresh_acquire_ms\": %.3f, \"commits\": %llu, \"conflicts\": %llu, "
         "\"blocks_recycled\": %llu}\n",
         ok ? "true" : "false", n, seconds, sh->calls, sh->max_wall * 1e6, sh->p9999 * 1e6, sh->max_cpu * 1e6, sh->over_1ms, suicides, murders,
         errors, reaped, (unsigned long long)used_total, fresh_s * 1e3, (unsigned long long)commits, (unsigned long long)conflicts,
         (unsigned long long)recycled);
  gemhook_pool_detach(p);
  gemhook_pool_close(p);
  unlink(path);
  return ok ? 0 : 1;
}
