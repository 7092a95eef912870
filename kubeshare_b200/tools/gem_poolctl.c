/*
 * gem-poolctl -- inspect / administer a gemhook credit pool (usage export).
 *   gem-poolctl POOL dump            JSON: one object per client
 *   gem-poolctl POOL prom            Prometheus text exposition (gauge names mirror kubeshare-aggregator's
 *                                    gpu_requirement labels: reference pkg/aggregator/aggregator.go:22-38)
 *   gem-poolctl POOL ledger          the token ledger in the shape gem-schd's _DEBUG build dumps on SIGINT
 *                                    (reference scheduler.cpp:693-714): [{"container", "start", "end"}] in seconds
 *   gem-poolctl POOL load FILE [limit_request]   (re)load a quota file into the pool
 *   gem-poolctl POOL reap            reclaim bytes / token of dead clients
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/gemhook.h"

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: gem-poolctl POOL dump|prom|ledger|load FILE [limit_request]|reap\n");
    return 2;
  }
  gemhook_pool* p = gemhook_pool_open(argv[1], !strcmp(argv[2], "load"), 300.0, 20.0, 10000.0, 0);
  if (!p) {
    fprintf(stderr, "gem-poolctl: %s\n", gemhook_last_error());
    return 1;
  }
  int n = gemhook_pool_nslots(p);
  gemhook_slot_info s;
  if (!strcmp(argv[2], "dump")) {
    printf("[");
    for (int i = 0; i < n; i++) {
      gemhook_pool_slot_info(p, i, &s);
      printf("%s\n {\"pod\": \"%s\", \"request\": %.17g, \"limit\": %.17g, \"mem_limit\": %llu, \"mem_used\": %llu, "
             "\"gpu_ms\": %.6f, \"launches\": %llu, \"tokens\": %llu, \"quota_ms\": %.6f, \"accumulated_token_ms\": %.6f, "
             "\"holds_token\": %d, \"waiting\": %d}",
             i ? "," : "", s.name, s.min_frac, s.max_frac, (unsigned long long)s.mem_limit, (unsigned long long)s.mem_used,
             s.gpu_ns / 1e6, (unsigned long long)s.launches, (unsigned long long)s.tokens, s.quota_ms, s.accumulated_ms,
             s.holds_token, s.waiting);
    }
    printf("\n]\n");
  } else if (!strcmp(argv[2], "prom")) {
    printf("# TYPE gemhook_gpu_seconds_total counter\n# TYPE gemhook_token_seconds_total counter\n"
           "# TYPE gemhook_launches_total counter\n# TYPE gemhook_mem_used_bytes gauge\n# TYPE gemhook_mem_limit_bytes gauge\n");
    for (int i = 0; i < n; i++) {
      gemhook_pool_slot_info(p, i, &s);
      printf("gemhook_gpu_seconds_total{pod=\"%s\"} %.9f\n", s.name, s.gpu_ns / 1e9);
      printf("gemhook_token_seconds_total{pod=\"%s\"} %.6f\n", s.name, s.accumulated_ms / 1e3);
      printf("gemhook_launches_total{pod=\"%s\"} %llu\n", s.name, (unsigned long long)s.launches);
      printf("gemhook_mem_used_bytes{pod=\"%s\",request=\"%g\",limit=\"%g\"} %llu\n", s.name, s.min_frac, s.max_frac,
             (unsigned long long)s.mem_used);
      printf("gemhook_mem_limit_bytes{pod=\"%s\"} %llu\n", s.name, (unsigned long long)s.mem_limit);
    }
  } else if (!strcmp(argv[2], "ledger")) {
    size_t k = gemhook_pool_history(p, NULL, NULL, NULL, 0);
    int* sl = (int*)calloc(k + 1, sizeof(int));
    double* a = (double*)calloc(k + 1, sizeof(double));
    double* b = (double*)calloc(k + 1, sizeof(double));
    size_t got = gemhook_pool_history(p, sl, a, b, k);
    if (got < k) k = got;
    printf("[\n");
    for (size_t i = 0; i < k; i++) {
      gemhook_pool_slot_info(p, sl[i], &s);
      printf("\t{\"container\": \"%s\", \"start\": %.3lf, \"end\" : %.3lf}%s\n", s.name, a[i] / 1000.0, b[i] / 1000.0, i + 1 < k ? "," : "");
    }
    printf("]\n");
    free(sl);
    free(a);
    free(b);
  } else if (!strcmp(argv[2], "load") && argc >= 4) {
    FILE* f = fopen(argv[3], "r");
    if (!f) {
      perror(argv[3]);
      return 1;
    }
    static char text[1 << 16];
    size_t k = fread(text, 1, sizeof(text) - 1, f);
    text[k] = 0;
    fclose(f);
    int swap = argc >= 5 && !strcmp(argv[4], "limit_request");
    int c = gemhook_pool_load_config(p, text, swap);
    printf("%d\n", c);
    if (c < 0) return 1;
  } else if (!strcmp(argv[2], "reap")) {
    printf("%d\n", gemhook_pool_reap(p));
  } else {
    fprintf(stderr, "gem-poolctl: unknown command\n");
    return 2;
  }
  gemhook_pool_close(p);
  return 0;
}
