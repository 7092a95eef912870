/*
 * gem-storm -- the synthetic CUDA client of the benchmark and of the LD_PRELOAD tests.
 * Driver API only (the thinnest possible un-hooked launch path, so hook overhead is measured against
 * the hardest baseline).  Workloads follow SURVEY.md 8(d):
 *
 *   storm    W warm-up + K timed steps; a step = B launches of noop<<<1,32>>> on the default stream with
 *            cuCtxSynchronize every S launches                                          (configs 1, 2)
 *   bursty   R rounds of { L ~ U{16..4096} launches of a ~5 us spin kernel; sync; sleep Exp(2 ms) },
 *            seed 0xB200                                                                (config 3)
 *   memsweep cudaMalloc-style sweep: 256 MiB x i cumulative toward --sweep-bytes, then 1000 odd-sized
 *            allocations U[1, 64 MiB], seed 4                                           (config 4)
 *   mnist    iterations of 100 conv launches (N=64, 1->32->64 channels, 28x28, 3x3) + one DtoH (config 5)
 *   truth    R bursts of L self-timing spin kernels (on 1 or 2 streams), sync, optional idle sleep; reports the sum
 *            of the per-burst busy spans measured with %globaltimer INSIDE the kernels  (independent SM-time truth)
 *   graph    capture a CUDA graph on a non-blocking stream while hooked, replay it        (capture safety)
 *   probe    host cost of the primitives the hook builds on (launch, event record, elapsed, stamp)
 *
 * Options shared by the co-residency modes: --start-barrier-dir DIR (all clients wait for each other after context creation,
 * before the first call a hook could intercept); storm, mnist: --track-blocked (rdtsc around every launch; calls over 5 ms
 * are token waits -> "blocked_s", so the client itself reports the time it really ran); storm: --pace-ns N (busy-wait after
 * every launch: how a slower launcher looks to the driver's queue).
 *
 * Output: ONE JSON object on stdout (or --out FILE).  Timing: CUDA events around the timed region on the
 * stream used (device time) AND CLOCK_MONOTONIC around the same region (host wall time).
 */
#define _GNU_SOURCE
#include <cuda.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

extern const unsigned char _binary_storm_kernels_cubin_start[];

#define CK(x)                                                                 \
  do {                                                                        \
    CUresult _r = (x);                                                        \
    if (_r != CUDA_SUCCESS) {                                                 \
      const char* _s = NULL;                                                  \
      cuGetErrorString(_r, &_s);                                              \
      fprintf(stderr, "gem-storm: %s -> %d (%s)\n", #x, (int)_r, _s ? _s : "?"); \
      exit(3);                                                                \
    }                                                                         \
  } while (0)

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static uint64_t rng_state;
static uint32_t rng_u32(void) {  /* splitmix64 */
  uint64_t z = (rng_state += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return (uint32_t)((z ^ (z >> 31)) >> 16);
}
static double rng_unit(void) { return (rng_u32() + 0.5) / 4294967296.0; }

/* file barrier so that co-resident clients enter the timed region together */
static void barrier(const char* dir, int id, int n, const char* tag) {
  if (!dir || n <= 1) return;
  char path[600];
  snprintf(path, sizeof(path), "%s/%s.%d", dir, tag, id);
  FILE* f = fopen(path, "w");
  if (f) fclose(f);
  for (;;) {
    int seen = 0;
    for (int i = 0; i < n; i++) {
      struct stat st;
      snprintf(path, sizeof(path), "%s/%s.%d", dir, tag, i);
      if (stat(path, &st) == 0) seen++;
    }
    if (seen == n) return;
    usleep(500);
  }
}

/* --track-blocked: every launch is timed with rdtsc; a call that takes longer than 5 ms waited for a token (a launch costs
 * ~2 us, draining 1024 noops ~2 ms).  What is left of the client's run time after those calls is the time it really had
 * the GPU -- measured by the client itself, whatever hook is loaded. */
static int track_blocked = 0;
static double ticks_per_ns = 0;
static unsigned long long blocked_thr = 0, blocked_ticks = 0;
static void calibrate_tsc(void) { /* TSC ticks per ns, over 20 ms */
  if (ticks_per_ns > 0) return;
  double a = now_s();
  unsigned long long c0 = __builtin_ia32_rdtsc();
  while (now_s() - a < 0.02) {}
  ticks_per_ns = (double)(__builtin_ia32_rdtsc() - c0) / ((now_s() - a) * 1e9);
  blocked_thr = (unsigned long long)(ticks_per_ns * 5e6);
}
#define TRACKED(call)                                             \
  do {                                                            \
    if (track_blocked) {                                          \
      unsigned long long c_ = __builtin_ia32_rdtsc();             \
      CK(call);                                                   \
      unsigned long long d_ = __builtin_ia32_rdtsc() - c_;        \
      if (d_ > blocked_thr) blocked_ticks += d_;                  \
    } else {                                                      \
      CK(call);                                                   \
    }                                                             \
  } while (0)
static double blocked_seconds(void) { return ticks_per_ns > 0 ? (double)blocked_ticks / ticks_per_ns / 1e9 : 0.0; }

static CUfunction f_noop, f_spin, f_conv, f_spin_stamp;

/* multi-threaded client: every thread launches on its own stream and synchronises now and then */
#include <pthread.h>
struct mt_arg { CUcontext ctx; long launches; int id; };
static void* mt_worker(void* p) {
  struct mt_arg* a = (struct mt_arg*)p;
  CUstream st;
  CK(cuCtxSetCurrent(a->ctx));
  CK(cuStreamCreate(&st, CU_STREAM_DEFAULT));
  for (long i = 1; i <= a->launches; i++) {
    CK(cuLaunchKernel(f_noop, 1, 1, 1, 32, 1, 1, 0, (i & 1) ? st : NULL, NULL, NULL));
    if (i % (97 + a->id * 13) == 0) CK(cuCtxSynchronize());
    if (i % 1000 == 0) {
      CUdeviceptr d;
      if (cuMemAlloc(&d, 4096 + a->id) == CUDA_SUCCESS) cuMemFree(d);
    }
  }
  CK(cuCtxSynchronize());
  return NULL;
}

int main(int argc, char** argv) {
  const char* mode = "storm";
  long step_launches = 65536, sync_every = 1024, steps = 16, warmup = 3, rounds = 2000;
  double pace_ns = 0; /* storm mode: host-side delay added after every launch (how a slower launcher would look) */
  double spin_us = 5.0, sleep_mean_ms = 2.0;
  unsigned long long sweep_bytes = 40ULL << 30;
  const char* out_path = NULL;
  const char* barrier_dir = NULL;
  /* a barrier BEFORE the first call a hook could intercept (after context creation and module load): co-resident clients
   * do not create contexts -- which stalls the device for everybody -- while a peer already runs on a token */
  const char* start_barrier_dir = NULL;
  int client_id = 0, nclients = 1, iters = 20;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--mode") && i + 1 < argc) mode = argv[++i];
    else if (!strcmp(argv[i], "--step-launches") && i + 1 < argc) step_launches = atol(argv[++i]);
    else if (!strcmp(argv[i], "--sync-every") && i + 1 < argc) sync_every = atol(argv[++i]);
    else if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps = atol(argv[++i]);
    else if (!strcmp(argv[i], "--warmup") && i + 1 < argc) warmup = atol(argv[++i]);
    else if (!strcmp(argv[i], "--pace-ns") && i + 1 < argc) pace_ns = atof(argv[++i]);
    else if (!strcmp(argv[i], "--track-blocked")) track_blocked = 1;
    else if (!strcmp(argv[i], "--rounds") && i + 1 < argc) rounds = atol(argv[++i]);
    else if (!strcmp(argv[i], "--spin-us") && i + 1 < argc) spin_us = atof(argv[++i]);
    else if (!strcmp(argv[i], "--sleep-mean-ms") && i + 1 < argc) sleep_mean_ms = atof(argv[++i]);
    else if (!strcmp(argv[i], "--sweep-bytes") && i + 1 < argc) sweep_bytes = strtoull(argv[++i], NULL, 0);
    else if (!strcmp(argv[i], "--out") && i + 1 < argc) out_path = argv[++i];
    else if (!strcmp(argv[i], "--barrier-dir") && i + 1 < argc) barrier_dir = argv[++i];
    else if (!strcmp(argv[i], "--start-barrier-dir") && i + 1 < argc) start_barrier_dir = argv[++i];
    else if (!strcmp(argv[i], "--client-id") && i + 1 < argc) client_id = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--nclients") && i + 1 < argc) nclients = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--iters") && i + 1 < argc) iters = atoi(argv[++i]);
    else {
      fprintf(stderr, "gem-storm: unknown argument %s\n", argv[i]);
      return 2;
    }
  }
  FILE* out = out_path ? fopen(out_path, "w") : stdout;
  if (!out) return 2;

  CUdevice dev;
  CUcontext ctx;
  CUmodule mod;
  CK(cuInit(0));
  CK(cuDeviceGet(&dev, 0));
  CK(cuDevicePrimaryCtxRetain(&ctx, dev));
  CK(cuCtxSetCurrent(ctx));
  CK(cuModuleLoadData(&mod, _binary_storm_kernels_cubin_start));
  CK(cuModuleGetFunction(&f_noop, mod, "noop"));
  CK(cuModuleGetFunction(&f_spin, mod, "spin"));
  CK(cuModuleGetFunction(&f_conv, mod, "conv3x3"));
  CK(cuModuleGetFunction(&f_spin_stamp, mod, "spin_stamp"));
  CUevent e0, e1;
  CK(cuEventCreate(&e0, CU_EVENT_DEFAULT));
  CK(cuEventCreate(&e1, CU_EVENT_DEFAULT));
  barrier(start_barrier_dir, client_id, nclients, "start");

  if (!strcmp(mode, "storm")) {
    /* un-timed warm-up steps, barrier, then exactly K timed steps */
    double t_first = now_s(); /* the hook asks for its first token inside the first launch */
    for (long s = 0; s < warmup; s++) {
      for (long i = 1; i <= step_launches; i++) {
        CK(cuLaunchKernel(f_noop, 1, 1, 1, 32, 1, 1, 0, NULL, NULL, NULL));
        if (i % sync_every == 0) CK(cuCtxSynchronize());
      }
      CK(cuCtxSynchronize());
    }
    barrier(barrier_dir, client_id, nclients, "ready");
    double* step_s = (double*)calloc((size_t)steps + 1, sizeof(double));
    unsigned long long pace_ticks = 0;
    if (pace_ns > 0 || track_blocked) {
      calibrate_tsc();
      pace_ticks = (unsigned long long)(ticks_per_ns * pace_ns);
    }
    TRACKED(cuCtxSynchronize());
    double t0 = now_s();
    CK(cuEventRecord(e0, NULL));
    for (long s = 0; s < steps; s++) {
      double ts = now_s();
      for (long i = 1; i <= step_launches; i++) {
        TRACKED(cuLaunchKernel(f_noop, 1, 1, 1, 32, 1, 1, 0, NULL, NULL, NULL));
        if (pace_ticks) {
          unsigned long long c = __builtin_ia32_rdtsc();
          while (__builtin_ia32_rdtsc() - c < pace_ticks) {}
        }
        if (i % sync_every == 0) CK(cuCtxSynchronize()); /* never a token wait: tokens are asked for at launches */
      }
      CK(cuCtxSynchronize());
      step_s[s] = now_s() - ts;
    }
    CK(cuEventRecord(e1, NULL));
    CK(cuEventSynchronize(e1));
    double t1 = now_s();
    float ev_ms = 0;
    CK(cuEventElapsedTime(&ev_ms, e0, e1));
    fprintf(out,
            "{\"mode\": \"storm\", \"client\": %d, \"launches\": %ld, \"steps\": %ld, \"warmup\": %ld, "
            "\"step_launches\": %ld, \"sync_every\": %ld, \"wall_s\": %.9f, \"event_ms\": %.6f, \"t0\": %.9f, "
            "\"t1\": %.9f, \"t_first\": %.9f, \"t_last\": %.9f, \"blocked_s\": %.9f, \"step_s\": [",
            client_id, steps * step_launches, steps, warmup, step_launches, sync_every, t1 - t0, ev_ms, t0, t1, t_first, t1,
            blocked_seconds());
    for (long s = 0; s < steps; s++) fprintf(out, "%s%.9f", s ? ", " : "", step_s[s]);
    fprintf(out, "]}\n");
  } else if (!strcmp(mode, "bursty")) {
    rng_state = 0xB200ULL + (uint64_t)client_id;
    unsigned long long ns = (unsigned long long)(spin_us * 1000.0);
    void* args[] = {&ns};
    long total = 0;
    barrier(barrier_dir, client_id, nclients, "ready");
    double t0 = now_s();
    CK(cuEventRecord(e0, NULL));
    for (long r = 0; r < rounds; r++) {
      long L = 16 + (long)(rng_u32() % (4096 - 16 + 1));
      for (long i = 0; i < L; i++) CK(cuLaunchKernel(f_spin, 1, 1, 1, 32, 1, 1, 0, NULL, args, NULL));
      total += L;
      CK(cuCtxSynchronize());
      double sl = -log(rng_unit()) * sleep_mean_ms;
      struct timespec ts = {(time_t)(sl / 1e3), (long)(fmod(sl, 1e3) * 1e6)};
      nanosleep(&ts, NULL);
    }
    CK(cuEventRecord(e1, NULL));
    CK(cuEventSynchronize(e1));
    double t1 = now_s();
    float ev_ms = 0;
    CK(cuEventElapsedTime(&ev_ms, e0, e1));
    fprintf(out, "{\"mode\": \"bursty\", \"client\": %d, \"rounds\": %ld, \"launches\": %ld, \"wall_s\": %.9f, \"event_ms\": %.6f}\n",
            client_id, rounds, total, t1 - t0, ev_ms);
  } else if (!strcmp(mode, "memsweep")) {
    /* sweep 1: s_i = 256 MiB * i, cumulative, until the running total would pass sweep_bytes */
    fprintf(out, "{\"mode\": \"memsweep\", \"sweep1\": [");
    unsigned long long total = 0;
    long first_fail = -1;
    CUdeviceptr held[512];
    int nheld = 0;
    for (long i = 1; total < sweep_bytes && i < 512; i++) {
      size_t sz = (size_t)(256ULL << 20) * (size_t)i;
      CUdeviceptr p = 0;
      CUresult r = cuMemAlloc(&p, sz);
      size_t fr = 0, tot = 0;
      cuMemGetInfo(&fr, &tot);
      fprintf(out, "%s{\"i\": %ld, \"bytes\": %zu, \"rc\": %d, \"free\": %zu, \"total\": %zu}", i > 1 ? ", " : "", i, sz, (int)r, fr, tot);
      if (r == CUDA_SUCCESS) {
        held[nheld++] = p;
        total += sz;
      } else {
        if (first_fail < 0) first_fail = i;
        if (r != CUDA_ERROR_OUT_OF_MEMORY) break;
        total += sz; /* keep sweeping "toward 40 GiB" as the config says */
      }
    }
    for (int i = 0; i < nheld; i++) cuMemFree(held[i]);
    size_t fr = 0, tot = 0;
    cuMemGetInfo(&fr, &tot);
    fprintf(out, "], \"first_fail\": %ld, \"free_after_release\": %zu, \"total\": %zu, \"sweep2\": [", first_fail, fr, tot);
    /* sweep 2: 1000 odd-sized allocations, never freed until the end */
    rng_state = 4;
    CUdeviceptr* h2 = (CUdeviceptr*)calloc(1000, sizeof(CUdeviceptr));
    int n2 = 0;
    for (int i = 0; i < 1000; i++) {
      size_t sz = 1 + (size_t)(((uint64_t)rng_u32() << 10 | (rng_u32() & 1023)) % (64ULL << 20));
      CUdeviceptr p = 0;
      CUresult r = cuMemAlloc(&p, sz);
      cuMemGetInfo(&fr, &tot);
      fprintf(out, "%s[%zu, %d, %zu]", i ? ", " : "", sz, (int)r, fr);
      if (r == CUDA_SUCCESS) h2[n2++] = p;
    }
    for (int i = 0; i < n2; i++) cuMemFree(h2[i]);
    cuMemGetInfo(&fr, &tot);
    fprintf(out, "], \"free_end\": %zu}\n", fr);
  } else if (!strcmp(mode, "mnist")) {
    const int N = 64;
    /* the very first intercepted call is a launch, so that every hook flavour asks for its first token at the same
     * point (the reference initialises -- and requests -- inside whichever hooked call comes first, e.g. cuMemAlloc) */
    if (track_blocked) calibrate_tsc();
    double t_first = now_s();
    TRACKED(cuLaunchKernel(f_noop, 1, 1, 1, 32, 1, 1, 0, NULL, NULL, NULL));
    CK(cuCtxSynchronize());
    CUdeviceptr d_in, d_w1, d_a1, d_w2, d_a2;
    CK(cuMemAlloc(&d_in, (size_t)N * 1 * 784 * 4));
    CK(cuMemAlloc(&d_w1, 32 * 1 * 9 * 4));
    CK(cuMemAlloc(&d_a1, (size_t)N * 32 * 784 * 4));
    CK(cuMemAlloc(&d_w2, 64 * 32 * 9 * 4));
    CK(cuMemAlloc(&d_a2, (size_t)N * 64 * 784 * 4));
    CK(cuMemsetD8(d_in, 0, (size_t)N * 784 * 4));
    CK(cuMemsetD8(d_w1, 0, 32 * 9 * 4));
    CK(cuMemsetD8(d_w2, 0, 64 * 32 * 9 * 4));
    float* host = (float*)malloc(64 * 4);
    int c1i = 1, c1o = 32, c2i = 32, c2o = 64;
    void* a1[] = {&d_in, &d_w1, &d_a1, &c1i, &c1o};
    void* a2[] = {&d_a1, &d_w2, &d_a2, &c2i, &c2o};
    /* the blocking copy is reached the way a cudart application reaches it -- through cuGetProcAddress -- because the
     * reference hook exports cuMemcpyDtoH C++-mangled by accident (hook.cpp:925-926) and would not see a direct call */
    typedef CUresult (*dtoh_t)(void*, CUdeviceptr, size_t);
    dtoh_t dtoh = NULL;
    CUdriverProcAddressQueryResult qst;
    CK(cuGetProcAddress("cuMemcpyDtoH", (void**)&dtoh, 12000, CU_GET_PROC_ADDRESS_DEFAULT, &qst));
    barrier(barrier_dir, client_id, nclients, "ready");
    long launches = 0;
    double t0 = now_s();
    CK(cuEventRecord(e0, NULL));
    for (int it = 0; it < iters; it++) {
      for (int k = 0; k < 50; k++) {
        TRACKED(cuLaunchKernel(f_conv, 32, N, 1, 28, 28, 1, 0, NULL, a1, NULL));
        TRACKED(cuLaunchKernel(f_conv, 64, N, 1, 28, 28, 1, 0, NULL, a2, NULL));
        launches += 2;
      }
      CK(dtoh(host, d_a2, 64 * 4));
    }
    CK(cuEventRecord(e1, NULL));
    CK(cuEventSynchronize(e1));
    double t1 = now_s();
    float ev_ms = 0;
    CK(cuEventElapsedTime(&ev_ms, e0, e1));
    fprintf(out, "{\"mode\": \"mnist\", \"client\": %d, \"iters\": %d, \"launches\": %ld, \"wall_s\": %.9f, \"event_ms\": %.6f, "
            "\"t_first\": %.9f, \"t_last\": %.9f, \"blocked_s\": %.9f}\n", client_id, iters, launches, t1 - t0, ev_ms, t_first, t1,
            blocked_seconds());
  } else if (!strcmp(mode, "resolve")) {
    /* the three ways an application reaches the driver: direct symbol, dlsym(), cuGetProcAddress */
    typedef CUresult (*launch_t)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                 CUstream, void**, void**);
    typedef CUresult (*alloc_t)(CUdeviceptr*, size_t);
    typedef CUresult (*free_t)(CUdeviceptr);
    void* h = dlopen("libcuda.so.1", RTLD_NOW);
    launch_t via_dlsym = (launch_t)dlsym(h, "cuLaunchKernel");
    alloc_t alloc_dlsym = (alloc_t)dlsym(h, "cuMemAlloc_v2");
    launch_t via_gpa = NULL;
    alloc_t alloc_gpa = NULL;
    free_t free_gpa = NULL;
    CUdriverProcAddressQueryResult st;
    CK(cuGetProcAddress("cuLaunchKernel", (void**)&via_gpa, 12000, CU_GET_PROC_ADDRESS_DEFAULT, &st));
    CK(cuGetProcAddress("cuMemAlloc", (void**)&alloc_gpa, 12000, CU_GET_PROC_ADDRESS_DEFAULT, &st));
    CK(cuGetProcAddress("cuMemFree", (void**)&free_gpa, 12000, CU_GET_PROC_ADDRESS_DEFAULT, &st));
    void* gpa_again = NULL; /* cudart asks the driver for cuGetProcAddress itself */
    CK(cuGetProcAddress("cuGetProcAddress", &gpa_again, 12000, CU_GET_PROC_ADDRESS_DEFAULT, &st));
    for (int i = 0; i < 10; i++) CK(cuLaunchKernel(f_noop, 1, 1, 1, 32, 1, 1, 0, NULL, NULL, NULL));
    for (int i = 0; i < 10; i++) CK(via_dlsym(f_noop, 1, 1, 1, 32, 1, 1, 0, NULL, NULL, NULL));
    for (int i = 0; i < 10; i++) CK(via_gpa(f_noop, 1, 1, 1, 32, 1, 1, 0, NULL, NULL, NULL));
    /* per-thread-default-stream flavour, as a cudart built with --default-stream per-thread asks for it */
    launch_t via_gpa_pt = NULL;
    CK(cuGetProcAddress("cuLaunchKernel", (void**)&via_gpa_pt, 12000, CU_GET_PROC_ADDRESS_PER_THREAD_DEFAULT_STREAM, &st));
    for (int i = 0; i < 10; i++) CK(via_gpa_pt(f_noop, 1, 1, 1, 32, 1, 1, 0, NULL, NULL, NULL));
    CK(cuCtxSynchronize());
    CUdeviceptr a = 0, b = 0, c = 0;
    CUresult r1 = cuMemAlloc(&a, 1000), r2 = alloc_dlsym(&b, 2000), r3 = alloc_gpa(&c, 3000);
    size_t fr = 0, tot = 0;
    cuMemGetInfo(&fr, &tot);
    fprintf(out, "{\"mode\": \"resolve\", \"rc\": [%d, %d, %d], \"free\": %zu, \"total\": %zu, \"gpa_is_hooked\": %d, \"ptsz_distinct\": %d}\n",
            (int)r1, (int)r2, (int)r3, fr, tot, gpa_again != NULL, via_gpa_pt != via_gpa);
    free_gpa(c);
    cuMemFree(b);
    cuMemFree(a);
  } else if (!strcmp(mode, "arrays")) {
    /* pitch and array allocations: what is charged against gpu_mem (reference hook.cpp:629-680) */
    size_t fr[6] = {0}, tot = 0, pitch = 0;
    CUdeviceptr dp = 0;
    CUarray a2 = NULL, a3 = NULL;
    CUDA_ARRAY_DESCRIPTOR d2;
    CUDA_ARRAY3D_DESCRIPTOR d3;
    memset(&d2, 0, sizeof(d2));
    memset(&d3, 0, sizeof(d3));
    d2.Width = 16; d2.Height = 8; d2.NumChannels = 4; d2.Format = CU_AD_FORMAT_FLOAT;        /* 16*8*4*4 = 2048 B */
    d3.Width = 8; d3.Height = 4; d3.Depth = 2; d3.NumChannels = 2; d3.Format = CU_AD_FORMAT_HALF; /* 8*4*2*2*2 = 256 B */
    cuMemGetInfo(&fr[0], &tot);
    CUresult r1 = cuMemAllocPitch(&dp, &pitch, 100, 7, 4);
    cuMemGetInfo(&fr[1], &tot);
    CUresult r2 = cuArrayCreate(&a2, &d2);
    cuMemGetInfo(&fr[2], &tot);
    CUresult r3 = cuArray3DCreate(&a3, &d3);
    cuMemGetInfo(&fr[3], &tot);
    if (r3 == CUDA_SUCCESS) cuArrayDestroy(a3);
    if (r2 == CUDA_SUCCESS) cuArrayDestroy(a2);
    cuMemGetInfo(&fr[4], &tot);
    if (r1 == CUDA_SUCCESS) cuMemFree(dp);
    cuMemGetInfo(&fr[5], &tot);
    fprintf(out, "{\"mode\": \"arrays\", \"rc\": [%d, %d, %d], \"pitch\": %zu, \"free\": [%zu, %zu, %zu, %zu, %zu, %zu], \"total\": %zu}\n",
            (int)r1, (int)r2, (int)r3, pitch, fr[0], fr[1], fr[2], fr[3], fr[4], fr[5], tot);
  } else if (!strcmp(mode, "optin")) {
    /* allocations the reference does not account (hook.cpp:619-627, 682-694) and pinned host memory: what is charged
     * with GEMHOOK_ACCOUNT_MANAGED / GEMHOOK_ACCOUNT_HOST */
    size_t fr[5] = {0}, tot = 0;
    CUdeviceptr dm = 0;
    CUmipmappedArray mm = NULL;
    void* hp = NULL;
    CUDA_ARRAY3D_DESCRIPTOR d3;
    memset(&d3, 0, sizeof(d3));
    d3.Width = 8; d3.Height = 8; d3.Depth = 0; d3.NumChannels = 1; d3.Format = CU_AD_FORMAT_FLOAT; /* 256 + 64 + 16 B over 3 levels */
    cuMemGetInfo(&fr[0], &tot);
    CUresult r1 = cuMemAllocManaged(&dm, 4096, CU_MEM_ATTACH_GLOBAL);
    cuMemGetInfo(&fr[1], &tot);
    CUresult r2 = cuMipmappedArrayCreate(&mm, &d3, 3);
    cuMemGetInfo(&fr[2], &tot);
    CUresult r3 = cuMemAllocHost(&hp, 2048);
    cuMemGetInfo(&fr[3], &tot);
    if (r3 == CUDA_SUCCESS) cuMemFreeHost(hp);
    if (r2 == CUDA_SUCCESS) cuMipmappedArrayDestroy(mm);
    if (r1 == CUDA_SUCCESS) cuMemFree(dm);
    cuMemGetInfo(&fr[4], &tot);
    fprintf(out, "{\"mode\": \"optin\", \"rc\": [%d, %d, %d], \"free\": [%zu, %zu, %zu, %zu, %zu], \"total\": %zu}\n", (int)r1, (int)r2,
            (int)r3, fr[0], fr[1], fr[2], fr[3], fr[4], tot);
  } else if (!strcmp(mode, "mt")) {
    int T = nclients > 1 ? nclients : 4;  /* --nclients doubles as thread count here */
    pthread_t tid[64];
    struct mt_arg args[64];
    double t0 = now_s();
    for (int i = 0; i < T && i < 64; i++) {
      args[i].ctx = ctx;
      args[i].launches = step_launches;
      args[i].id = i;
      pthread_create(&tid[i], NULL, mt_worker, &args[i]);
    }
    for (int i = 0; i < T && i < 64; i++) pthread_join(tid[i], NULL);
    fprintf(out, "{\"mode\": \"mt\", \"threads\": %d, \"launches\": %ld, \"wall_s\": %.6f}\n", T, (long)T * step_launches, now_s() - t0);
  } else if (!strcmp(mode, "modern")) {
    /* entry points newer than the reference: cuLaunchKernelEx, stream-ordered allocation, cuStreamSynchronize */
    CUstream st;
    CK(cuStreamCreate(&st, CU_STREAM_NON_BLOCKING));
    CUlaunchConfig cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDimX = cfg.gridDimY = cfg.gridDimZ = 1;
    cfg.blockDimX = 32;
    cfg.blockDimY = cfg.blockDimZ = 1;
    cfg.hStream = st;
    for (int b = 0; b < 5; b++) {
      for (int i = 0; i < 20; i++) CK(cuLaunchKernelEx(&cfg, f_noop, NULL, NULL));
      CK(cuStreamSynchronize(st));
    }
    CUdeviceptr a = 0, b2 = 0, c = 0;
    CUresult r1 = cuMemAllocAsync(&a, 1000, st), r2 = cuMemAllocAsync(&b2, 2000, st), r3 = cuMemAllocAsync(&c, 3000, st);
    size_t fr = 0, tot = 0;
    cuMemGetInfo(&fr, &tot);
    if (r2 == CUDA_SUCCESS) cuMemFreeAsync(b2, st);
    CK(cuStreamSynchronize(st));
    size_t fr2 = 0;
    cuMemGetInfo(&fr2, &tot);
    /* virtual memory management: the physical handle is what is charged */
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = 0;
    size_t gran = 0;
    cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
    CUmemGenericAllocationHandle h1 = 0, h2 = 0;
    CUresult v1 = cuMemCreate(&h1, gran ? gran : 4000, &prop, 0);
    size_t fr3 = 0, fr4 = 0;
    cuMemGetInfo(&fr3, &tot);
    CUresult v2 = cuMemCreate(&h2, (gran ? gran : 4000) * 4096, &prop, 0);
    if (v2 == CUDA_SUCCESS) cuMemRelease(h2);
    if (v1 == CUDA_SUCCESS) cuMemRelease(h1);
    cuMemGetInfo(&fr4, &tot);
    fprintf(out, "{\"mode\": \"modern\", \"rc\": [%d, %d, %d], \"free\": %zu, \"free_after\": %zu, \"total\": %zu, "
            "\"vmm\": {\"gran\": %zu, \"rc\": [%d, %d], \"free_held\": %zu, \"free_released\": %zu}}\n",
            (int)r1, (int)r2, (int)r3, fr, fr2, tot, gran, (int)v1, (int)v2, fr3, fr4);
    if (r1 == CUDA_SUCCESS) cuMemFreeAsync(a, st);
    CK(cuStreamSynchronize(st));
  } else if (!strcmp(mode, "truth")) {
    /* --rounds R bursts; --step-launches L kernels per burst; --spin-us; --sleep-mean-ms = FIXED idle after each sync
     * (0 = relaunch at once); --nclients doubles as the number of streams (1 = legacy default stream, 2 = two
     * non-blocking streams used alternately) */
    int nstreams = nclients > 1 ? 2 : 0;
    CUstream st[2] = {NULL, NULL};
    for (int i = 0; i < nstreams; i++) CK(cuStreamCreate(&st[i], CU_STREAM_NON_BLOCKING));
    CUdeviceptr d_bounds;
    size_t bb = (size_t)rounds * 2 * sizeof(unsigned long long);
    unsigned long long* h_bounds = (unsigned long long*)malloc(bb);
    for (long r = 0; r < rounds; r++) {
      h_bounds[2 * r] = ~0ULL;
      h_bounds[2 * r + 1] = 0ULL;
    }
    CK(cuMemAlloc(&d_bounds, bb));
    CK(cuMemcpyHtoD(d_bounds, h_bounds, bb));
    unsigned long long ns = (unsigned long long)(spin_us * 1000.0);
    double t0 = now_s();
    for (long r = 0; r < rounds; r++) {
      unsigned burst = (unsigned)r;
      void* args[] = {&ns, &d_bounds, &burst};
      for (long i = 0; i < step_launches; i++)
        CK(cuLaunchKernel(f_spin_stamp, 1, 1, 1, 32, 1, 1, 0, nstreams ? st[i & 1] : NULL, args, NULL));
      CK(cuCtxSynchronize());
      if (sleep_mean_ms > 0) {
        struct timespec ts = {(time_t)(sleep_mean_ms / 1e3), (long)(fmod(sleep_mean_ms, 1e3) * 1e6)};
        nanosleep(&ts, NULL);
      }
    }
    double t1 = now_s();
    CK(cuMemcpyDtoH(h_bounds, d_bounds, bb));
    unsigned long long truth = 0;
    for (long r = 0; r < rounds; r++)
      if (h_bounds[2 * r + 1] > h_bounds[2 * r]) truth += h_bounds[2 * r + 1] - h_bounds[2 * r];
    fprintf(out, "{\"mode\": \"truth\", \"rounds\": %ld, \"launches\": %ld, \"streams\": %d, \"spin_us\": %.3f, "
            "\"idle_ms\": %.3f, \"truth_ns\": %llu, \"wall_s\": %.9f}\n",
            rounds, rounds * step_launches, nstreams ? 2 : 1, spin_us, sleep_mean_ms, truth, t1 - t0);
  } else if (!strcmp(mode, "graph")) {
    /* capture `step_launches` kernels into a graph on a non-blocking stream (short tokens may expire meanwhile),
     * instantiate, replay `rounds` times; every API result is reported */
    CUstream cs;
    CK(cuStreamCreate(&cs, CU_STREAM_NON_BLOCKING));
    for (int i = 0; i < 64; i++) CK(cuLaunchKernel(f_noop, 1, 1, 1, 32, 1, 1, 0, cs, NULL, NULL)); /* open a segment on cs */
    CUgraph g = NULL;
    CUgraphExec ge = NULL;
    unsigned long long ns = (unsigned long long)(spin_us * 1000.0);
    void* args[] = {&ns};
    CUresult rb = cuStreamBeginCapture(cs, CU_STREAM_CAPTURE_MODE_GLOBAL);
    CUresult rl = CUDA_SUCCESS;
    for (long i = 0; i < step_launches && rl == CUDA_SUCCESS; i++) {
      rl = cuLaunchKernel(f_spin, 1, 1, 1, 32, 1, 1, 0, cs, args, NULL);
      if (i % 16 == 15) usleep(2000); /* let short tokens expire during the capture */
    }
    CUresult re = cuStreamEndCapture(cs, &g);
    size_t nodes = 0;
    if (re == CUDA_SUCCESS && g) cuGraphGetNodes(g, NULL, &nodes);
    CUresult ri = (re == CUDA_SUCCESS && g) ? cuGraphInstantiateWithFlags(&ge, g, 0) : CUDA_ERROR_INVALID_VALUE;
    CUresult rr = CUDA_SUCCESS;
    double replay_ms = 0; /* the application's own measurement of its replays: an event pair around each */
    for (long r = 0; r < rounds && ri == CUDA_SUCCESS && rr == CUDA_SUCCESS; r++) {
      cuEventRecord(e0, cs);
      rr = cuGraphLaunch(ge, cs);
      cuEventRecord(e1, cs);
      if (rr == CUDA_SUCCESS) rr = cuStreamSynchronize(cs);
      float ms = 0;
      if (rr == CUDA_SUCCESS && cuEventElapsedTime(&ms, e0, e1) == CUDA_SUCCESS) replay_ms += ms;
    }
    CUresult rs = cuCtxSynchronize();
    CUresult rd = cuStreamDestroy(cs);
    for (int i = 0; i < 64; i++) CK(cuLaunchKernel(f_noop, 1, 1, 1, 32, 1, 1, 0, NULL, NULL, NULL));
    CK(cuCtxSynchronize());
    fprintf(out, "{\"mode\": \"graph\", \"begin\": %d, \"launch\": %d, \"end\": %d, \"nodes\": %zu, \"instantiate\": %d, "
            "\"replay\": %d, \"sync\": %d, \"destroy\": %d, \"captured\": %ld, \"replays\": %ld, \"replay_ms\": %.6f}\n",
            (int)rb, (int)rl, (int)re, nodes, (int)ri, (int)rr, (int)rs, (int)rd, step_launches, rounds, replay_ms);
  } else if (!strcmp(mode, "probe")) {
    /* host cost (ns) of the building blocks; medians would be nicer, means over 20k are stable enough */
    const int N = 20000;
    for (int i = 0; i < 2000; i++) CK(cuLaunchKernel(f_noop, 1, 1, 1, 32, 1, 1, 0, NULL, NULL, NULL));
    CK(cuCtxSynchronize());
    double t = now_s();
    for (int i = 0; i < N; i++) {
      CK(cuLaunchKernel(f_noop, 1, 1, 1, 32, 1, 1, 0, NULL, NULL, NULL));
      if ((i & 1023) == 1023) CK(cuCtxSynchronize());
    }
    CK(cuCtxSynchronize());
    double launch_ns = (now_s() - t) / N * 1e9;
    CUevent ev[64];
    for (int i = 0; i < 64; i++) CK(cuEventCreate(&ev[i], CU_EVENT_DEFAULT));
    t = now_s();
    for (int i = 0; i < N; i++) CK(cuEventRecord(ev[i & 63], NULL));
    double rec_ns = (now_s() - t) / N * 1e9;
    CK(cuCtxSynchronize());
    float ms;
    t = now_s();
    for (int i = 0; i < N; i++) cuEventElapsedTime(&ms, ev[i & 31], ev[32 + (i & 31)]);
    double el_ns = (now_s() - t) / N * 1e9;
    t = now_s();
    for (int i = 0; i < N; i++) cuEventQuery(ev[i & 63]);
    double q_ns = (now_s() - t) / N * 1e9;
    t = now_s();
    for (int i = 0; i < 200; i++) CK(cuCtxSynchronize());
    double sync_ns = (now_s() - t) / 200 * 1e9;
    CUstream ps;
    CK(cuStreamCreate(&ps, CU_STREAM_NON_BLOCKING));
    CUstreamCaptureStatus cst;
    t = now_s();
    for (int i = 0; i < N; i++) cuStreamIsCapturing(NULL, &cst);
    double cap0_ns = (now_s() - t) / N * 1e9;
    t = now_s();
    for (int i = 0; i < N; i++) cuStreamIsCapturing(ps, &cst);
    double cap1_ns = (now_s() - t) / N * 1e9;
    struct timespec ts;
    t = now_s();
    for (int i = 0; i < 1000000; i++) clock_gettime(CLOCK_MONOTONIC, &ts);
    double clk_ns = (now_s() - t) / 1e6 * 1e9;
    fprintf(out,
            "{\"mode\": \"probe\", \"launch_ns\": %.1f, \"event_record_ns\": %.1f, \"event_elapsed_ns\": %.1f, "
            "\"event_query_ns\": %.1f, \"idle_ctx_sync_ns\": %.1f, \"clock_gettime_ns\": %.1f, \"is_capturing_legacy_ns\": %.1f, "
            "\"is_capturing_stream_ns\": %.1f, \"cpus\": %ld}\n",
            launch_ns, rec_ns, el_ns, q_ns, sync_ns, clk_ns, cap0_ns, cap1_ns, sysconf(_SC_NPROCESSORS_ONLN));
  } else {
    fprintf(stderr, "gem-storm: unknown mode %s\n", mode);
    return 2;
  }
  if (out != stdout) fclose(out);
  return 0;
}
