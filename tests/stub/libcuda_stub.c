/*
 * tests/stub/libcuda_stub.c -- TEST INFRASTRUCTURE: a fake libcuda.so.1 for the CPU-only dry run
 * (BASELINE.json configs[0]).  It exports the driver symbols the hook and the
 * storm client use, counts calls, and models a GPU as a single timeline: each launch occupies
 * STUB_KERNEL_US microseconds (default 2) after max(host now, previous work); events capture that
 * timeline; synchronising calls sleep until the host clock catches up.  No kernel is ever executed.
 *
 * STUB_REPORT=<file>: JSON with the call counters at process exit.
 * STUB_TOTAL_MEM=<bytes>: what the REAL cuMemGetInfo/cuDeviceTotalMem would say (default 180 GiB).
 */
#define _GNU_SOURCE
#include <cuda.h>
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#undef cuGetProcAddress

static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
static int64_t gpu_free_at_ns;  /* the timeline: when already-issued work completes */
static uint64_t n_launch, n_coop, n_alloc, n_free, n_sync, n_event_record, n_memcpy, n_module, n_getproc;
static uint64_t n_event_on_capturing, n_stream_destroy;
static uint64_t bytes_live, next_addr = 0x700000000000ULL;
static int fake_ctx_obj, fake_mod_obj;
static __thread CUcontext cur_ctx;

static int64_t now_ns(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (int64_t)ts.tv_sec * 1000000000LL + ts.tv_nsec;
}
static int64_t kernel_ns(void) {
  static int64_t v = -1;
  if (v < 0) {
    const char* e = getenv("STUB_KERNEL_US");
    v = (int64_t)((e ? atof(e) : 2.0) * 1000.0);
  }
  return v;
}
static void wait_until(int64_t t) {
  int64_t n;
  while ((n = now_ns()) < t) {
    int64_t d = t - n;
    if (d > 200000) {
      struct timespec ts = {0, (long)(d - 100000)};
      nanosleep(&ts, NULL);
    }
  }
}
static void enqueue(int64_t dur) {
  pthread_mutex_lock(&mu);
  int64_t n = now_ns();
  if (gpu_free_at_ns < n) gpu_free_at_ns = n;
  gpu_free_at_ns += dur;
  pthread_mutex_unlock(&mu);
}
static int64_t timeline(void) {
  pthread_mutex_lock(&mu);
  int64_t n = now_ns();
  int64_t t = gpu_free_at_ns < n ? n : gpu_free_at_ns;
  pthread_mutex_unlock(&mu);
  return t;
}

static void report(void) {
  const char* path = getenv("STUB_REPORT");
  if (!path) return;
  FILE* f = fopen(path, "w");
  if (!f) return;
  fprintf(f,
          "{\"launches\": %llu, \"coop_launches\": %llu, \"allocs\": %llu, \"frees\": %llu, \"syncs\": %llu, "
          "\"event_records\": %llu, \"memcpys\": %llu, \"modules\": %llu, \"getproc\": %llu, \"bytes_live\": %llu, "
          "\"events_on_capturing_streams\": %llu, \"stream_destroys\": %llu}\n",
          (unsigned long long)n_launch, (unsigned long long)n_coop, (unsigned long long)n_alloc,
          (unsigned long long)n_free, (unsigned long long)n_sync, (unsigned long long)n_event_record,
          (unsigned long long)n_memcpy, (unsigned long long)n_module, (unsigned long long)n_getproc,
          (unsigned long long)bytes_live, (unsigned long long)n_event_on_capturing, (unsigned long long)n_stream_destroy);
  fclose(f);
}

CUresult cuInit(unsigned int flags) {
  static int once;
  if (!once) {
    once = 1;
    atexit(report);
  }
  return CUDA_SUCCESS;
}
CUresult cuDriverGetVersion(int* v) { *v = 12090; return CUDA_SUCCESS; }
CUresult cuGetErrorString(CUresult e, const char** s) { *s = e == CUDA_SUCCESS ? "no error" : "stub error"; return CUDA_SUCCESS; }
CUresult cuGetErrorName(CUresult e, const char** s) { *s = e == CUDA_SUCCESS ? "CUDA_SUCCESS" : "CUDA_ERROR_STUB"; return CUDA_SUCCESS; }
CUresult cuDeviceGetCount(int* n) { *n = 1; return CUDA_SUCCESS; }
CUresult cuDeviceGet(CUdevice* d, int ord) { *d = 0; return ord == 0 ? CUDA_SUCCESS : CUDA_ERROR_INVALID_DEVICE; }
CUresult cuDeviceGetName(char* name, int len, CUdevice d) { snprintf(name, len, "STUB B200"); return CUDA_SUCCESS; }
CUresult cuDeviceGetAttribute(int* pi, CUdevice_attribute a, CUdevice d) {
  switch (a) {
    case CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR: *pi = 10; break;
    case CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR: *pi = 0; break;
    case CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT: *pi = 148; break;
    default: *pi = 0; break;
  }
  return CUDA_SUCCESS;
}
static size_t total_mem(void) {
  const char* e = getenv("STUB_TOTAL_MEM");
  return e ? (size_t)strtoull(e, NULL, 0) : (size_t)180ULL << 30;
}
CUresult cuDeviceTotalMem_v2(size_t* b, CUdevice d) { *b = total_mem(); return CUDA_SUCCESS; }
CUresult cuMemGetInfo_v2(size_t* f, size_t* t) { *t = total_mem(); *f = total_mem() - bytes_live; return CUDA_SUCCESS; }

CUresult cuCtxCreate_v2(CUcontext* c, unsigned int fl, CUdevice d) { *c = (CUcontext)&fake_ctx_obj; cur_ctx = *c; return CUDA_SUCCESS; }
CUresult cuDevicePrimaryCtxRetain(CUcontext* c, CUdevice d) { *c = (CUcontext)&fake_ctx_obj; return CUDA_SUCCESS; }
CUresult cuDevicePrimaryCtxRelease_v2(CUdevice d) { return CUDA_SUCCESS; }
CUresult cuCtxDestroy_v2(CUcontext c) { return CUDA_SUCCESS; }
CUresult cuCtxSetCurrent(CUcontext c) { cur_ctx = c; return CUDA_SUCCESS; }
CUresult cuCtxGetCurrent(CUcontext* c) { *c = cur_ctx; return CUDA_SUCCESS; }
CUresult cuCtxGetDevice(CUdevice* d) { *d = 0; return CUDA_SUCCESS; }
CUresult cuCtxSynchronize(void) { n_sync++; wait_until(timeline()); return CUDA_SUCCESS; }

CUresult cuModuleLoadData(CUmodule* m, const void* image) {
  /* STUB_MODULE_LOAD_US: the second and later loads (the first is the application's own module; the next one is the
   * hook's cubin, loaded during its first-use initialisation) take this long -- a real driver needs milliseconds */
  const char* e = getenv("STUB_MODULE_LOAD_US");
  if (e && n_module >= 1) {
    long us = atol(e);
    struct timespec ts = {us / 1000000, (us % 1000000) * 1000};
    nanosleep(&ts, NULL);
  }
  n_module++;
  *m = (CUmodule)&fake_mod_obj;
  return CUDA_SUCCESS;
}
CUresult cuModuleUnload(CUmodule m) { return CUDA_SUCCESS; }
CUresult cuModuleGetFunction(CUfunction* f, CUmodule m, const char* name) {
  *f = (CUfunction)(uintptr_t)(0x1000 + (uintptr_t)strlen(name));
  return CUDA_SUCCESS;
}
CUresult cuFuncSetAttribute(CUfunction f, CUfunction_attribute a, int v) { return CUDA_SUCCESS; }
CUresult cuOccupancyMaxActiveBlocksPerMultiprocessor(int* n, CUfunction f, int bs, size_t smem) { *n = 2; return CUDA_SUCCESS; }

CUresult cuStreamCreate(CUstream* s, unsigned int fl) { *s = (CUstream)calloc(1, 8); return CUDA_SUCCESS; }
CUresult cuStreamDestroy_v2(CUstream s) { n_stream_destroy++; free(s); return CUDA_SUCCESS; }
CUresult cuStreamSynchronize(CUstream s) { wait_until(timeline()); return CUDA_SUCCESS; }
/* stream capture: a flag in the fake stream object; events recorded on a capturing stream are counted (the hook must
 * keep its events off such a stream) */
static int stub_capturing(CUstream s) { return s && *(int*)s; }
CUresult cuStreamIsCapturing(CUstream s, CUstreamCaptureStatus* st) {
  *st = stub_capturing(s) ? CU_STREAM_CAPTURE_STATUS_ACTIVE : CU_STREAM_CAPTURE_STATUS_NONE;
  return CUDA_SUCCESS;
}
static int fake_graph_obj;
CUresult cuStreamBeginCapture_v2(CUstream s, CUstreamCaptureMode m) { if (!s) return CUDA_ERROR_INVALID_VALUE; *(int*)s = 1; return CUDA_SUCCESS; }
CUresult cuStreamEndCapture(CUstream s, CUgraph* g) { if (!s) return CUDA_ERROR_INVALID_VALUE; *(int*)s = 0; *g = (CUgraph)&fake_graph_obj; return CUDA_SUCCESS; }
CUresult cuGraphGetNodes(CUgraph g, CUgraphNode* nodes, size_t* n) { *n = 0; return CUDA_SUCCESS; }
CUresult cuGraphLaunch(CUgraphExec ge, CUstream s) { __atomic_add_fetch(&n_launch, 1, __ATOMIC_RELAXED); enqueue(kernel_ns()); return CUDA_SUCCESS; }
CUresult cuGraphInstantiateWithFlags(CUgraphExec* ge, CUgraph g, unsigned long long fl) { *ge = (CUgraphExec)&fake_graph_obj; return CUDA_SUCCESS; }
CUresult cuThreadExchangeStreamCaptureMode(CUstreamCaptureMode* m) { return CUDA_SUCCESS; }

struct stub_event { int64_t t; int recorded; };
CUresult cuEventCreate(CUevent* e, unsigned int fl) { *e = (CUevent)calloc(1, sizeof(struct stub_event)); return CUDA_SUCCESS; }
CUresult cuEventDestroy_v2(CUevent e) { free(e); return CUDA_SUCCESS; }
CUresult cuEventRecord(CUevent e, CUstream s) {
  n_event_record++;
  if (stub_capturing(s)) n_event_on_capturing++;
  struct stub_event* ev = (struct stub_event*)e;
  ev->t = timeline();
  ev->recorded = 1;
  return CUDA_SUCCESS;
}
CUresult cuEventSynchronize(CUevent e) { wait_until(((struct stub_event*)e)->t); return CUDA_SUCCESS; }
CUresult cuEventQuery(CUevent e) { return now_ns() >= ((struct stub_event*)e)->t ? CUDA_SUCCESS : CUDA_ERROR_NOT_READY; }
CUresult cuEventElapsedTime(float* ms, CUevent a, CUevent b) {
  struct stub_event *x = (struct stub_event*)a, *y = (struct stub_event*)b;
  if (!x->recorded || !y->recorded) return CUDA_ERROR_INVALID_HANDLE;
  if (now_ns() < x->t || now_ns() < y->t) return CUDA_ERROR_NOT_READY;
  *ms = (float)((double)(y->t - x->t) / 1e6);
  return CUDA_SUCCESS;
}

static struct { CUdeviceptr p; size_t n; } live_tab[1 << 16];
static int live_n;
static CUresult fake_alloc(CUdeviceptr* p, size_t bytes) {
  pthread_mutex_lock(&mu);
  if (bytes_live + bytes > total_mem()) {
    pthread_mutex_unlock(&mu);
    return CUDA_ERROR_OUT_OF_MEMORY;
  }
  n_alloc++;
  *p = next_addr;
  next_addr += (bytes + 0x1fffff) & ~(uint64_t)0x1fffff;  /* 2 MiB granularity like the real driver */
  next_addr += 0x200000;
  bytes_live += bytes;
  if (live_n < (1 << 16)) { live_tab[live_n].p = *p; live_tab[live_n].n = bytes; live_n++; }
  pthread_mutex_unlock(&mu);
  return CUDA_SUCCESS;
}
CUresult cuMemAlloc_v2(CUdeviceptr* p, size_t bytes) { return fake_alloc(p, bytes); }
CUresult cuMemAllocManaged(CUdeviceptr* p, size_t bytes, unsigned int fl) { return fake_alloc(p, bytes); }
CUresult cuMemAllocPitch_v2(CUdeviceptr* p, size_t* pitch, size_t w, size_t h, unsigned int es) {
  *pitch = (w + 511) & ~(size_t)511;
  return fake_alloc(p, *pitch * h);
}
CUresult cuMemFree_v2(CUdeviceptr p) {
  pthread_mutex_lock(&mu);
  n_free++;
  for (int i = 0; i < live_n; i++)
    if (live_tab[i].p == p) {
      bytes_live -= live_tab[i].n;
      live_tab[i] = live_tab[--live_n];
      break;
    }
  pthread_mutex_unlock(&mu);
  return CUDA_SUCCESS;
}
CUresult cuMemHostAlloc(void** pp, size_t bytes, unsigned int fl) { *pp = calloc(1, bytes); return *pp ? CUDA_SUCCESS : CUDA_ERROR_OUT_OF_MEMORY; }
CUresult cuMemFreeHost(void* p) { free(p); return CUDA_SUCCESS; }
CUresult cuMemAllocHost_v2(void** pp, size_t bytes) { *pp = calloc(1, bytes); return *pp ? CUDA_SUCCESS : CUDA_ERROR_OUT_OF_MEMORY; }
CUresult cuMemHostGetDevicePointer_v2(CUdeviceptr* d, void* p, unsigned int fl) { *d = (CUdeviceptr)(uintptr_t)p; return CUDA_SUCCESS; }
CUresult cuMemHostRegister_v2(void* p, size_t bytes, unsigned int fl) { return CUDA_SUCCESS; }
CUresult cuMemHostUnregister(void* p) { return CUDA_SUCCESS; }
CUresult cuMemcpyHtoDAsync_v2(CUdeviceptr d, const void* s, size_t n, CUstream st) { n_memcpy++; return CUDA_SUCCESS; }
CUresult cuMemcpyDtoHAsync_v2(void* d, CUdeviceptr s, size_t n, CUstream st) { n_memcpy++; return CUDA_SUCCESS; }
CUresult cuMemsetD8Async(CUdeviceptr d, unsigned char v, size_t n, CUstream st) { return CUDA_SUCCESS; }
CUresult cuMemcpyHtoD_v2(CUdeviceptr d, const void* s, size_t n) { n_memcpy++; wait_until(timeline()); return CUDA_SUCCESS; }
CUresult cuMemcpyDtoH_v2(void* d, CUdeviceptr s, size_t n) { n_memcpy++; wait_until(timeline()); return CUDA_SUCCESS; }
CUresult cuMemcpyAtoH_v2(void* d, CUarray a, size_t off, size_t n) { n_memcpy++; wait_until(timeline()); return CUDA_SUCCESS; }
CUresult cuMemcpyHtoA_v2(CUarray a, size_t off, const void* s, size_t n) { n_memcpy++; wait_until(timeline()); return CUDA_SUCCESS; }

CUresult cuArrayCreate_v2(CUarray* h, const CUDA_ARRAY_DESCRIPTOR* d) { *h = (CUarray)malloc(16); return CUDA_SUCCESS; }
CUresult cuArray3DCreate_v2(CUarray* h, const CUDA_ARRAY3D_DESCRIPTOR* d) { *h = (CUarray)malloc(16); return CUDA_SUCCESS; }
CUresult cuArrayDestroy(CUarray h) { free(h); return CUDA_SUCCESS; }
CUresult cuMipmappedArrayCreate(CUmipmappedArray* h, const CUDA_ARRAY3D_DESCRIPTOR* d, unsigned int l) { *h = (CUmipmappedArray)malloc(16); return CUDA_SUCCESS; }
CUresult cuMipmappedArrayDestroy(CUmipmappedArray h) { free(h); return CUDA_SUCCESS; }

CUresult cuLaunchKernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                        unsigned sh, CUstream s, void** p, void** e) {
  __atomic_add_fetch(&n_launch, 1, __ATOMIC_RELAXED);
  if (stub_capturing(s)) return CUDA_SUCCESS; /* captured, not executed */
  if (kernel_ns() == 0) return CUDA_SUCCESS;  /* STUB_KERNEL_US=0: a free GPU, for measuring the hook's own host cost */
  enqueue(kernel_ns());
  return CUDA_SUCCESS;
}
CUresult cuLaunchCooperativeKernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by,
                                   unsigned bz, unsigned sh, CUstream s, void** p) {
  __atomic_add_fetch(&n_coop, 1, __ATOMIC_RELAXED);
  enqueue(kernel_ns());
  return CUDA_SUCCESS;
}
CUresult cuLaunchKernelEx(const CUlaunchConfig* c, CUfunction f, void** p, void** e) {
  __atomic_add_fetch(&n_launch, 1, __ATOMIC_RELAXED);
  enqueue(kernel_ns());
  return CUDA_SUCCESS;
}

/* cuGetProcAddress: resolve "<symbol>_v2" / "<symbol>" inside this library */
static void* self_handle(void) {
  static void* h;
  if (!h) {
    Dl_info info;
    if (dladdr((void*)&cuInit, &info)) h = dlopen(info.dli_fname, RTLD_NOW | RTLD_NOLOAD);
  }
  return h;
}
static void* lookup(const char* symbol) {
  char buf[128];
  void* h = self_handle();
  if (!h) return NULL;
  void* (*real_dlsym)(void*, const char*) = (void* (*)(void*, const char*))dlvsym(RTLD_NEXT, "dlsym", "GLIBC_2.2.5");
  if (!real_dlsym) real_dlsym = (void* (*)(void*, const char*))dlvsym(RTLD_NEXT, "dlsym", "GLIBC_2.34");
  snprintf(buf, sizeof(buf), "%s_v2", symbol);
  void* p = real_dlsym(h, buf);
  if (!p) p = real_dlsym(h, symbol);
  return p;
}
CUresult cuGetProcAddress_v2(const char* symbol, void** pfn, int ver, cuuint64_t flags, CUdriverProcAddressQueryResult* st) {
  n_getproc++;
  *pfn = lookup(symbol);
  if (st) *st = *pfn ? CU_GET_PROC_ADDRESS_SUCCESS : CU_GET_PROC_ADDRESS_SYMBOL_NOT_FOUND;
  return *pfn ? CUDA_SUCCESS : CUDA_ERROR_NOT_FOUND;
}
CUresult cuGetProcAddress(const char* symbol, void** pfn, int ver, cuuint64_t flags) {
  n_getproc++;
  *pfn = lookup(symbol);
  return *pfn ? CUDA_SUCCESS : CUDA_ERROR_NOT_FOUND;
}
CUresult cuMemsetD8_v2(CUdeviceptr d, unsigned char v, size_t n) { return CUDA_SUCCESS; }
CUresult cuMemAllocAsync(CUdeviceptr* p, size_t bytes, CUstream s) { return fake_alloc(p, bytes); }
CUresult cuMemAllocFromPoolAsync(CUdeviceptr* p, size_t bytes, CUmemoryPool pool, CUstream s) { return fake_alloc(p, bytes); }
CUresult cuMemFreeAsync(CUdeviceptr p, CUstream s) { return cuMemFree_v2(p); }
CUresult cuMemGetAllocationGranularity(size_t* g, const CUmemAllocationProp* prop, CUmemAllocationGranularity_flags o) { *g = 1000; return CUDA_SUCCESS; }
CUresult cuMemCreate(CUmemGenericAllocationHandle* h, size_t size, const CUmemAllocationProp* prop, unsigned long long flags) {
  CUdeviceptr p;
  CUresult r = fake_alloc(&p, size);
  *h = (CUmemGenericAllocationHandle)p;
  return r;
}
CUresult cuMemRelease(CUmemGenericAllocationHandle h) { return cuMemFree_v2((CUdeviceptr)h); }
