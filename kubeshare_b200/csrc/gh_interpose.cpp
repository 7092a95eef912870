// gh_interpose.cpp -- the hook surface: every symbol the reference libgemhook.so.1 interposes
// (reference hook.cpp:109-159 dlsym, :875-980 cuGetProcAddress, :1005-1062 wrappers, :857-872 mem-info),
// plus cuGetProcAddress_v2 -- without which a CUDA >= 12 runtime never sees a hook -- the _ptsz/_ptds
// twins, cuLaunchKernelEx and a correctly exported cuMemcpyDtoH_v2.
//
// Launch fast path = one load of gh_gate_open + one relaxed add + the indirect call into the driver.
#include <string.h>

#include "gh_internal.h"

#define GH_HOOK extern "C" __attribute__((visibility("default"))) CUresult CUDAAPI

// CU_HOOK_DEBUG=1: per-symbol call counters (reference hook.cpp:87-100, 783, 860, 868, 991 keep the same counts in
// hookInfo::call_count and never print them; ours are readable through gemhook_call_counts()).
#define GH_COUNTED(X)                                                                                              \
  X(cuLaunchKernel) X(cuLaunchCooperativeKernel) X(cuLaunchKernelEx) X(cuGraphLaunch) X(cuMemAlloc) X(cuMemAllocManaged) \
  X(cuMemAllocPitch) X(cuMemFree) X(cuArrayCreate) X(cuArray3DCreate) X(cuMipmappedArrayCreate) X(cuArrayDestroy)  \
  X(cuMipmappedArrayDestroy) X(cuMemGetInfo) X(cuDeviceTotalMem) X(cuCtxSynchronize) X(cuMemcpyAtoH)               \
  X(cuMemcpyDtoH) X(cuMemcpyHtoA) X(cuMemcpyHtoD) X(cuMemAllocAsync) X(cuMemAllocFromPoolAsync) X(cuMemFreeAsync)   \
  X(cuMemCreate) X(cuMemRelease) X(cuStreamSynchronize) X(cuEventSynchronize) X(cuStreamDestroy) X(cuMemAllocHost) \
  X(cuMemHostAlloc) X(cuMemFreeHost) X(cuGetProcAddress) X(dlsym)
enum {
#define X(n) CNT_##n,
  GH_COUNTED(X)
#undef X
  CNT_MAX
};
static uint64_t g_calls[CNT_MAX];
static const char* const g_call_names[CNT_MAX + 1] = {
#define X(n) #n,
    GH_COUNTED(X)
#undef X
    nullptr};
#define GH_COUNT(n)                                                                                   \
  do {                                                                                                \
    if (__builtin_expect(__atomic_load_n(&gh_hook_debug, __ATOMIC_RELAXED), 0)) __atomic_add_fetch(&g_calls[CNT_##n], 1, __ATOMIC_RELAXED); \
  } while (0)
extern "C" __attribute__((visibility("default"))) size_t gemhook_call_counts(const char* const** names, const uint64_t** counts) {
  if (names) *names = g_call_names;
  if (counts) *counts = g_calls;
  return CNT_MAX;
}

// late resolution for names that are not part of the core table (ptsz/ptds twins, Ex)
static void* resolve_late(void** slot, const char* name) {
  if (gh_driver_init() != 0) return nullptr;
  void* p = gh_true_dlsym(gh_real.handle, name);
  if (!p) {  // a driver without the per-thread-default-stream twin: fall back to the plain entry point
    size_t n = strlen(name);
    if (n > 5 && n < 96 && (!strcmp(name + n - 5, "_ptsz") || !strcmp(name + n - 5, "_ptds"))) {
      char base[96];
      memcpy(base, name, n - 5);
      base[n - 5] = 0;
      p = gh_true_dlsym(gh_real.handle, base);
    }
  }
  __atomic_store_n(slot, p, __ATOMIC_RELEASE);
  return p;
}
#define GH_REAL_CORE(name) \
  ((decltype(&name))(gh_real.name ? gh_real.name : resolve_late(&gh_real.name, #name)))

// ---- launches ------------------------------------------------------------------------------------------
// FAST PATH, shaped by hand because it decides the overhead where the launch storm is host-bound (measured on one of
// the pool's boxes: 2.01 us per launch un-hooked, so every 20 ns of hook is 1 %).  A launch hook is a leaf that either
// tail-jumps into the driver with its arguments untouched, or tail-jumps into a same-signature slow twin:
//     load the driver entry + gh_gate_fast (one cache line), plain increment of the thread's TLS counter, mask test, jmp.
// No lock prefix (round 1: lock xadd on a shared counter), no register saves (round 1: six pushes and five stack
// reloads, because the slow-path call sat in the same function).
static inline __attribute__((always_inline)) bool launch_fast_ok(void) {  // (C twin of the assembly below, other ISAs)
  if (__builtin_expect(!__atomic_load_n(&gh_gate_fast, __ATOMIC_RELAXED) || !gh_tl.registered, 0)) return false;
  uint64_t n = gh_tl.count + 1;
  if (__builtin_expect(((uint32_t)n & gh_seg_mask) == 0, 0)) return false;  // segment tick due: slow twin
  __atomic_store_n(&gh_tl.count, n, __ATOMIC_RELAXED);
  return true;
}
// everything else: gate closed (burst edge / token), first launch of a thread, CU_HOOK_DEBUG counting, segment tick
static __attribute__((noinline)) void launch_slow_common(CUstream hStream, int counter) {
  if (__atomic_load_n(&gh_hook_debug, __ATOMIC_RELAXED)) __atomic_add_fetch(&g_calls[counter], 1, __ATOMIC_RELAXED);
  gh_thread_register();
  if (__atomic_load_n(&gh_gate_open, __ATOMIC_RELAXED) == 0) gh_launch_slow(hStream);
  uint64_t n = gh_tl.count + 1;
  __atomic_store_n(&gh_tl.count, n, __ATOMIC_RELAXED);
  if (((uint32_t)n & gh_seg_mask) == 0 && __atomic_load_n(&gh_gate_open, __ATOMIC_RELAXED)) gh_segment_tick(hStream);
}

typedef CUresult(CUDAAPI* launch_fn)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                     CUstream, void**, void**);
typedef CUresult(CUDAAPI* coop_fn)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                   CUstream, void**);
typedef CUresult(CUDAAPI* launchex_fn)(const CUlaunchConfig*, CUfunction, void**, void**);
typedef CUresult(CUDAAPI* graphlaunch_fn)(CUgraphExec, CUstream);

static void* p_launch_ptsz;
static void* p_coop_ptsz;
static void* p_launchex;
static void* p_launchex_ptsz;
static void *p_graphlaunch, *p_graphlaunch_pt;
#define LATE(slot, name) ((slot) ? (slot) : resolve_late(&(slot), name))

// The slow twin has the hook's exact signature, so both exits of the hook are plain jumps.  On x86-64 the hook itself
// is written in assembly: left to the compiler the "leaf with two tail calls" still saved six registers and copied the
// five stack arguments to registers and back (it needs scratch registers and does not see that the outgoing stack
// arguments ARE the incoming ones).  12 instructions, scratch registers rax/r10/r11 only (never argument registers),
// two cache lines (gh_hot and the thread's TLS line):
//     real = gh_hot.fn[i]; if (!real || !gh_hot.gate_fast) slow;  tl = %fs + gh_hot.tls_off;
//     if (!tl->registered) slow;  n = tl->count + 1;  if (!(n & gh_hot.seg_mask)) slow;  tl->count = n;  jmp *real
#define GH_STR2(x) #x
#define GH_STR(x) GH_STR2(x)
#if defined(__x86_64__) && !defined(GEMHOOK_NO_ASM_FASTPATH)
#define GH_LAUNCH_ENTRY(name, fn_t, idx, params, args)                                                          \
  asm(".text\n.p2align 5\n.globl " #name "\n.type " #name ",@function\n" #name ":\n"                          \
      "  endbr64\n"                                                                                            \
      "  movq gh_hot+16+8*" GH_STR(idx) "(%rip), %rax\n"                                                       \
      "  testq %rax, %rax\n"                                                                                   \
      "  jz 1f\n"                                                                                              \
      "  cmpl $0, gh_hot(%rip)\n"                                                                              \
      "  je 1f\n"                                                                                              \
      "  movq gh_hot+8(%rip), %r10\n"                                                                          \
      "  cmpq $0, %fs:8(%r10)\n"                                                                               \
      "  je 1f\n"                                                                                              \
      "  movq %fs:(%r10), %r11\n"                                                                              \
      "  incq %r11\n"                                                                                          \
      "  testl %r11d, gh_hot+4(%rip)\n"                                                                        \
      "  jz 1f\n"                                                                                              \
      "  movq %r11, %fs:(%r10)\n"                                                                              \
      "  jmp *%rax\n"                                                                                          \
      "1:\n"                                                                                                   \
      "  jmp " #name "_slowtwin\n"                                                                             \
      ".size " #name ", .-" #name "\n");
#else
#define GH_LAUNCH_ENTRY(name, fn_t, idx, params, args)                                                          \
  GH_HOOK name params {                                                                                         \
    fn_t real = (fn_t)__atomic_load_n(&gh_hot.fn[idx], __ATOMIC_RELAXED);                                       \
    if (__builtin_expect(real != nullptr && launch_fast_ok(), 1)) return real args;                             \
    return name##_slowtwin args;                                                                                \
  }
#endif
#define GH_FASTFN_PUBLISH(idx, ptr) __atomic_store_n(&gh_hot.fn[idx], (void*)(ptr), __ATOMIC_RELEASE)

#define GH_LAUNCH_HOOK(name, fn_t, idx, slot, symbol, counter, stream_expr, params, args)                      \
  extern "C" __attribute__((visibility("hidden"), used, noinline)) CUresult CUDAAPI name##_slowtwin params;     \
  extern "C" __attribute__((visibility("default"))) CUresult CUDAAPI name params;                               \
  GH_LAUNCH_ENTRY(name, fn_t, idx, params, args)                                                                \
  CUresult CUDAAPI name##_slowtwin params {                                                                     \
    launch_slow_common(stream_expr, counter);                                                                   \
    fn_t real = (fn_t)LATE(slot, symbol);                                                                       \
    GH_FASTFN_PUBLISH(idx, real);                                                                               \
    return real args;                                                                                           \
  }
// hooks that are not worth a place in the hot line (one graph launch stands for many kernels): plain C, same gate
#define GH_LAUNCH_HOOK_COLD(name, fn_t, slot, symbol, counter, stream_expr, params, args)                       \
  GH_HOOK name params {                                                                                         \
    launch_slow_common(stream_expr, counter);                                                                   \
    return ((fn_t)LATE(slot, symbol))args;                                                                      \
  }

#define LAUNCH_PARAMS (CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz, unsigned shmem, \
                       CUstream hStream, void** params, void** extra)
#define LAUNCH_ARGS (f, gx, gy, gz, bx, by, bz, shmem, hStream, params, extra)
#define COOP_PARAMS (CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz, unsigned shmem, \
                     CUstream hStream, void** params)
#define COOP_ARGS (f, gx, gy, gz, bx, by, bz, shmem, hStream, params)
#define EX_PARAMS (const CUlaunchConfig* config, CUfunction f, void** params, void** extra)
#define EX_ARGS (config, f, params, extra)

GH_LAUNCH_HOOK(cuLaunchKernel, launch_fn, 0, gh_real.cuLaunchKernel, "cuLaunchKernel", CNT_cuLaunchKernel, hStream, LAUNCH_PARAMS, LAUNCH_ARGS)
GH_LAUNCH_HOOK(cuLaunchCooperativeKernel, coop_fn, 1, gh_real.cuLaunchCooperativeKernel, "cuLaunchCooperativeKernel",
               CNT_cuLaunchCooperativeKernel, hStream, COOP_PARAMS, COOP_ARGS)
GH_LAUNCH_HOOK(cuLaunchKernel_ptsz, launch_fn, 2, p_launch_ptsz, "cuLaunchKernel_ptsz", CNT_cuLaunchKernel, hStream, LAUNCH_PARAMS, LAUNCH_ARGS)
GH_LAUNCH_HOOK(cuLaunchCooperativeKernel_ptsz, coop_fn, 3, p_coop_ptsz, "cuLaunchCooperativeKernel_ptsz", CNT_cuLaunchCooperativeKernel,
               hStream, COOP_PARAMS, COOP_ARGS)
GH_LAUNCH_HOOK(cuLaunchKernelEx, launchex_fn, 4, p_launchex, "cuLaunchKernelEx", CNT_cuLaunchKernelEx, (config ? config->hStream : nullptr),
               EX_PARAMS, EX_ARGS)
GH_LAUNCH_HOOK(cuLaunchKernelEx_ptsz, launchex_fn, 5, p_launchex_ptsz, "cuLaunchKernelEx_ptsz", CNT_cuLaunchKernelEx,
               (config ? config->hStream : nullptr), EX_PARAMS, EX_ARGS)
// graph launches pass the token gate like a kernel launch
GH_LAUNCH_HOOK_COLD(cuGraphLaunch, graphlaunch_fn, p_graphlaunch, "cuGraphLaunch", CNT_cuGraphLaunch, hStream, (CUgraphExec g, CUstream hStream),
               (g, hStream))
GH_LAUNCH_HOOK_COLD(cuGraphLaunch_ptsz, graphlaunch_fn, p_graphlaunch_pt, "cuGraphLaunch_ptsz", CNT_cuGraphLaunch, hStream,
               (CUgraphExec g, CUstream hStream), (g, hStream))

// ---- gpu_mem cap ----------------------------------------------------------------------------------------
GH_HOOK cuMemAlloc_v2(CUdeviceptr* dptr, size_t bytesize) {
  GH_COUNT(cuMemAlloc);
  if (!gh_mem_reserve(bytesize)) return CUDA_ERROR_OUT_OF_MEMORY;  // the driver is not called (hook.cpp:995)
  CUresult r = GH_REAL_CORE(cuMemAlloc_v2)(dptr, bytesize);
  if (r != CUDA_SUCCESS) {
    gh_mem_unreserve(bytesize);
    return r;
  }
  gh_mem_commit((uint64_t)*dptr, bytesize);
  return r;
}

GH_HOOK cuMemAllocManaged(CUdeviceptr* dptr, size_t bytesize, unsigned int flags) {
  GH_COUNT(cuMemAllocManaged);
  gh_live_get();
  // managed memory is not accounted by the reference (hook.cpp:619-627: empty pre/post hooks) -- kept as the default;
  // GEMHOOK_ACCOUNT_MANAGED=1 charges it like cuMemAlloc
  if (!gh_cfg.account_managed) return GH_REAL_CORE(cuMemAllocManaged)(dptr, bytesize, flags);
  if (!gh_mem_reserve(bytesize)) return CUDA_ERROR_OUT_OF_MEMORY;
  CUresult r = GH_REAL_CORE(cuMemAllocManaged)(dptr, bytesize, flags);
  if (r != CUDA_SUCCESS) {
    gh_mem_unreserve(bytesize);
    return r;
  }
  gh_mem_commit((uint64_t)*dptr, bytesize);
  return r;
}

GH_HOOK cuMemAllocPitch_v2(CUdeviceptr* dptr, size_t* pPitch, size_t WidthInBytes, size_t Height,
                           unsigned int ElementSizeBytes) {
  GH_COUNT(cuMemAllocPitch);
  // the charge is pitch * Height with the REAL pitch (reference post-hook, hook.cpp:633-636; its pre-hook
  // reads *pPitch before the driver wrote it, :629-632).  The pitch is only known after the call, so the
  // reservation follows it and a denied reservation frees the allocation again.
  gh_live_get();
  CUresult r = GH_REAL_CORE(cuMemAllocPitch_v2)(dptr, pPitch, WidthInBytes, Height, ElementSizeBytes);
  if (r != CUDA_SUCCESS) return r;
  uint64_t bytes = (uint64_t)(*pPitch) * Height;
  if (!gh_mem_reserve(bytes)) {
    GH_REAL_CORE(cuMemFree_v2)(*dptr);
    return CUDA_ERROR_OUT_OF_MEMORY;
  }
  gh_mem_commit((uint64_t)*dptr, bytes);
  return r;
}

GH_HOOK cuMemFree_v2(CUdeviceptr dptr) {
  GH_COUNT(cuMemFree);
  gh_mem_free_key((uint64_t)dptr);
  return GH_REAL_CORE(cuMemFree_v2)(dptr);
}

GH_HOOK cuArrayCreate_v2(CUarray* pHandle, const CUDA_ARRAY_DESCRIPTOR* d) {
  GH_COUNT(cuArrayCreate);
  uint64_t bytes = gemhook_array_bytes(d->Width, d->Height, 0, d->NumChannels, (uint32_t)d->Format, 0);
  if (!gh_mem_reserve(bytes)) return CUDA_ERROR_OUT_OF_MEMORY;
  CUresult r = GH_REAL_CORE(cuArrayCreate_v2)(pHandle, d);
  if (r != CUDA_SUCCESS) {
    gh_mem_unreserve(bytes);
    return r;
  }
  gh_mem_commit((uint64_t)(uintptr_t)*pHandle, bytes);
  return r;
}

GH_HOOK cuArray3DCreate_v2(CUarray* pHandle, const CUDA_ARRAY3D_DESCRIPTOR* d) {
  GH_COUNT(cuArray3DCreate);
  uint64_t bytes = gemhook_array_bytes(d->Width, d->Height, d->Depth, d->NumChannels, (uint32_t)d->Format, 1);
  if (!gh_mem_reserve(bytes)) return CUDA_ERROR_OUT_OF_MEMORY;
  CUresult r = GH_REAL_CORE(cuArray3DCreate_v2)(pHandle, d);
  if (r != CUDA_SUCCESS) {
    gh_mem_unreserve(bytes);
    return r;
  }
  gh_mem_commit((uint64_t)(uintptr_t)*pHandle, bytes);
  return r;
}

GH_HOOK cuMipmappedArrayCreate(CUmipmappedArray* pHandle, const CUDA_ARRAY3D_DESCRIPTOR* d, unsigned int levels) {
  GH_COUNT(cuMipmappedArrayCreate);
  gh_live_get();
  // not accounted by the reference (hook.cpp:682-694); with GEMHOOK_ACCOUNT_MANAGED=1 every level is charged with
  // the array rule of hook.cpp:668-680, each extent halving (floor, at least 1) per level
  if (!gh_cfg.account_managed || !d) return GH_REAL_CORE(cuMipmappedArrayCreate)(pHandle, d, levels);
  uint64_t bytes = gemhook_mipmap_bytes(d->Width, d->Height, d->Depth, d->NumChannels, (uint32_t)d->Format, levels);
  if (!gh_mem_reserve(bytes)) return CUDA_ERROR_OUT_OF_MEMORY;
  CUresult r = GH_REAL_CORE(cuMipmappedArrayCreate)(pHandle, d, levels);
  if (r != CUDA_SUCCESS) {
    gh_mem_unreserve(bytes);
    return r;
  }
  gh_mem_commit((uint64_t)(uintptr_t)*pHandle, bytes);
  return r;
}

GH_HOOK cuArrayDestroy(CUarray hArray) {
  GH_COUNT(cuArrayDestroy);
  gh_mem_free_key((uint64_t)(uintptr_t)hArray);
  return GH_REAL_CORE(cuArrayDestroy)(hArray);
}

GH_HOOK cuMipmappedArrayDestroy(CUmipmappedArray h) {
  GH_COUNT(cuMipmappedArrayDestroy);
  gh_mem_free_key((uint64_t)(uintptr_t)h);
  return GH_REAL_CORE(cuMipmappedArrayDestroy)(h);
}

// mem-info virtualisation: the driver is never asked (hook.cpp:857-872)
extern "C" int gh_live_enabled(void);
GH_HOOK cuMemGetInfo_v2(size_t* free_b, size_t* total_b) {
  GH_COUNT(cuMemGetInfo);
  if (!gh_live_get() || !gh_live_enabled()) return GH_REAL_CORE(cuMemGetInfo_v2)(free_b, total_b);
  uint64_t f = 0, t = 0;
  gh_mem_info(&f, &t);
  if (free_b) *free_b = (size_t)f;
  if (total_b) *total_b = (size_t)t;
  return CUDA_SUCCESS;
}
GH_HOOK cuDeviceTotalMem_v2(size_t* bytes, CUdevice dev) {
  GH_COUNT(cuDeviceTotalMem);
  if (!gh_live_get() || !gh_live_enabled()) return GH_REAL_CORE(cuDeviceTotalMem_v2)(bytes, dev);
  uint64_t f = 0, t = 0;
  gh_mem_info(&f, &t);
  if (bytes) *bytes = (size_t)t;
  return CUDA_SUCCESS;
}

// ---- synchronising calls: burst end / window start (hook.cpp:696-722) -------------------------------------
GH_HOOK cuCtxSynchronize(void) {
  GH_COUNT(cuCtxSynchronize);
  gh_host_sync_pre();
  CUresult r = GH_REAL_CORE(cuCtxSynchronize)();
  if (r == CUDA_SUCCESS) gh_host_sync_post();
  return r;
}

#define GH_SYNC_COPY(counter, export_name, real_expr, params, args) \
  GH_HOOK export_name params {                             \
    if (__builtin_expect(__atomic_load_n(&gh_hook_debug, __ATOMIC_RELAXED), 0)) __atomic_add_fetch(&g_calls[counter], 1, __ATOMIC_RELAXED); \
    gh_host_sync_pre();                                    \
    CUresult r = (real_expr)args;                          \
    if (r == CUDA_SUCCESS) gh_host_sync_post();            \
    return r;                                              \
  }

typedef CUresult(CUDAAPI* atoh_fn)(void*, CUarray, size_t, size_t);
typedef CUresult(CUDAAPI* dtoh_fn)(void*, CUdeviceptr, size_t);
typedef CUresult(CUDAAPI* htoa_fn)(CUarray, size_t, const void*, size_t);
typedef CUresult(CUDAAPI* htod_fn)(CUdeviceptr, const void*, size_t);
static void *p_atoh_ptds, *p_dtoh_ptds, *p_htoa_ptds, *p_htod_ptds;

GH_SYNC_COPY(CNT_cuMemcpyAtoH, cuMemcpyAtoH_v2, GH_REAL_CORE(cuMemcpyAtoH_v2), (void* dst, CUarray src, size_t off, size_t n), (dst, src, off, n))
GH_SYNC_COPY(CNT_cuMemcpyDtoH, cuMemcpyDtoH_v2, GH_REAL_CORE(cuMemcpyDtoH_v2), (void* dst, CUdeviceptr src, size_t n), (dst, src, n))
GH_SYNC_COPY(CNT_cuMemcpyHtoA, cuMemcpyHtoA_v2, GH_REAL_CORE(cuMemcpyHtoA_v2), (CUarray dst, size_t off, const void* src, size_t n), (dst, off, src, n))
GH_SYNC_COPY(CNT_cuMemcpyHtoD, cuMemcpyHtoD_v2, GH_REAL_CORE(cuMemcpyHtoD_v2), (CUdeviceptr dst, const void* src, size_t n), (dst, src, n))
GH_SYNC_COPY(CNT_cuMemcpyAtoH, cuMemcpyAtoH_v2_ptds, (atoh_fn)LATE(p_atoh_ptds, "cuMemcpyAtoH_v2_ptds"), (void* dst, CUarray src, size_t off, size_t n), (dst, src, off, n))
GH_SYNC_COPY(CNT_cuMemcpyDtoH, cuMemcpyDtoH_v2_ptds, (dtoh_fn)LATE(p_dtoh_ptds, "cuMemcpyDtoH_v2_ptds"), (void* dst, CUdeviceptr src, size_t n), (dst, src, n))
GH_SYNC_COPY(CNT_cuMemcpyHtoA, cuMemcpyHtoA_v2_ptds, (htoa_fn)LATE(p_htoa_ptds, "cuMemcpyHtoA_v2_ptds"), (CUarray dst, size_t off, const void* src, size_t n), (dst, off, src, n))
GH_SYNC_COPY(CNT_cuMemcpyHtoD, cuMemcpyHtoD_v2_ptds, (htod_fn)LATE(p_htod_ptds, "cuMemcpyHtoD_v2_ptds"), (CUdeviceptr dst, const void* src, size_t n), (dst, src, n))

// ---- modern entry points the reference never saw ------------------------------------------
// Stream-ordered allocations are charged like cuMemAlloc; graph launches pass the token gate like a kernel
// launch; cuStreamSynchronize / cuEventSynchronize count as host syncs only with GEMHOOK_EXTRA_HOOKS=1 (the
// reference's burst detection knows cuCtxSynchronize and the four blocking copies only, hook.cpp:696-722).
typedef CUresult(CUDAAPI* allocasync_fn)(CUdeviceptr*, size_t, CUstream);
typedef CUresult(CUDAAPI* allocpool_fn)(CUdeviceptr*, size_t, CUmemoryPool, CUstream);
typedef CUresult(CUDAAPI* freeasync_fn)(CUdeviceptr, CUstream);
typedef CUresult(CUDAAPI* streamsync_fn)(CUstream);
typedef CUresult(CUDAAPI* eventsync_fn)(CUevent);
static void *p_allocasync, *p_allocasync_pt, *p_allocpool, *p_allocpool_pt, *p_freeasync, *p_freeasync_pt;
static void *p_streamsync, *p_streamsync_pt, *p_eventsync;

#define GH_ALLOC_ASYNC(name, slot, sym)                                        \
  GH_HOOK name(CUdeviceptr* dptr, size_t bytesize, CUstream hStream) {         \
    GH_COUNT(cuMemAllocAsync);                                                 \
    if (!gh_mem_reserve(bytesize)) return CUDA_ERROR_OUT_OF_MEMORY;            \
    CUresult r = ((allocasync_fn)LATE(slot, sym))(dptr, bytesize, hStream);    \
    if (r != CUDA_SUCCESS) {                                                   \
      gh_mem_unreserve(bytesize);                                              \
      return r;                                                                \
    }                                                                          \
    gh_mem_commit((uint64_t)*dptr, bytesize);                                  \
    return r;                                                                  \
  }
GH_ALLOC_ASYNC(cuMemAllocAsync, p_allocasync, "cuMemAllocAsync")
GH_ALLOC_ASYNC(cuMemAllocAsync_ptsz, p_allocasync_pt, "cuMemAllocAsync_ptsz")

#define GH_ALLOC_POOL(name, slot, sym)                                                  \
  GH_HOOK name(CUdeviceptr* dptr, size_t bytesize, CUmemoryPool pool, CUstream hStream) { \
    GH_COUNT(cuMemAllocFromPoolAsync);                                                  \
    if (!gh_mem_reserve(bytesize)) return CUDA_ERROR_OUT_OF_MEMORY;                     \
    CUresult r = ((allocpool_fn)LATE(slot, sym))(dptr, bytesize, pool, hStream);        \
    if (r != CUDA_SUCCESS) {                                                            \
      gh_mem_unreserve(bytesize);                                                       \
      return r;                                                                         \
    }                                                                                   \
    gh_mem_commit((uint64_t)*dptr, bytesize);                                           \
    return r;                                                                           \
  }
GH_ALLOC_POOL(cuMemAllocFromPoolAsync, p_allocpool, "cuMemAllocFromPoolAsync")
GH_ALLOC_POOL(cuMemAllocFromPoolAsync_ptsz, p_allocpool_pt, "cuMemAllocFromPoolAsync_ptsz")

GH_HOOK cuMemFreeAsync(CUdeviceptr dptr, CUstream hStream) {
  GH_COUNT(cuMemFreeAsync);
  gh_mem_free_key((uint64_t)dptr);
  return ((freeasync_fn)LATE(p_freeasync, "cuMemFreeAsync"))(dptr, hStream);
}
GH_HOOK cuMemFreeAsync_ptsz(CUdeviceptr dptr, CUstream hStream) {
  GH_COUNT(cuMemFreeAsync);
  gh_mem_free_key((uint64_t)dptr);
  return ((freeasync_fn)LATE(p_freeasync_pt, "cuMemFreeAsync_ptsz"))(dptr, hStream);
}
// Virtual-memory-management allocations (PyTorch "expandable segments"): the physical handle is what consumes
// device memory, so cuMemCreate is charged and cuMemRelease gives it back; mapping / address reservation is free.
typedef CUresult(CUDAAPI* memcreate_fn)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
typedef CUresult(CUDAAPI* memrelease_fn)(CUmemGenericAllocationHandle);
static void *p_memcreate, *p_memrelease;
static inline uint64_t handle_key(CUmemGenericAllocationHandle h) { return (uint64_t)h ^ 0x8000000000000000ULL; }
GH_HOOK cuMemCreate(CUmemGenericAllocationHandle* handle, size_t size, const CUmemAllocationProp* prop, unsigned long long flags) {
  GH_COUNT(cuMemCreate);
  bool device = !prop || prop->location.type == CU_MEM_LOCATION_TYPE_DEVICE;
  if (device && !gh_mem_reserve(size)) return CUDA_ERROR_OUT_OF_MEMORY;
  CUresult r = ((memcreate_fn)LATE(p_memcreate, "cuMemCreate"))(handle, size, prop, flags);
  if (!device) return r;
  if (r != CUDA_SUCCESS) {
    gh_mem_unreserve(size);
    return r;
  }
  gh_mem_commit(handle_key(*handle), size);
  return r;
}
GH_HOOK cuMemRelease(CUmemGenericAllocationHandle handle) {
  GH_COUNT(cuMemRelease);
  gh_mem_free_key(handle_key(handle));
  return ((memrelease_fn)LATE(p_memrelease, "cuMemRelease"))(handle);
}

static inline bool extra_syncs(void) { return gh_live_get() && gh_cfg.extra_hooks; }
GH_HOOK cuStreamSynchronize(CUstream hStream) {
  GH_COUNT(cuStreamSynchronize);
  bool x = extra_syncs();
  if (x) gh_host_sync_pre();
  CUresult r = ((streamsync_fn)LATE(p_streamsync, "cuStreamSynchronize"))(hStream);
  if (x && r == CUDA_SUCCESS) gh_host_sync_post();
  return r;
}
GH_HOOK cuStreamSynchronize_ptsz(CUstream hStream) {
  GH_COUNT(cuStreamSynchronize);
  bool x = extra_syncs();
  if (x) gh_host_sync_pre();
  CUresult r = ((streamsync_fn)LATE(p_streamsync_pt, "cuStreamSynchronize_ptsz"))(hStream);
  if (x && r == CUDA_SUCCESS) gh_host_sync_post();
  return r;
}
GH_HOOK cuEventSynchronize(CUevent ev) {
  GH_COUNT(cuEventSynchronize);
  bool x = extra_syncs();
  if (x) gh_host_sync_pre();
  CUresult r = ((eventsync_fn)LATE(p_eventsync, "cuEventSynchronize"))(ev);
  if (x && r == CUDA_SUCCESS) gh_host_sync_post();
  return r;
}

// cuStreamDestroy: a segment still open on the dying stream gets its end marker first (gh_hook.cpp)
typedef CUresult(CUDAAPI* streamdestroy_fn)(CUstream);
GH_HOOK cuStreamDestroy_v2(CUstream hStream) {
  GH_COUNT(cuStreamDestroy);
  if (gh_live_get()) gh_stream_destroyed(hStream);
  return GH_REAL_CORE(cuStreamDestroy_v2)(hStream);
}

// Pinned host memory is not device memory and the reference does not see it at all; GEMHOOK_ACCOUNT_HOST=1 charges
// it against gpu_mem anyway (a pod that pins host RAM takes it from every tenant of the node).  Off by default.
typedef CUresult(CUDAAPI* allochost_fn)(void**, size_t);
typedef CUresult(CUDAAPI* hostalloc_fn)(void**, size_t, unsigned int);
typedef CUresult(CUDAAPI* freehost_fn)(void*);
static inline uint64_t host_key(void* p) { return (uint64_t)(uintptr_t)p ^ 0x4000000000000000ULL; }
GH_HOOK cuMemAllocHost_v2(void** pp, size_t bytesize) {
  GH_COUNT(cuMemAllocHost);
  bool charge = gh_live_get() && gh_cfg.account_host;
  if (charge && !gh_mem_reserve(bytesize)) return CUDA_ERROR_OUT_OF_MEMORY;
  CUresult r = GH_REAL_CORE(cuMemAllocHost_v2)(pp, bytesize);
  if (!charge) return r;
  if (r != CUDA_SUCCESS) {
    gh_mem_unreserve(bytesize);
    return r;
  }
  gh_mem_commit(host_key(*pp), bytesize);
  return r;
}
GH_HOOK cuMemHostAlloc(void** pp, size_t bytesize, unsigned int flags) {
  GH_COUNT(cuMemHostAlloc);
  bool charge = gh_live_get() && gh_cfg.account_host;
  if (charge && !gh_mem_reserve(bytesize)) return CUDA_ERROR_OUT_OF_MEMORY;
  CUresult r = GH_REAL_CORE(cuMemHostAlloc)(pp, bytesize, flags);
  if (!charge) return r;
  if (r != CUDA_SUCCESS) {
    gh_mem_unreserve(bytesize);
    return r;
  }
  gh_mem_commit(host_key(*pp), bytesize);
  return r;
}
GH_HOOK cuMemFreeHost(void* p) {
  GH_COUNT(cuMemFreeHost);
  if (gh_live_get() && gh_cfg.account_host) gh_mem_free_key(host_key(p));
  return GH_REAL_CORE(cuMemFreeHost)(p);
}

// ---- symbol tables ---------------------------------------------------------------------------------------
extern "C" __attribute__((visibility("default"))) void* dlsym(void* handle, const char* symbol);
GH_HOOK cuGetProcAddress_v2(const char* symbol, void** pfn, int cudaVersion, cuuint64_t flags,
                            CUdriverProcAddressQueryResult* status);
#undef cuGetProcAddress
GH_HOOK cuGetProcAddress(const char* symbol, void** pfn, int cudaVersion, cuuint64_t flags);

struct HookEntry {
  const char* exported;  // dynamic symbol name (what dlsym / the linker sees)
  const char* base;      // name cuGetProcAddress is asked for
  void* fn;              // legacy-stream hook
  void* fn_pt;           // per-thread-default-stream hook, or NULL when the call has no stream semantics
};
static const HookEntry kHooks[] = {
    {"cuLaunchKernel", "cuLaunchKernel", (void*)&cuLaunchKernel, (void*)&cuLaunchKernel_ptsz},
    {"cuLaunchCooperativeKernel", "cuLaunchCooperativeKernel", (void*)&cuLaunchCooperativeKernel, (void*)&cuLaunchCooperativeKernel_ptsz},
    {"cuLaunchKernelEx", "cuLaunchKernelEx", (void*)&cuLaunchKernelEx, (void*)&cuLaunchKernelEx_ptsz},
    {"cuMemAlloc_v2", "cuMemAlloc", (void*)&cuMemAlloc_v2, nullptr},
    {"cuMemAllocManaged", "cuMemAllocManaged", (void*)&cuMemAllocManaged, nullptr},
    {"cuMemAllocPitch_v2", "cuMemAllocPitch", (void*)&cuMemAllocPitch_v2, nullptr},
    {"cuMemFree_v2", "cuMemFree", (void*)&cuMemFree_v2, nullptr},
    {"cuArrayCreate_v2", "cuArrayCreate", (void*)&cuArrayCreate_v2, nullptr},
    {"cuArray3DCreate_v2", "cuArray3DCreate", (void*)&cuArray3DCreate_v2, nullptr},
    {"cuMipmappedArrayCreate", "cuMipmappedArrayCreate", (void*)&cuMipmappedArrayCreate, nullptr},
    {"cuArrayDestroy", "cuArrayDestroy", (void*)&cuArrayDestroy, nullptr},
    {"cuMipmappedArrayDestroy", "cuMipmappedArrayDestroy", (void*)&cuMipmappedArrayDestroy, nullptr},
    {"cuMemGetInfo_v2", "cuMemGetInfo", (void*)&cuMemGetInfo_v2, nullptr},
    {"cuDeviceTotalMem_v2", "cuDeviceTotalMem", (void*)&cuDeviceTotalMem_v2, nullptr},
    {"cuCtxSynchronize", "cuCtxSynchronize", (void*)&cuCtxSynchronize, nullptr},
    {"cuMemcpyAtoH_v2", "cuMemcpyAtoH", (void*)&cuMemcpyAtoH_v2, (void*)&cuMemcpyAtoH_v2_ptds},
    {"cuMemcpyDtoH_v2", "cuMemcpyDtoH", (void*)&cuMemcpyDtoH_v2, (void*)&cuMemcpyDtoH_v2_ptds},
    {"cuMemcpyHtoA_v2", "cuMemcpyHtoA", (void*)&cuMemcpyHtoA_v2, (void*)&cuMemcpyHtoA_v2_ptds},
    {"cuMemcpyHtoD_v2", "cuMemcpyHtoD", (void*)&cuMemcpyHtoD_v2, (void*)&cuMemcpyHtoD_v2_ptds},
    {"cuMemAllocAsync", "cuMemAllocAsync", (void*)&cuMemAllocAsync, (void*)&cuMemAllocAsync_ptsz},
    {"cuMemAllocFromPoolAsync", "cuMemAllocFromPoolAsync", (void*)&cuMemAllocFromPoolAsync, (void*)&cuMemAllocFromPoolAsync_ptsz},
    {"cuMemFreeAsync", "cuMemFreeAsync", (void*)&cuMemFreeAsync, (void*)&cuMemFreeAsync_ptsz},
    {"cuGraphLaunch", "cuGraphLaunch", (void*)&cuGraphLaunch, (void*)&cuGraphLaunch_ptsz},
    {"cuMemCreate", "cuMemCreate", (void*)&cuMemCreate, nullptr},
    {"cuMemRelease", "cuMemRelease", (void*)&cuMemRelease, nullptr},
    {"cuStreamSynchronize", "cuStreamSynchronize", (void*)&cuStreamSynchronize, (void*)&cuStreamSynchronize_ptsz},
    {"cuEventSynchronize", "cuEventSynchronize", (void*)&cuEventSynchronize, nullptr},
    {"cuStreamDestroy_v2", "cuStreamDestroy", (void*)&cuStreamDestroy_v2, nullptr},
    {"cuMemAllocHost_v2", "cuMemAllocHost", (void*)&cuMemAllocHost_v2, nullptr},
    {"cuMemHostAlloc", "cuMemHostAlloc", (void*)&cuMemHostAlloc, nullptr},
    {"cuMemFreeHost", "cuMemFreeHost", (void*)&cuMemFreeHost, nullptr},
};
static const size_t kNumHooks = sizeof(kHooks) / sizeof(kHooks[0]);

static const char* const kHookedNames[] = {
    "dlsym", "cuGetProcAddress", "cuGetProcAddress_v2", "cuLaunchKernel", "cuLaunchCooperativeKernel",
    "cuLaunchKernelEx", "cuMemAlloc_v2", "cuMemAllocManaged", "cuMemAllocPitch_v2", "cuMemFree_v2",
    "cuArrayCreate_v2", "cuArray3DCreate_v2", "cuMipmappedArrayCreate", "cuArrayDestroy",
    "cuMipmappedArrayDestroy", "cuMemGetInfo_v2", "cuDeviceTotalMem_v2", "cuCtxSynchronize",
    "cuMemcpyAtoH_v2", "cuMemcpyDtoH_v2", "cuMemcpyHtoA_v2", "cuMemcpyHtoD_v2",
    "cuMemAllocAsync", "cuMemAllocFromPoolAsync", "cuMemFreeAsync", "cuGraphLaunch", "cuMemCreate", "cuMemRelease", "cuStreamSynchronize",
    "cuEventSynchronize", "cuStreamDestroy_v2", "cuMemAllocHost_v2", "cuMemHostAlloc", "cuMemFreeHost", nullptr};

extern "C" __attribute__((visibility("default"))) const char* const* gemhook_hooked_symbols(size_t* count) {
  if (count) *count = sizeof(kHookedNames) / sizeof(kHookedNames[0]) - 1;
  return kHookedNames;
}

// dlsym interposer (hook.cpp:109-159): anything that is not a hooked driver symbol goes to the libc dlsym.
// The pass-through is a chain of sibling calls (dlsym -> gh_true_dlsym -> libc dlsym, all `jmp` at -O2, checked in
// the disassembly), so glibc still sees the ORIGINAL caller's return address and dlsym(RTLD_NEXT, ...) issued by
// other libraries keeps its meaning -- the reference calls the real dlsym from inside its own (hook.cpp:80-84,
// 158), which re-anchors RTLD_NEXT at the hook library.
__attribute__((no_sanitize("thread", "address", "undefined"))) void* dlsym(void* handle, const char* symbol) {
  if (symbol && symbol[0] == 'c' && symbol[1] == 'u') {
    GH_COUNT(dlsym);
    if (!strcmp(symbol, "cuGetProcAddress_v2")) return (void*)&cuGetProcAddress_v2;
    if (!strcmp(symbol, "cuGetProcAddress")) return (void*)&cuGetProcAddress;
    for (size_t i = 0; i < kNumHooks; i++)
      if (!strcmp(symbol, kHooks[i].exported)) return kHooks[i].fn;
    // exported twins
    static const struct { const char* n; void* f; } twins[] = {
        {"cuLaunchKernel_ptsz", (void*)&cuLaunchKernel_ptsz},
        {"cuLaunchCooperativeKernel_ptsz", (void*)&cuLaunchCooperativeKernel_ptsz},
        {"cuLaunchKernelEx_ptsz", (void*)&cuLaunchKernelEx_ptsz},
        {"cuMemcpyAtoH_v2_ptds", (void*)&cuMemcpyAtoH_v2_ptds},
        {"cuMemcpyDtoH_v2_ptds", (void*)&cuMemcpyDtoH_v2_ptds},
        {"cuMemcpyHtoA_v2_ptds", (void*)&cuMemcpyHtoA_v2_ptds},
        {"cuMemcpyHtoD_v2_ptds", (void*)&cuMemcpyHtoD_v2_ptds},
        {"cuMemAllocAsync_ptsz", (void*)&cuMemAllocAsync_ptsz},
        {"cuMemAllocFromPoolAsync_ptsz", (void*)&cuMemAllocFromPoolAsync_ptsz},
        {"cuMemFreeAsync_ptsz", (void*)&cuMemFreeAsync_ptsz},
        {"cuGraphLaunch_ptsz", (void*)&cuGraphLaunch_ptsz},
        {"cuStreamSynchronize_ptsz", (void*)&cuStreamSynchronize_ptsz}};
    for (size_t i = 0; i < sizeof(twins) / sizeof(twins[0]); i++)
      if (!strcmp(symbol, twins[i].n)) return twins[i].f;
  }
  return gh_true_dlsym(handle, symbol);
}

static void swap_in_hook(const char* symbol, void** pfn, int cudaVersion, cuuint64_t flags) {
  if (!symbol || !pfn || !*pfn) return;
  if (!strcmp(symbol, "cuGetProcAddress")) {
    *pfn = cudaVersion >= 12000 ? (void*)&cuGetProcAddress_v2 : (void*)&cuGetProcAddress;
    return;
  }
  for (size_t i = 0; i < kNumHooks; i++) {
    if (!strcmp(symbol, kHooks[i].base)) {
      bool per_thread = (flags & CU_GET_PROC_ADDRESS_PER_THREAD_DEFAULT_STREAM) != 0;
      *pfn = (per_thread && kHooks[i].fn_pt) ? kHooks[i].fn_pt : kHooks[i].fn;
      return;
    }
  }
}

GH_HOOK cuGetProcAddress_v2(const char* symbol, void** pfn, int cudaVersion, cuuint64_t flags,
                            CUdriverProcAddressQueryResult* status) {
  typedef CUresult(CUDAAPI * fn_t)(const char*, void**, int, cuuint64_t, CUdriverProcAddressQueryResult*);
  GH_COUNT(cuGetProcAddress);
  if (gh_driver_init() != 0 || !gh_real.gpa_v2) return CUDA_ERROR_NOT_INITIALIZED;
  CUresult r = ((fn_t)gh_real.gpa_v2)(symbol, pfn, cudaVersion, flags, status);
  if (r == CUDA_SUCCESS) swap_in_hook(symbol, pfn, cudaVersion, flags);
  return r;
}

GH_HOOK cuGetProcAddress(const char* symbol, void** pfn, int cudaVersion, cuuint64_t flags) {
  typedef CUresult(CUDAAPI * fn_t)(const char*, void**, int, cuuint64_t);
  GH_COUNT(cuGetProcAddress);
  if (gh_driver_init() != 0 || !gh_real.gpa_legacy) return CUDA_ERROR_NOT_INITIALIZED;
  CUresult r = ((fn_t)gh_real.gpa_legacy)(symbol, pfn, cudaVersion, flags);
  if (r == CUDA_SUCCESS) swap_in_hook(symbol, pfn, cudaVersion, flags);
  return r;
}
