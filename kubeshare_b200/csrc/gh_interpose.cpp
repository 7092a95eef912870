// gh_interpose.cpp -- the hook surface: every symbol the reference libgemhook.so.1 interposes
// (reference hook.cpp:109-159 dlsym, :875-980 cuGetProcAddress, :1005-1062 wrappers, :857-872 mem-info),
// plus cuGetProcAddress_v2 -- without which a CUDA >= 12 runtime never sees a hook -- the _ptsz/_ptds
// twins, cuLaunchKernelEx and a correctly exported cuMemcpyDtoH_v2.
//
// Launch fast path = one load of gh_gate_open + one relaxed add + the indirect call into the driver.
#include <string.h>

#include "gh_internal.h"

#define GH_HOOK extern "C" __attribute__((visibility("default"))) CUresult CUDAAPI

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
static inline __attribute__((always_inline)) void launch_gate(CUstream hStream) {
  if (__builtin_expect(__atomic_load_n(&gh_gate_open, __ATOMIC_RELAXED) != 0, 1)) {
    uint64_t n = __atomic_add_fetch(&gh_launch_count, 1, __ATOMIC_RELAXED);
    if (__builtin_expect(((uint32_t)n & gh_seg_mask) == 0, 0)) gh_segment_tick(hStream);
  } else {
    gh_launch_slow(hStream);
    __atomic_add_fetch(&gh_launch_count, 1, __ATOMIC_RELAXED);
  }
}

typedef CUresult(CUDAAPI* launch_fn)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                     CUstream, void**, void**);
typedef CUresult(CUDAAPI* coop_fn)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                   CUstream, void**);
typedef CUresult(CUDAAPI* launchex_fn)(const CUlaunchConfig*, CUfunction, void**, void**);

GH_HOOK cuLaunchKernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                       unsigned shmem, CUstream hStream, void** params, void** extra) {
  launch_gate(hStream);
  return GH_REAL_CORE(cuLaunchKernel)(f, gx, gy, gz, bx, by, bz, shmem, hStream, params, extra);
}
GH_HOOK cuLaunchCooperativeKernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by,
                                  unsigned bz, unsigned shmem, CUstream hStream, void** params) {
  launch_gate(hStream);
  return GH_REAL_CORE(cuLaunchCooperativeKernel)(f, gx, gy, gz, bx, by, bz, shmem, hStream, params);
}

static void* p_launch_ptsz;
static void* p_coop_ptsz;
static void* p_launchex;
static void* p_launchex_ptsz;
#define LATE(slot, name) ((slot) ? (slot) : resolve_late(&(slot), name))

GH_HOOK cuLaunchKernel_ptsz(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                            unsigned shmem, CUstream hStream, void** params, void** extra) {
  launch_gate(hStream);
  return ((launch_fn)LATE(p_launch_ptsz, "cuLaunchKernel_ptsz"))(f, gx, gy, gz, bx, by, bz, shmem, hStream, params, extra);
}
GH_HOOK cuLaunchCooperativeKernel_ptsz(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by,
                                       unsigned bz, unsigned shmem, CUstream hStream, void** params) {
  launch_gate(hStream);
  return ((coop_fn)LATE(p_coop_ptsz, "cuLaunchCooperativeKernel_ptsz"))(f, gx, gy, gz, bx, by, bz, shmem, hStream, params);
}
GH_HOOK cuLaunchKernelEx(const CUlaunchConfig* config, CUfunction f, void** params, void** extra) {
  launch_gate(config ? config->hStream : nullptr);
  return ((launchex_fn)LATE(p_launchex, "cuLaunchKernelEx"))(config, f, params, extra);
}
GH_HOOK cuLaunchKernelEx_ptsz(const CUlaunchConfig* config, CUfunction f, void** params, void** extra) {
  launch_gate(config ? config->hStream : nullptr);
  return ((launchex_fn)LATE(p_launchex_ptsz, "cuLaunchKernelEx_ptsz"))(config, f, params, extra);
}

// ---- gpu_mem cap ----------------------------------------------------------------------------------------
GH_HOOK cuMemAlloc_v2(CUdeviceptr* dptr, size_t bytesize) {
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
  gh_live_get();  // managed memory is not accounted (hook.cpp:619-627)
  return GH_REAL_CORE(cuMemAllocManaged)(dptr, bytesize, flags);
}

GH_HOOK cuMemAllocPitch_v2(CUdeviceptr* dptr, size_t* pPitch, size_t WidthInBytes, size_t Height,
                           unsigned int ElementSizeBytes) {
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
  gh_mem_free_key((uint64_t)dptr);
  return GH_REAL_CORE(cuMemFree_v2)(dptr);
}

GH_HOOK cuArrayCreate_v2(CUarray* pHandle, const CUDA_ARRAY_DESCRIPTOR* d) {
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
  gh_live_get();  // not accounted (hook.cpp:682-694)
  return GH_REAL_CORE(cuMipmappedArrayCreate)(pHandle, d, levels);
}

GH_HOOK cuArrayDestroy(CUarray hArray) {
  gh_mem_free_key((uint64_t)(uintptr_t)hArray);
  return GH_REAL_CORE(cuArrayDestroy)(hArray);
}

GH_HOOK cuMipmappedArrayDestroy(CUmipmappedArray h) {
  gh_mem_free_key((uint64_t)(uintptr_t)h);
  return GH_REAL_CORE(cuMipmappedArrayDestroy)(h);
}

// mem-info virtualisation: the driver is never asked (hook.cpp:857-872)
extern "C" int gh_live_enabled(void);
GH_HOOK cuMemGetInfo_v2(size_t* free_b, size_t* total_b) {
  if (!gh_live_get() || !gh_live_enabled()) return GH_REAL_CORE(cuMemGetInfo_v2)(free_b, total_b);
  uint64_t f = 0, t = 0;
  gh_mem_info(&f, &t);
  if (free_b) *free_b = (size_t)f;
  if (total_b) *total_b = (size_t)t;
  return CUDA_SUCCESS;
}
GH_HOOK cuDeviceTotalMem_v2(size_t* bytes, CUdevice dev) {
  if (!gh_live_get() || !gh_live_enabled()) return GH_REAL_CORE(cuDeviceTotalMem_v2)(bytes, dev);
  uint64_t f = 0, t = 0;
  gh_mem_info(&f, &t);
  if (bytes) *bytes = (size_t)t;
  return CUDA_SUCCESS;
}

// ---- synchronising calls: burst end / window start (hook.cpp:696-722) -------------------------------------
GH_HOOK cuCtxSynchronize(void) {
  gh_host_sync_pre();
  CUresult r = GH_REAL_CORE(cuCtxSynchronize)();
  if (r == CUDA_SUCCESS) gh_host_sync_post();
  return r;
}

#define GH_SYNC_COPY(export_name, real_expr, params, args) \
  GH_HOOK export_name params {                             \
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

GH_SYNC_COPY(cuMemcpyAtoH_v2, GH_REAL_CORE(cuMemcpyAtoH_v2), (void* dst, CUarray src, size_t off, size_t n), (dst, src, off, n))
GH_SYNC_COPY(cuMemcpyDtoH_v2, GH_REAL_CORE(cuMemcpyDtoH_v2), (void* dst, CUdeviceptr src, size_t n), (dst, src, n))
GH_SYNC_COPY(cuMemcpyHtoA_v2, GH_REAL_CORE(cuMemcpyHtoA_v2), (CUarray dst, size_t off, const void* src, size_t n), (dst, off, src, n))
GH_SYNC_COPY(cuMemcpyHtoD_v2, GH_REAL_CORE(cuMemcpyHtoD_v2), (CUdeviceptr dst, const void* src, size_t n), (dst, src, n))
GH_SYNC_COPY(cuMemcpyAtoH_v2_ptds, (atoh_fn)LATE(p_atoh_ptds, "cuMemcpyAtoH_v2_ptds"), (void* dst, CUarray src, size_t off, size_t n), (dst, src, off, n))
GH_SYNC_COPY(cuMemcpyDtoH_v2_ptds, (dtoh_fn)LATE(p_dtoh_ptds, "cuMemcpyDtoH_v2_ptds"), (void* dst, CUdeviceptr src, size_t n), (dst, src, n))
GH_SYNC_COPY(cuMemcpyHtoA_v2_ptds, (htoa_fn)LATE(p_htoa_ptds, "cuMemcpyHtoA_v2_ptds"), (CUarray dst, size_t off, const void* src, size_t n), (dst, off, src, n))
GH_SYNC_COPY(cuMemcpyHtoD_v2_ptds, (htod_fn)LATE(p_htod_ptds, "cuMemcpyHtoD_v2_ptds"), (CUdeviceptr dst, const void* src, size_t n), (dst, src, n))

// ---- modern entry points the reference never saw ------------------------------------------
// Stream-ordered allocations are charged like cuMemAlloc; graph launches pass the token gate like a kernel
// launch; cuStreamSynchronize / cuEventSynchronize count as host syncs only with GEMHOOK_EXTRA_HOOKS=1 (the
// reference's burst detection knows cuCtxSynchronize and the four blocking copies only, hook.cpp:696-722).
typedef CUresult(CUDAAPI* allocasync_fn)(CUdeviceptr*, size_t, CUstream);
typedef CUresult(CUDAAPI* allocpool_fn)(CUdeviceptr*, size_t, CUmemoryPool, CUstream);
typedef CUresult(CUDAAPI* freeasync_fn)(CUdeviceptr, CUstream);
typedef CUresult(CUDAAPI* graphlaunch_fn)(CUgraphExec, CUstream);
typedef CUresult(CUDAAPI* streamsync_fn)(CUstream);
typedef CUresult(CUDAAPI* eventsync_fn)(CUevent);
static void *p_allocasync, *p_allocasync_pt, *p_allocpool, *p_allocpool_pt, *p_freeasync, *p_freeasync_pt;
static void *p_graphlaunch, *p_graphlaunch_pt, *p_streamsync, *p_streamsync_pt, *p_eventsync;

#define GH_ALLOC_ASYNC(name, slot, sym)                                        \
  GH_HOOK name(CUdeviceptr* dptr, size_t bytesize, CUstream hStream) {         \
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
  gh_mem_free_key((uint64_t)dptr);
  return ((freeasync_fn)LATE(p_freeasync, "cuMemFreeAsync"))(dptr, hStream);
}
GH_HOOK cuMemFreeAsync_ptsz(CUdeviceptr dptr, CUstream hStream) {
  gh_mem_free_key((uint64_t)dptr);
  return ((freeasync_fn)LATE(p_freeasync_pt, "cuMemFreeAsync_ptsz"))(dptr, hStream);
}
GH_HOOK cuGraphLaunch(CUgraphExec g, CUstream hStream) {
  launch_gate(hStream);
  return ((graphlaunch_fn)LATE(p_graphlaunch, "cuGraphLaunch"))(g, hStream);
}
GH_HOOK cuGraphLaunch_ptsz(CUgraphExec g, CUstream hStream) {
  launch_gate(hStream);
  return ((graphlaunch_fn)LATE(p_graphlaunch_pt, "cuGraphLaunch_ptsz"))(g, hStream);
}
// Virtual-memory-management allocations (PyTorch "expandable segments"): the physical handle is what consumes
// device memory, so cuMemCreate is charged and cuMemRelease gives it back; mapping / address reservation is free.
typedef CUresult(CUDAAPI* memcreate_fn)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
typedef CUresult(CUDAAPI* memrelease_fn)(CUmemGenericAllocationHandle);
static void *p_memcreate, *p_memrelease;
static inline uint64_t handle_key(CUmemGenericAllocationHandle h) { return (uint64_t)h ^ 0x8000000000000000ULL; }
GH_HOOK cuMemCreate(CUmemGenericAllocationHandle* handle, size_t size, const CUmemAllocationProp* prop, unsigned long long flags) {
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
  gh_mem_free_key(handle_key(handle));
  return ((memrelease_fn)LATE(p_memrelease, "cuMemRelease"))(handle);
}

static inline bool extra_syncs(void) { return gh_live_get() && gh_cfg.extra_hooks; }
GH_HOOK cuStreamSynchronize(CUstream hStream) {
  bool x = extra_syncs();
  if (x) gh_host_sync_pre();
  CUresult r = ((streamsync_fn)LATE(p_streamsync, "cuStreamSynchronize"))(hStream);
  if (x && r == CUDA_SUCCESS) gh_host_sync_post();
  return r;
}
GH_HOOK cuStreamSynchronize_ptsz(CUstream hStream) {
  bool x = extra_syncs();
  if (x) gh_host_sync_pre();
  CUresult r = ((streamsync_fn)LATE(p_streamsync_pt, "cuStreamSynchronize_ptsz"))(hStream);
  if (x && r == CUDA_SUCCESS) gh_host_sync_post();
  return r;
}
GH_HOOK cuEventSynchronize(CUevent ev) {
  bool x = extra_syncs();
  if (x) gh_host_sync_pre();
  CUresult r = ((eventsync_fn)LATE(p_eventsync, "cuEventSynchronize"))(ev);
  if (x && r == CUDA_SUCCESS) gh_host_sync_post();
  return r;
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
};
static const size_t kNumHooks = sizeof(kHooks) / sizeof(kHooks[0]);

static const char* const kHookedNames[] = {
    "dlsym", "cuGetProcAddress", "cuGetProcAddress_v2", "cuLaunchKernel", "cuLaunchCooperativeKernel",
    "cuLaunchKernelEx", "cuMemAlloc_v2", "cuMemAllocManaged", "cuMemAllocPitch_v2", "cuMemFree_v2",
    "cuArrayCreate_v2", "cuArray3DCreate_v2", "cuMipmappedArrayCreate", "cuArrayDestroy",
    "cuMipmappedArrayDestroy", "cuMemGetInfo_v2", "cuDeviceTotalMem_v2", "cuCtxSynchronize",
    "cuMemcpyAtoH_v2", "cuMemcpyDtoH_v2", "cuMemcpyHtoA_v2", "cuMemcpyHtoD_v2",
    "cuMemAllocAsync", "cuMemAllocFromPoolAsync", "cuMemFreeAsync", "cuGraphLaunch", "cuMemCreate", "cuMemRelease", "cuStreamSynchronize",
    "cuEventSynchronize", nullptr};

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
  if (gh_driver_init() != 0 || !gh_real.gpa_v2) return CUDA_ERROR_NOT_INITIALIZED;
  CUresult r = ((fn_t)gh_real.gpa_v2)(symbol, pfn, cudaVersion, flags, status);
  if (r == CUDA_SUCCESS) swap_in_hook(symbol, pfn, cudaVersion, flags);
  return r;
}

GH_HOOK cuGetProcAddress(const char* symbol, void** pfn, int cudaVersion, cuuint64_t flags) {
  typedef CUresult(CUDAAPI * fn_t)(const char*, void**, int, cuuint64_t);
  if (gh_driver_init() != 0 || !gh_real.gpa_legacy) return CUDA_ERROR_NOT_INITIALIZED;
  CUresult r = ((fn_t)gh_real.gpa_legacy)(symbol, pfn, cudaVersion, flags);
  if (r == CUDA_SUCCESS) swap_in_hook(symbol, pfn, cudaVersion, flags);
  return r;
}
