// gh_acct.cpp -- host side of the device accounting path (include/gemhook.h section 2d).
//
// Owns: the embedded sm_100a cubin (acct_kernels.cu), a NON_BLOCKING accounting stream (so our work
// never serialises with the application's legacy default stream), the device-resident record ring and
// running totals, and the mapped pinned totals page.  Driver API only -- the hook must not drag a
// second cudart into the application (the reference links cudart for five event calls, reference
// hook.cpp:482-491, 543, 754).
//
// There is NO CPU fallback here: without a device (or with the wrong architecture) create() fails and
// says why.
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "gh_internal.h"

extern "C" {
extern const unsigned char _binary_acct_kernels_cubin_start[];
extern const unsigned char _binary_acct_kernels_cubin_end[];
}

namespace {

struct totals_page {  // mirrors gemhook_totals_page in acct_kernels.cu
  volatile uint64_t epoch;
  volatile uint64_t nslots;
  volatile uint64_t mem_slot;
  uint64_t reserved;
  volatile uint64_t buf[2][GEMHOOK_MAX_SLOTS * 3];
  volatile uint64_t mem[2][2];
};
struct mem_mirror {  // mirrors gemhook_mem_mirror (kernel parameter, by value)
  uint64_t slot, used, limit;
};

#ifndef GEMHOOK_UNROLL
#define GEMHOOK_UNROLL 8
#endif
const unsigned TILE_RECORDS = 32u * GEMHOOK_UNROLL;  // records per warp iteration (32 lanes x GEMHOOK_UNROLL)
const unsigned STAGED_MIN_SLOTS = 22;                // above this many client slots the TMA-staged kernel runs (measured crossover)
const size_t SMALL_N = 512;                          // up to here one warp does everything (gemhook_acct_reduce_small);
                                                     // measured: 10.5 vs 12.7 us at 2-64 records, break-even near 1024
// shared memory per warp: (nslots + 1) rows of 32 16-byte cells (the extra row swallows out-of-range slots) + the
// warp's u64 accumulators
#ifndef GEMHOOK_COLS
#define GEMHOOK_COLS 32
#endif
inline unsigned warp_smem(unsigned nslots) { return (nslots + 1u) * GEMHOOK_COLS * 16u + nslots * 24u; }

const char* cu_err(CUresult r) {
  const char* s = nullptr;
  if (gh_real.cuGetErrorString) GH_CALL(cuGetErrorString, r, &s);
  return s ? s : "unknown CUDA error";
}

}  // namespace

#define CU_TRY(expr)                                                             \
  do {                                                                           \
    CUresult _r = (expr);                                                        \
    if (_r != CUDA_SUCCESS) {                                                    \
      gh_set_error("%s failed: %d (%s) at %s:%d", #expr, (int)_r, cu_err(_r), __FILE__, __LINE__); \
      return -1;                                                                 \
    }                                                                            \
  } while (0)

struct gemhook_acct {
  CUcontext ctx = nullptr;
  CUmodule mod = nullptr;
  CUfunction f_reduce = nullptr, f_small = nullptr, f_clear = nullptr, f_peek = nullptr, f_staged = nullptr;
  CUstream stream = nullptr;
  CUevent ev0 = nullptr, ev1 = nullptr;
  CUdeviceptr d_ring = 0, d_totals = 0, d_ticket = 0, d_page = 0, d_mem = 0;
  totals_page* page = nullptr;
  size_t ring_cap = 0;
  uint32_t nslots = 0;
  unsigned warps = 8, smem_bytes = 0, small_smem = 0, max_blocks = 0, flush_every = 8000;
  unsigned staged_cols = 32;
  const unsigned stage_rows = GEMHOOK_UNROLL;  // 32-record rows per ring buffer (8 -> 4 KB; 2 KB buffers measured slower)
  unsigned stages = 0;  // > 0: the TMA-staged kernel (many client slots) with this many 4 KB buffers per warp
  int sm_count = 0;
  mem_mirror mm = {0, 0, 0};
  bool small_enabled = true;
  std::atomic<uint64_t> kernel_launches{0};
  std::atomic<uint64_t> reduce_launches{0};  // each publishes one epoch of the totals page
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
};

static int acct_init(gemhook_acct* a, uint32_t nslots, size_t ring_cap) {
  if (gh_driver_init() != 0) return -1;
  if (nslots == 0 || nslots > GEMHOOK_MAX_SLOTS) {
    gh_set_error("nslots %u out of range 1..%d", nslots, GEMHOOK_MAX_SLOTS);
    return -1;
  }
  CU_TRY(GH_CALL(cuCtxGetCurrent, &a->ctx));
  if (!a->ctx) {
    gh_set_error("no CUDA context is current on the calling thread");
    return -1;
  }
  CUdevice dev;
  CU_TRY(GH_CALL(cuCtxGetDevice, &dev));
  int major = 0, minor = 0;
  CU_TRY(GH_CALL(cuDeviceGetAttribute, &major, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR, dev));
  CU_TRY(GH_CALL(cuDeviceGetAttribute, &minor, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR, dev));
  CU_TRY(GH_CALL(cuDeviceGetAttribute, &a->sm_count, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev));
  if (major != 10) {
    gh_set_error("accounting kernels are built for sm_100a only; device is sm_%d%d", major, minor);
    return -1;
  }
  CU_TRY(GH_CALL(cuModuleLoadData, &a->mod, (const void*)_binary_acct_kernels_cubin_start));
  // Launch shape.  Bins cost warp_smem(nslots) per warp (8.9 KB at 16 slots, 34.8 KB at 64); the kernel keeps two tiles of
  // UNROLL 16-byte loads per lane in flight per warp (register double buffer, 128 registers at UNROLL 8), so the
  // register file allows 16 warps per SM and a handful of warps already covers the HBM latency x bandwidth product
  // (~36 KB in flight per SM).  Warps per SM = min(16, what fits into ~216 KB of shared memory), split into one or two
  // blocks; the grid is one full wave (a multiple of the SM count), grid-stride over 256-record warp tiles.
  const unsigned per_warp = warp_smem(nslots);
  unsigned total = (216u * 1024u) / per_warp;
  if (total > 16u) total = 16u;
  if (total < 1u) total = 1u;
  unsigned blocks_per_sm = total > 8u ? 2u : 1u;
  a->warps = total / blocks_per_sm;
  if (const char* e = getenv("GEMHOOK_ACCT_WARPS")) {
    unsigned w = (unsigned)atoi(e);
    if (w >= 1u && w <= 8u && w * per_warp <= 227u * 1024u) a->warps = w;
  }
  if (const char* e = getenv("GEMHOOK_ACCT_BLOCKS_PER_SM")) {
    unsigned b = (unsigned)atoi(e);
    if (b >= 1u && b <= 8u) blocks_per_sm = b;
  }
  if (const char* e = getenv("GEMHOOK_ACCT_FLUSH_EVERY")) {  // tests: force the in-kernel bin flush (default: every 8000 tiles)
    unsigned f = (unsigned)atoi(e);
    if (f >= 1u && f <= 8000u) a->flush_every = f;
  }
  if (const char* e = getenv("GEMHOOK_ACCT_SMALL")) a->small_enabled = atoi(e) != 0;
  // Many client slots: the bins leave room for only a few warps, and a few warps with register-staged loads cannot keep
  // HBM busy.  gemhook_acct_reduce_staged feeds the same accumulation from per-warp rings of 4 KB buffers filled by
  // cp.async.bulk, so the bytes in flight are set by the ring: pick the largest warp count (one block per SM) that still
  // leaves every warp >= 2 buffers and the SM >= 48 KB in flight (the HBM latency x bandwidth product is ~36 KB per SM).
  const unsigned max_warps = 8;
  const unsigned SMEM_MAX = 227u * 1024u, STG = 32u * a->stage_rows * 16u + 8u;
  bool staged = nslots > STAGED_MIN_SLOTS;
  if (const char* e = getenv("GEMHOOK_ACCT_STAGED")) staged = atoi(e) != 0;
  // 32 columns while eight warps with two buffers each fit (up to 37 slots), 16 columns beyond: half the bins, twice the
  // warps (measured at 48 / 64 slots: 0.99 / 0.88 of the roofline with 32 columns, 1.00 / 0.95 with 16)
  unsigned per_warp_staged = per_warp;
  if (8u * (per_warp + 2u * STG) + 16u > SMEM_MAX) a->staged_cols = 16;
  if (const char* e = getenv("GEMHOOK_ACCT_STAGED_COLS")) {  // sweeps
    if (atoi(e) == 16 || atoi(e) == 32) a->staged_cols = (unsigned)atoi(e);
  }
  if (a->staged_cols == 16) per_warp_staged = (nslots + 1u) * 16u * 16u + nslots * 24u;
  if (staged) {
    const unsigned per_warp = per_warp_staged;  // (shadows the register-staged kernel's figure inside this block)
    unsigned best_w = 0, best_s = 0;
    for (unsigned w = max_warps; w >= 1 && !best_w; w--) {
      if (w * per_warp + 16u + w * 2u * STG > SMEM_MAX) continue;
      unsigned s_ = (SMEM_MAX - w * per_warp - 16u) / (w * STG);
      if (s_ > 3u) s_ = 3u;  // measured at 20-22 slots, eight warps: 2 / 3 / 6 buffers -> 1.00 / 1.01 / 0.97 of the roofline
      if (w * s_ >= 12u || w == 1u) best_w = w, best_s = s_;
    }
    if (const char* e = getenv("GEMHOOK_ACCT_WARPS")) {
      unsigned w = (unsigned)atoi(e);
      if (w >= 1u && w <= max_warps && w * per_warp + 16u + w * 2u * STG <= SMEM_MAX) {
        best_w = w;
        best_s = (SMEM_MAX - w * per_warp - 16u) / (w * STG);
        if (best_s > 8u) best_s = 8u;
      }
    }
    if (const char* e = getenv("GEMHOOK_ACCT_STAGES")) {
      unsigned s_ = (unsigned)atoi(e);
      if (s_ >= 1u && best_w && best_w * per_warp + 16u + best_w * s_ * STG <= SMEM_MAX) best_s = s_;
    }
    if (best_w && best_s) {
      a->warps = best_w;
      a->stages = best_s;
      blocks_per_sm = 1;
    }
  }
  CU_TRY(GH_CALL(cuModuleGetFunction, &a->f_reduce, a->mod, "gemhook_acct_reduce"));
  CU_TRY(GH_CALL(cuModuleGetFunction, &a->f_staged, a->mod,
                 a->staged_cols == 16 ? "gemhook_acct_reduce_staged_c16" : "gemhook_acct_reduce_staged"));
  CU_TRY(GH_CALL(cuModuleGetFunction, &a->f_small, a->mod, "gemhook_acct_reduce_small"));
  CU_TRY(GH_CALL(cuModuleGetFunction, &a->f_clear, a->mod, "gemhook_acct_clear"));
  CU_TRY(GH_CALL(cuModuleGetFunction, &a->f_peek, a->mod, "gemhook_peek_pool"));
  CU_TRY(GH_CALL(cuStreamCreate, &a->stream, CU_STREAM_NON_BLOCKING));
  CU_TRY(GH_CALL(cuEventCreate, &a->ev0, CU_EVENT_DEFAULT));
  CU_TRY(GH_CALL(cuEventCreate, &a->ev1, CU_EVENT_DEFAULT));

  a->nslots = nslots;
  a->smem_bytes = a->warps * per_warp;
  if (a->stages) a->smem_bytes = ((a->warps * per_warp_staged + 15u) & ~15u) + a->warps * a->stages * STG;
  a->small_smem = (nslots + 1u) * GEMHOOK_COLS * 16u;
  CUfunction f_big = a->stages ? a->f_staged : a->f_reduce;
  if (a->smem_bytes > 48u * 1024u)
    CU_TRY(GH_CALL(cuFuncSetAttribute, f_big, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)a->smem_bytes));
  int per_sm = 0;
  CU_TRY(GH_CALL(cuOccupancyMaxActiveBlocksPerMultiprocessor, &per_sm, f_big, (int)(a->warps * 32u), (size_t)a->smem_bytes));
  if (per_sm < 1) per_sm = 1;
  if ((unsigned)per_sm > blocks_per_sm) per_sm = (int)blocks_per_sm;
  a->max_blocks = (unsigned)(per_sm * a->sm_count);  // one full wave: a multiple of the SM count

  a->ring_cap = ring_cap ? ring_cap : (1u << 16);
  CU_TRY(GH_CALL(cuMemAlloc_v2, &a->d_ring, a->ring_cap * sizeof(gemhook_record)));
  size_t tot_bytes = ((size_t)nslots * 3 + 1) * sizeof(uint64_t);
  CU_TRY(GH_CALL(cuMemAlloc_v2, &a->d_totals, tot_bytes));
  CU_TRY(GH_CALL(cuMemAlloc_v2, &a->d_ticket, 256));
  CU_TRY(GH_CALL(cuMemAlloc_v2, &a->d_mem, 256));  // device-resident mirror of the pod's gpu_mem counter + peek scratch
  CU_TRY(GH_CALL(cuMemsetD8Async, a->d_totals, 0, tot_bytes, a->stream));
  CU_TRY(GH_CALL(cuMemsetD8Async, a->d_ticket, 0, 256, a->stream));
  CU_TRY(GH_CALL(cuMemsetD8Async, a->d_mem, 0, 256, a->stream));
  void* hp = nullptr;
  CU_TRY(GH_CALL(cuMemHostAlloc, &hp, sizeof(totals_page), CU_MEMHOSTALLOC_PORTABLE | CU_MEMHOSTALLOC_DEVICEMAP));
  memset(hp, 0, sizeof(totals_page));
  a->page = (totals_page*)hp;
  CU_TRY(GH_CALL(cuMemHostGetDevicePointer_v2, &a->d_page, hp, 0));
  CU_TRY(GH_CALL(cuStreamSynchronize, a->stream));
  GH_INFO("acct: %d SMs, nslots %u, %u warps/block, %u B smem, wave %u blocks (%d per SM), %u TMA stages per warp", a->sm_count,
          nslots, a->warps, a->smem_bytes, a->max_blocks, per_sm, a->stages);
  return 0;
}

GH_EXPORT gemhook_acct* gemhook_acct_create(uint32_t nslots, size_t ring_capacity_records) {
  gemhook_acct* a = new gemhook_acct();
  if (acct_init(a, nslots, ring_capacity_records) != 0) {
    gemhook_acct_destroy(a);
    return nullptr;
  }
  return a;
}

GH_EXPORT void gemhook_acct_destroy(gemhook_acct* a) {
  if (!a) return;
  if (a->stream) GH_CALL(cuStreamSynchronize, a->stream);
  if (a->d_ring) GH_CALL(cuMemFree_v2, a->d_ring);
  if (a->d_totals) GH_CALL(cuMemFree_v2, a->d_totals);
  if (a->d_ticket) GH_CALL(cuMemFree_v2, a->d_ticket);
  if (a->d_mem) GH_CALL(cuMemFree_v2, a->d_mem);
  if (a->page) GH_CALL(cuMemFreeHost, (void*)a->page);
  if (a->ev0) GH_CALL(cuEventDestroy_v2, a->ev0);
  if (a->ev1) GH_CALL(cuEventDestroy_v2, a->ev1);
  if (a->stream) GH_CALL(cuStreamDestroy_v2, a->stream);
  if (a->mod) GH_CALL(cuModuleUnload, a->mod);
  delete a;
}

GH_EXPORT void gemhook_acct_launch_shape(const gemhook_acct* a, uint32_t out[6]) {
  out[0] = a->warps;
  out[1] = a->max_blocks;
  out[2] = a->smem_bytes;
  out[3] = a->stages;
  out[4] = a->stages ? a->staged_cols : (unsigned)GEMHOOK_COLS;
  out[5] = a->small_smem;
}

GH_EXPORT uint32_t gemhook_acct_grid_for(const gemhook_acct* a, size_t n) {
  if (a->small_enabled && n <= SMALL_N) return 1;
  size_t per_block = (size_t)a->warps * (a->stages ? 32u * a->stage_rows : TILE_RECORDS);
  size_t want = (n + per_block - 1) / per_block;
  if (want < 1) want = 1;
  if (want > a->max_blocks) want = a->max_blocks;
  return (uint32_t)want;
}

// launch the reduction over n records at device address d_rec (stream-ordered, no host sync)
static int launch_reduce(gemhook_acct* a, CUdeviceptr d_rec, size_t n) {
  unsigned ns = a->nslots;
  if (a->small_enabled && n <= SMALL_N) {  // the live hook's regime: one warp, no ticket
    unsigned nn = (unsigned)n;
    void* args[] = {&d_rec, &nn, &ns, &a->d_totals, &a->d_page, &a->mm, &a->d_mem};
    CU_TRY(GH_CALL(cuLaunchKernel, a->f_small, 1, 1, 1, 32, 1, 1, a->small_smem, a->stream, args, nullptr));
  } else {
    unsigned long long nn = n;
    void* args[] = {&d_rec, &nn, &ns, &a->d_totals, &a->d_ticket, &a->d_page, &a->mm, &a->d_mem, &a->flush_every, &a->stages};
    unsigned grid = gemhook_acct_grid_for(a, n);  // (the staged kernel takes one more parameter: the ring depth)
    CU_TRY(GH_CALL(cuLaunchKernel, a->stages ? a->f_staged : a->f_reduce, grid, 1, 1, a->warps * 32u, 1, 1, a->smem_bytes, a->stream,
                   args, nullptr));
  }
  a->kernel_launches.fetch_add(1, std::memory_order_relaxed);
  a->reduce_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}
uint64_t gh_acct_reduce_launches(const gemhook_acct* a) { return a ? a->reduce_launches.load(std::memory_order_relaxed) : 0; }

// ---- gpu_mem mirror -----------------------------------------------------------------------------------------------
// The authoritative counter is the CAS word in the shared-pinned credit pool (gh_pool.cpp SlotShared); the process
// hands its pod's current (used, limit) to every reduce launch, whose publish step leaves them in device memory and
// in the totals page next to the SM-time of the same epoch.
GH_EXPORT void gemhook_acct_set_mem(gemhook_acct* a, uint32_t slot, uint64_t used, uint64_t limit) {
  if (!a) return;
  pthread_mutex_lock(&a->mu);
  a->mm.slot = slot;
  a->mm.used = used;
  a->mm.limit = limit;
  pthread_mutex_unlock(&a->mu);
}
// from_device = 0: the copy in the totals page (no CUDA call); 1: copied back from the device-resident words
GH_EXPORT int gemhook_acct_read_mem(gemhook_acct* a, int from_device, uint64_t* slot, uint64_t* used, uint64_t* limit, uint64_t* epoch) {
  if (!a) return -1;
  uint64_t w[4] = {0, 0, 0, 0};
  if (from_device) {
    pthread_mutex_lock(&a->mu);
    CUresult r = GH_CALL(cuStreamSynchronize, a->stream);
    if (r == CUDA_SUCCESS) r = GH_CALL(cuMemcpyDtoH_v2, w, a->d_mem, sizeof(w));
    pthread_mutex_unlock(&a->mu);
    if (r != CUDA_SUCCESS) {
      gh_set_error("reading the device mem mirror failed: %d", (int)r);
      return -1;
    }
    if (used) *used = w[0];
    if (limit) *limit = w[1];
    if (slot) *slot = w[2];
    if (epoch) *epoch = w[3];
    return 0;
  }
  for (int tries = 0; tries < 1000000; tries++) {
    uint64_t e1 = a->page->epoch;
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    uint64_t u = a->page->mem[e1 & 1][0], l = a->page->mem[e1 & 1][1], sl = a->page->mem_slot;
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    if (a->page->epoch == e1) {
      if (used) *used = u;
      if (limit) *limit = l;
      if (slot) *slot = sl;
      if (epoch) *epoch = e1;
      return 0;
    }
  }
  return -1;
}
// "shared-pinned" check: page-lock + device-map the 4 KiB page(s) holding `host_words` (no-op if the caller already
// registered them -- the hook registers the pool's counter pages), read four u64 words THROUGH THE DEVICE and return them.
GH_EXPORT int gemhook_acct_peek_host_words(gemhook_acct* a, const void* host_words, uint64_t out[4]) {
  if (!a || !host_words) return -1;
  uintptr_t lo = (uintptr_t)host_words & ~(uintptr_t)4095, hi = ((uintptr_t)host_words + 32 + 4095) & ~(uintptr_t)4095;
  CUresult reg = GH_CALL(cuMemHostRegister_v2, (void*)lo, hi - lo, CU_MEMHOSTREGISTER_PORTABLE | CU_MEMHOSTREGISTER_DEVICEMAP);
  bool mine = reg == CUDA_SUCCESS;
  if (!mine && reg != CUDA_ERROR_HOST_MEMORY_ALREADY_REGISTERED) {
    gh_set_error("cuMemHostRegister of the pool words failed: %d (%s)", (int)reg, cu_err(reg));
    return -1;
  }
  CUdeviceptr d = 0;
  int rc = -1;
  pthread_mutex_lock(&a->mu);
  if (GH_CALL(cuMemHostGetDevicePointer_v2, &d, (void*)host_words, 0) == CUDA_SUCCESS) {
    CUdeviceptr d_out = a->d_mem + 64;
    void* args[] = {&d, &d_out};
    if (GH_CALL(cuLaunchKernel, a->f_peek, 1, 1, 1, 32, 1, 1, 0, a->stream, args, nullptr) == CUDA_SUCCESS &&
        GH_CALL(cuStreamSynchronize, a->stream) == CUDA_SUCCESS && GH_CALL(cuMemcpyDtoH_v2, out, d_out, 32) == CUDA_SUCCESS) {
      a->kernel_launches.fetch_add(1, std::memory_order_relaxed);
      rc = 0;
    }
  }
  pthread_mutex_unlock(&a->mu);
  if (rc != 0) gh_set_error("device read of the pool words failed");
  if (mine) GH_CALL(cuMemHostUnregister, (void*)lo);
  return rc;
}

GH_EXPORT int gemhook_acct_reduce_device(gemhook_acct* a, uint64_t d_records, size_t n, float* kernel_ms_out) {
  if (!a) return -1;
  if (d_records & 15u) {
    gh_set_error("records must be 16-byte aligned");
    return -1;
  }
  pthread_mutex_lock(&a->mu);
  int rc = 0;
  if (kernel_ms_out) rc = (GH_CALL(cuEventRecord, a->ev0, a->stream) == CUDA_SUCCESS) ? 0 : -1;
  if (rc == 0 && n) rc = launch_reduce(a, (CUdeviceptr)d_records, n);
  if (rc == 0 && kernel_ms_out) {
    if (GH_CALL(cuEventRecord, a->ev1, a->stream) != CUDA_SUCCESS || GH_CALL(cuEventSynchronize, a->ev1) != CUDA_SUCCESS ||
        GH_CALL(cuEventElapsedTime, kernel_ms_out, a->ev0, a->ev1) != CUDA_SUCCESS) {
      gh_set_error("event timing of the reduce kernel failed");
      rc = -1;
    }
  }
  pthread_mutex_unlock(&a->mu);
  return rc;
}

static int read_page(gemhook_acct* a, uint64_t* totals_out, uint64_t* epoch_out) {
  // The device writes buf[(e+1) & 1], fences, then stores epoch = e+1.  A copy of buf[e & 1] is intact only if the
  // epoch did not move at all while it was taken: once the epoch reads e+1, kernel e+2 may already be filling
  // buf[(e+2) & 1] -- the very buffer being copied.  Anything else -> retry (a flush is microseconds apart at worst).
  for (int tries = 0; tries < 1000000; tries++) {
    uint64_t e1 = a->page->epoch;
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    if (totals_out) {
      if (e1 == 0) memset(totals_out, 0, sizeof(uint64_t) * a->nslots * 3);
      else
        for (uint32_t i = 0; i < a->nslots * 3; i++) totals_out[i] = a->page->buf[e1 & 1][i];
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    uint64_t e2 = a->page->epoch;
    if (e2 == e1) {
      if (epoch_out) *epoch_out = e1;
      return 0;
    }
  }
  gh_set_error("totals page never became stable");
  return -1;
}

GH_EXPORT int gemhook_acct_read_totals(gemhook_acct* a, uint64_t* totals_out, uint64_t* epoch_out) {
  if (!a) return -1;
  return read_page(a, totals_out, epoch_out);
}

GH_EXPORT int gemhook_acct_reduce_host(gemhook_acct* a, const gemhook_record* records, size_t n, uint64_t* totals_out) {
  if (!a) return -1;
  pthread_mutex_lock(&a->mu);
  int rc = 0;
  size_t done = 0;
  while (done < n && rc == 0) {
    size_t chunk = n - done < a->ring_cap ? n - done : a->ring_cap;
    // stream order makes the ring safe to reuse: copy k+1 cannot start before kernel k finished
    if (GH_CALL(cuMemcpyHtoDAsync_v2, a->d_ring, records + done, chunk * sizeof(gemhook_record), a->stream) != CUDA_SUCCESS) {
      gh_set_error("H2D copy of %zu records failed", chunk);
      rc = -1;
      break;
    }
    rc = launch_reduce(a, a->d_ring, chunk);
    done += chunk;
  }
  if (rc == 0 && GH_CALL(cuStreamSynchronize, a->stream) != CUDA_SUCCESS) {
    gh_set_error("accounting stream failed");
    rc = -1;
  }
  if (rc == 0 && totals_out) {
    rc = read_page(a, totals_out, nullptr);
  }
  pthread_mutex_unlock(&a->mu);
  return rc;
}

GH_EXPORT int gemhook_acct_sync(gemhook_acct* a) {
  if (!a) return -1;
  CU_TRY(GH_CALL(cuStreamSynchronize, a->stream));
  return 0;
}

GH_EXPORT int gemhook_acct_reset(gemhook_acct* a) {
  if (!a) return -1;
  pthread_mutex_lock(&a->mu);
  unsigned cnt = a->nslots * 3;  // the publish counter (last word) keeps counting
  void* args[] = {&a->d_totals, &cnt};
  CUresult r = GH_CALL(cuLaunchKernel, a->f_clear, 1, 1, 1, 256, 1, 1, 0, a->stream, args, nullptr);
  if (r == CUDA_SUCCESS) a->kernel_launches.fetch_add(1, std::memory_order_relaxed);
  if (r == CUDA_SUCCESS) r = GH_CALL(cuStreamSynchronize, a->stream);
  for (uint32_t i = 0; i < a->nslots * 3; i++) a->page->buf[0][i] = a->page->buf[1][i] = 0;
  // (the page itself is rewritten by the next launch; until then the host copy must not show stale sums)
  pthread_mutex_unlock(&a->mu);
  if (r != CUDA_SUCCESS) {
    gh_set_error("reset failed: %d", (int)r);
    return -1;
  }
  return 0;
}

GH_EXPORT uint64_t gemhook_acct_kernel_launches(const gemhook_acct* a) {
  return a ? a->kernel_launches.load(std::memory_order_relaxed) : 0;
}
GH_EXPORT uint64_t gemhook_acct_stream(const gemhook_acct* a) { return a ? (uint64_t)(uintptr_t)a->stream : 0; }

// internal helpers for the live hook (gh_hook.cpp)
int gh_acct_push_async(gemhook_acct* a, const gemhook_record* pinned_records, size_t n) {
  // records live in pinned host memory owned by the caller and stay valid until the next sync
  pthread_mutex_lock(&a->mu);
  int rc = 0;
  if (n > a->ring_cap) n = a->ring_cap;
  if (GH_CALL(cuMemcpyHtoDAsync_v2, a->d_ring, pinned_records, n * sizeof(gemhook_record), a->stream) != CUDA_SUCCESS) rc = -1;
  if (rc == 0) rc = launch_reduce(a, a->d_ring, n);
  pthread_mutex_unlock(&a->mu);
  return rc;
}
