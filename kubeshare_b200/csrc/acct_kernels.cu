// acct_kernels.cu -- sm_100a device code of the gemhook accounting path.
//
// The reference hook has no device code at all: its only GPU-time measurement is one event pair per
// token resolved on the host (reference Gemini/src/hook.cpp:456-502).  The B200-native hook stamps
// launch segments with CUDA events (csrc/gh_hook.cpp) and reduces the resulting 16-byte launch records
// on the device, so the per-client running totals live next to the ring in HBM and only one small
// snapshot per launch crosses to the mapped pinned totals page the host reads without a CUDA call.
//
// Record (16 B, one uint4, see include/gemhook.h gemhook_record):
//     x = slot            client slot in the credit pool; slot >= nslots is ignored
//     y = launches        kernel launches covered by the record
//     z,w = elapsed_ns    u64 little-endian: SM-time of the segment in nanoseconds
// Output: per slot {sum elapsed_ns, sum launches, record count}, all u64 -> integer sums are
// order-independent, so parity with the CPU oracle is bit-exact.
//
// Roofline: pure streaming read, 16 B per record, O(nslots) bytes written per block -> HBM-bound.
// Design:
//   * every lane loads whole records with 128-bit ld.global.nc.L1::no_allocate (a warp covers 512 contiguous
//     bytes per load); the loads of tile k+1 are issued BEFORE tile k is accumulated (register double buffer), so
//     every warp keeps UNROLL 16-byte loads per lane in flight all the time -- what decides the bandwidth once the
//     bin tables of many slots leave room for only a few warps per SM;
//   * privatised accumulation without atomics: each warp owns bins[slot][lane] in shared memory, one 16-byte cell
//     per (slot, lane): u64 ns | u64 (count << 48 | launches).  Lane L only ever touches column L: no races, and a
//     128-bit access per lane is conflict-free per quarter warp.  One 16-byte load + two 64-bit adds + one 16-byte
//     store per record, GEMHOOK_ILP records per lane at a time with same-slot records merged in registers first so the
//     read-modify-write chains are independent (bin_add_group).  The packed half holds < 2^16 records per column, so every FLUSH_EVERY tiles a warp folds its bins
//     into its own u64 accumulators (never in practice below 2^31 records per launch; tested with a small value);
//   * epilogue: up to four slots (a handful of pods per GPU, the common case) are folded with __shfl_down_sync trees;
//     beyond that lane L sums the 32 columns of slots L, L+32 itself with a rotated column index (conflict-free, 64
//     loads instead of 30 shuffles per slot); warps are folded through shared memory, ONE atomicAdd per (slot, field)
//     per block;
//   * the last block to finish (threadfence + ticket) publishes the running totals -- and the mirror of the pod's
//     gpu_mem counter the host passes along -- to the mapped pinned totals page: double-buffered by epoch parity,
//     one system fence, then the epoch store;
//   * gemhook_acct_reduce_small: one warp, no ticket, totals taken from the atomics' return values -- the live
//     hook's regime (a flush carries tens of records) where fixed costs are everything;
//   * gemhook_acct_reduce_staged[_c16] (above 22 client slots): the same bins fed from per-warp shared-memory rings that
//     cp.async.bulk (TMA, mbarrier-tracked) keeps filled, software-pipelined bin update -- see the comment at the kernel.
//
// Build: nvcc -cubin -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 (csrc/Makefile); the cubin
// is embedded in libgemhook.so.1 and loaded with cuModuleLoadData (no cudart dependency).

#include <stdint.h>

#ifndef GEMHOOK_UNROLL
#define GEMHOOK_UNROLL 8               /* 16-byte loads per lane per tile; two tiles are in flight (host: gh_acct.cpp TILE_RECORDS) */
#endif
#define GEMHOOK_MAX_WARPS_PER_BLOCK 8

typedef unsigned long long u64;

extern "C" {

#define GEMHOOK_PAGE_MAX_SLOTS 64
struct gemhook_totals_page {   // mapped pinned page (host reads it without any CUDA call)
  u64 epoch;                   // number of reduce launches published; buf[epoch & 1] holds the current totals
  u64 nslots;
  u64 mem_slot;                // which slot the memory mirror below describes (the publishing process's own pod)
  u64 reserved;
  u64 buf[2][GEMHOOK_PAGE_MAX_SLOTS * 3];  // [slot][3]: elapsed_ns, launches, records
  u64 mem[2][2];               // [epoch parity]: mem_used, mem_limit of mem_slot as the host passed them at launch
};

// what the host passes along with every launch: the pod's gpu_mem counter (authoritative copy: the CAS word in the
// shared-pinned credit pool) to be mirrored into device memory and the totals page
struct gemhook_mem_mirror {
  u64 slot, used, limit;
};

__device__ __forceinline__ uint4 ld_stream_16(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

}  // extern "C"

#define PK_ONE (1ull << 48)
#define PK_MASK (PK_ONE - 1ull)

// bin columns per slot: 32 = every lane owns a column; 16 = lanes L and L+16 share column L and take turns (two phases
// per tile, separated by __syncwarp): half the shared memory per slot, i.e. twice the warps per SM when the slot table
// is large.  The host (gh_acct.cpp) must be compiled with the same value.
#ifndef GEMHOOK_COLS
#define GEMHOOK_COLS 32
#endif
#define COLS ((unsigned)GEMHOOK_COLS)

// 16-byte shared-memory accesses spelled out: left to itself the compiler splits the cell load into two LDS.64, which
// at a 16-byte lane stride is a 2-way bank conflict each; one LDS.128 per cell is conflict-free per quarter warp
__device__ __forceinline__ uint4 lds128(const uint4* p) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}
__device__ __forceinline__ void sts128(uint4* p, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

// one record into the warp's bins: cell (slot, lane) = {ns, count << 48 | launches}.  Branch-free: a slot outside
// [0, nslots) lands in the trash row `nslots`, which is zeroed with the others and never folded.
template <unsigned C = COLS>
__device__ __forceinline__ void bin_add(uint4* cells, unsigned nslots, unsigned col, const uint4& r) {
  uint4* c = cells + min(r.x, nslots) * C + col;
  uint4 v = lds128(c);
  u64 ns = (((u64)v.y << 32) | v.x) + (((u64)r.w << 32) | r.z);
  u64 pk = (((u64)v.w << 32) | v.z) + (PK_ONE | (u64)r.y);
  sts128(c, make_uint4((unsigned)ns, (unsigned)(ns >> 32), (unsigned)pk, (unsigned)(pk >> 32)));
}

// G records of one lane at once.  A read-modify-write of a shared-memory cell is a dependent chain (load -> add ->
// store -> the next load may hit the same cell), and with few warps per SM (bins of 64 slots leave room for six) that
// chain, not HBM, sets the pace: measured 0.74 of the roofline at 64 slots with one record at a time.  Records of the
// group that name the same slot are first merged in registers (the later one is redirected to the trash row), so the G
// cells are distinct and their loads, adds and stores are independent: G chains in flight per lane instead of one.
#ifndef GEMHOOK_ILP
#define GEMHOOK_ILP 2 /* measured (profiles/r02_acct_reduce_variants.jsonl): 1 -> 0.74, 2 -> 0.79, 4 -> 0.79 of the roofline at 64 slots */
#endif
template <int G, unsigned C = COLS>
__device__ __forceinline__ void bin_add_group(uint4* cells, unsigned nslots, unsigned col, const uint4* r) {
  unsigned sl[G];
  u64 ns[G], pk[G];
#pragma unroll
  for (int j = 0; j < G; j++) {
    sl[j] = min(r[j].x, nslots);
    ns[j] = ((u64)r[j].w << 32) | r[j].z;
    pk[j] = PK_ONE | (u64)r[j].y;
  }
#pragma unroll
  for (int j = 1; j < G; j++) {
#pragma unroll
    for (int i = 0; i < j; i++) {
      const bool same = sl[j] == sl[i];
      ns[i] += same ? ns[j] : 0ull;
      pk[i] += same ? pk[j] : 0ull;
      sl[j] = same ? nslots : sl[j];  // (what a redirected record still carries goes to the trash row: harmless)
    }
  }
  uint4 v[G];
#pragma unroll
  for (int j = 0; j < G; j++) v[j] = lds128(cells + sl[j] * C + col);
#pragma unroll
  for (int j = 0; j < G; j++) {
    u64 a = (((u64)v[j].y << 32) | v[j].x) + ns[j];
    u64 b = (((u64)v[j].w << 32) | v[j].z) + pk[j];
    v[j] = make_uint4((unsigned)a, (unsigned)(a >> 32), (unsigned)b, (unsigned)(b >> 32));
  }
  // two redirected records of one group share the trash cell: whichever store lands last wins, nobody reads it
#pragma unroll
  for (int j = 0; j < G; j++) sts128(cells + sl[j] * C + col, v[j]);
}
// U records of one lane, software-pipelined: the cell of record u+1 is loaded BEFORE the cell of record u is stored, so the
// load latency of one record overlaps the adds of the previous one; if the two records name the same cell the loaded value
// is stale and the value just computed is forwarded instead (one compare + four selects per record, no merging pass).
// Measured in the TMA-staged kernel at 64 slots: merged groups of 2 / 4 / 8 records 0.865 / 0.880 / 0.643, one record at a
// time 0.741, this 0.880 of the roofline -- and the best or equal at every other slot count.
template <int U, unsigned C = COLS>
__device__ __forceinline__ void bin_add_tile_fwd(uint4* cells, unsigned nslots, unsigned lane, const uint4* r) {
  unsigned off[U];
#pragma unroll
  for (int u = 0; u < U; u++) off[u] = min(r[u].x, nslots) * C + (lane & (C - 1u));
#pragma unroll
  for (unsigned ph = 0; ph < 32u / C; ph++) {  // lanes sharing a column take turns
    if (C == 32u || (lane / C) == ph) {
      uint4 cur = lds128(cells + off[0]);
#pragma unroll
      for (int u = 0; u < U; u++) {
        uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
        if (u + 1 < U) nxt = lds128(cells + off[u + 1]);
        const u64 a = (((u64)cur.y << 32) | cur.x) + (((u64)r[u].w << 32) | r[u].z);
        const u64 b = (((u64)cur.w << 32) | cur.z) + (PK_ONE | (u64)r[u].y);
        const uint4 nv = make_uint4((unsigned)a, (unsigned)(a >> 32), (unsigned)b, (unsigned)(b >> 32));
        sts128(cells + off[u], nv);
        if (u + 1 < U) {
          const bool same = off[u + 1] == off[u];
          cur.x = same ? nv.x : nxt.x;
          cur.y = same ? nv.y : nxt.y;
          cur.z = same ? nv.z : nxt.z;
          cur.w = same ? nv.w : nxt.w;
        }
      }
    }
    if (C != 32u) __syncwarp();
  }
}

template <int U, int ILP = GEMHOOK_ILP, unsigned C = COLS>
__device__ __forceinline__ void bin_add_tile(uint4* cells, unsigned nslots, unsigned lane, const uint4* r) {
  constexpr int G = (ILP <= U && U % ILP == 0) ? ILP : 1;
  const unsigned col = lane & (C - 1u);
#pragma unroll
  for (unsigned ph = 0; ph < 32u / C; ph++) {  // lanes sharing a column take turns
    if (C == 32u || (lane / C) == ph) {
#pragma unroll
      for (int u = 0; u < U; u += G) {
        if (G == 1) bin_add<C>(cells, nslots, col, r[u]);
        else bin_add_group<G, C>(cells, nslots, col, r + u);
      }
    }
    if (C != 32u) __syncwarp();
  }
}

// fold the warp's bins: lane L owns slots L, L+32, ...; column index rotated by the lane -> conflict-free LDS.128.
// acc[slot][3] (warp private, u64) += column sums; optionally the bins are zeroed for the next round.
// Few clients (the common case: a handful of pods per GPU): the classic warp tree.  Every lane contributes the cell of
// its own column, five __shfl_down_sync steps per field leave the slot's sums in lane 0.  For many slots the
// transposed fold below is cheaper (2 x 32 loads per lane instead of 30 shuffles per slot).
#ifndef GEMHOOK_SHFL_SLOTS
#define GEMHOOK_SHFL_SLOTS 4
#endif
__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  return v;
}
template <unsigned C = COLS>
__device__ __forceinline__ void warp_tree_slot(const uint4* cells, unsigned s, unsigned lane, u64& ns, u64& la, u64& rc) {
  const bool own = lane < C;
  uint4 v = own ? cells[s * C + lane] : make_uint4(0u, 0u, 0u, 0u);
  u64 pk = ((u64)v.w << 32) | v.z;
  ns = warp_sum_u64(((u64)v.y << 32) | v.x);
  la = warp_sum_u64(pk & PK_MASK);
  rc = warp_sum_u64(pk >> 48);
}

template <unsigned C = COLS>
__device__ __forceinline__ void zero_bins(uint4* cells, unsigned nslots, unsigned lane) {
  for (unsigned t = lane; t < (nslots + 1u) * C; t += 32u) cells[t] = make_uint4(0u, 0u, 0u, 0u);
}
template <unsigned C = COLS>
__device__ __forceinline__ void fold_bins(uint4* cells, u64* acc, unsigned nslots, unsigned lane, bool rezero) {
  __syncwarp();
  if (nslots <= GEMHOOK_SHFL_SLOTS) {
    for (unsigned s = 0; s < nslots; s++) {
      u64 ns, la, rc;
      warp_tree_slot<C>(cells, s, lane, ns, la, rc);
      if (lane == 0) {
        acc[s * 3u + 0u] += ns;
        acc[s * 3u + 1u] += la;
        acc[s * 3u + 2u] += rc;
      }
    }
  } else
  for (unsigned s = lane; s < nslots; s += 32u) {
    u64 ns = 0ull, la = 0ull, rc = 0ull;
#pragma unroll 8
    for (unsigned c = 0; c < C; c++) {
      uint4 v = cells[s * C + ((c + lane) & (C - 1u))];
      u64 pk = ((u64)v.w << 32) | v.z;
      ns += ((u64)v.y << 32) | v.x;
      la += pk & PK_MASK;
      rc += pk >> 48;
    }
    acc[s * 3u + 0u] += ns;
    acc[s * 3u + 1u] += la;
    acc[s * 3u + 2u] += rc;
  }
  __syncwarp();
  if (rezero) {
    zero_bins<C>(cells, nslots, lane);
    __syncwarp();
  }
}

// totals -> buf[(e+1) & 1] of the mapped pinned page, ONE system-scope fence, then the 8-byte epoch store that flips
// the reader over (reader rule: gh_acct.cpp read_page).  The publish counter lives in device memory
// (dev_totals[nslots*3]) so nothing is ever READ over PCIe here.  Called by all threads of ONE block.
__device__ __forceinline__ void publish_page(unsigned nslots, u64* __restrict__ dev_totals, gemhook_totals_page* __restrict__ page,
                                             const gemhook_mem_mirror& mm, u64* __restrict__ dev_mem) {
  __shared__ u64 e_sh;
  if (threadIdx.x == 0) e_sh = *reinterpret_cast<volatile u64*>(dev_totals + nslots * 3u) + 1ull;
  __syncthreads();
  const u64 e = e_sh;
  u64* dst = page->buf[e & 1ull];
  for (unsigned t = threadIdx.x; t < nslots * 3u; t += blockDim.x) {
    // read through L2 (the atomics were resolved there); volatile avoids a stale L1 line
    dst[t] = *reinterpret_cast<volatile u64*>(dev_totals + t);
  }
  if (threadIdx.x == 0) {
    page->mem[e & 1ull][0] = mm.used;
    page->mem[e & 1ull][1] = mm.limit;
    if (dev_mem) {  // device-resident mirror of the pod's gpu_mem counter
      dev_mem[0] = mm.used;
      dev_mem[1] = mm.limit;
      dev_mem[2] = mm.slot;
      dev_mem[3] = e;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    dev_totals[nslots * 3u] = e;
    page->nslots = nslots;
    page->mem_slot = mm.slot;
    *reinterpret_cast<volatile u64*>(&page->epoch) = e;
  }
}

// After every warp folded its bins into its accumulators: fold the warps (thread t handles (slot, field) t; ONE atomic per
// (slot, field) per block), then the last block of the launch (threadfence + ticket) publishes.  Called by the whole block.
template <unsigned C = COLS>
__device__ __forceinline__ void block_epilogue(unsigned char* smem, unsigned nwarps, unsigned nslots, u64* __restrict__ dev_totals,
                                               unsigned* __restrict__ ticket, gemhook_totals_page* __restrict__ page,
                                               const gemhook_mem_mirror& mm, u64* __restrict__ dev_mem) {
  __syncthreads();
  const u64* acc0 = reinterpret_cast<const u64*>(smem + (size_t)nwarps * (nslots + 1u) * C * 16u);
  for (unsigned t = threadIdx.x; t < nslots * 3u; t += blockDim.x) {
    u64 v = 0ull;
    for (unsigned w = 0; w < nwarps; w++) v += acc0[(size_t)w * nslots * 3u + t];
    if (v) atomicAdd(dev_totals + t, v);
  }
  __shared__ unsigned is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned prev = atomicAdd(ticket, 1u);
    is_last = (prev == gridDim.x - 1u) ? 1u : 0u;
    if (is_last) *ticket = 0u;  // self-reset for the next launch (stream-ordered)
  }
  __syncthreads();
  if (is_last && page) {
    __threadfence();
    publish_page(nslots, dev_totals, page, mm, dev_mem);
  }
}

extern "C" {

// dev_totals: [nslots][3] u64 running totals + 1 u64 publish counter (device memory, persistent)
// ticket:     u32 zero-initialised, self-resetting
// dynamic shared memory: warps * ((nslots + 1) * COLS * 16 + nslots * 24) bytes  (bins incl. the trash row, then the warps'
// u64 accumulators)
__global__ void __launch_bounds__(GEMHOOK_MAX_WARPS_PER_BLOCK * 32, 2)
gemhook_acct_reduce(const uint4* __restrict__ rec, u64 n, unsigned nslots, u64* __restrict__ dev_totals,
                    unsigned* __restrict__ ticket, gemhook_totals_page* __restrict__ page, gemhook_mem_mirror mm,
                    u64* __restrict__ dev_mem, unsigned flush_every) {
  extern __shared__ __align__(16) unsigned char smem[];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  const unsigned nwarps = blockDim.x >> 5;
  uint4* cells = reinterpret_cast<uint4*>(smem) + (size_t)warp * (nslots + 1u) * COLS;
  u64* acc = reinterpret_cast<u64*>(smem + (size_t)nwarps * (nslots + 1u) * COLS * 16u) + (size_t)warp * nslots * 3u;

  zero_bins(cells, nslots, lane);
  for (unsigned t = lane; t < nslots * 3u; t += 32u) acc[t] = 0ull;
  __syncwarp();

  const u64 warps_total = (u64)gridDim.x * nwarps;
  const u64 tile = 32ull * GEMHOOK_UNROLL;  // records per warp-iteration
  u64 base = ((u64)blockIdx.x * nwarps + warp) * tile;
  const u64 stride = warps_total * tile;
  unsigned since_flush = 0;

  uint4 a[GEMHOOK_UNROLL], b[GEMHOOK_UNROLL];
  bool va = base + tile <= n;
  if (va) {
#pragma unroll
    for (int u = 0; u < GEMHOOK_UNROLL; u++) a[u] = ld_stream_16(rec + base + (unsigned)u * 32u + lane);
  }
  while (va) {
    // tile A is in registers (or on its way): put tile B in flight, then accumulate A
    u64 nb = base + stride;
    const bool vb = nb + tile <= n;
    if (vb) {
#pragma unroll
      for (int u = 0; u < GEMHOOK_UNROLL; u++) b[u] = ld_stream_16(rec + nb + (unsigned)u * 32u + lane);
    }
    bin_add_tile<GEMHOOK_UNROLL>(cells, nslots, lane, a);
    base = nb;
    if (++since_flush >= flush_every) {
      fold_bins(cells, acc, nslots, lane, true);
      since_flush = 0;
    }
    if (!vb) break;
    u64 na = base + stride;
    va = na + tile <= n;
    if (va) {
#pragma unroll
      for (int u = 0; u < GEMHOOK_UNROLL; u++) a[u] = ld_stream_16(rec + na + (unsigned)u * 32u + lane);
    }
    bin_add_tile<GEMHOOK_UNROLL>(cells, nslots, lane, b);
    base = na;
    if (++since_flush >= flush_every) {
      fold_bins(cells, acc, nslots, lane, true);
      since_flush = 0;
    }
  }
  if (base < n) {  // ragged tail of this warp's last tile
#pragma unroll 1
    for (int u = 0; u < GEMHOOK_UNROLL; u++) {
      u64 i = base + (unsigned)u * 32u + lane;
      const uint4 r = i < n ? ld_stream_16(rec + i) : make_uint4(0xffffffffu, 0u, 0u, 0u);
      bin_add_tile<1>(cells, nslots, lane, &r);
    }
  }
  fold_bins(cells, acc, nslots, lane, false);
  block_epilogue(smem, nwarps, nslots, dev_totals, ticket, page, mm, dev_mem);
}

// ---- many client slots: the same accumulation fed by TMA bulk copies -------------------------------------------------
// Bins of 64 slots take 34.8 KB per warp, so only a handful of warps fit into an SM, and with register-staged loads
// (UNROLL x 16 B per lane, two tiles) a handful of warps cannot keep the ~36 KB per SM in flight that HBM needs: measured
// 0.55 / 0.65 / 0.74 of the roofline with 4 / 5 / 6 warps -- proportional to the warp count, i.e. bound by bytes in flight,
// not by the bin updates.  Here every warp owns a ring of `stages` 4 KB buffers in shared memory that lane 0 keeps filled
// with cp.async.bulk (SASS UBLKCP, completion counted on an mbarrier per buffer): the bytes in flight are set by the ring,
// not by the register file, and four warps are enough for the arithmetic (one 32-record row per ~90 cycles per warp).
// dynamic shared memory: bins + accumulators of the warps as above, then (16-byte aligned) warps x stages x 4096 bytes of
// staging and warps x stages mbarriers.
// bytes per ring buffer: R rows of 32 records (R = GEMHOOK_UNROLL = 8 -> 4 KB; 4 -> 2 KB where shared memory is tightest)
#define STG_BYTES(R) (32u * (unsigned)(R) * 16u)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void stage_fill(unsigned dst, const uint4* src, unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void stage_wait(unsigned bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}

}  // extern "C"

template <unsigned C, int R>
__device__ __forceinline__ void reduce_staged_body(const uint4* __restrict__ rec, u64 n, unsigned nslots, u64* __restrict__ dev_totals,
                                                   unsigned* __restrict__ ticket, gemhook_totals_page* __restrict__ page,
                                                   const gemhook_mem_mirror& mm, u64* __restrict__ dev_mem, unsigned flush_every,
                                                   unsigned stages) {
  extern __shared__ __align__(16) unsigned char smem_st[];  // (own name: C++ linkage here, C linkage in the kernels above)
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  const unsigned nwarps = blockDim.x >> 5;
  uint4* cells = reinterpret_cast<uint4*>(smem_st) + (size_t)warp * (nslots + 1u) * C;
  u64* acc = reinterpret_cast<u64*>(smem_st + (size_t)nwarps * (nslots + 1u) * C * 16u) + (size_t)warp * nslots * 3u;
  const unsigned bins_bytes = nwarps * ((nslots + 1u) * C * 16u + nslots * 24u);
  unsigned char* stg_all = smem_st + ((bins_bytes + 15u) & ~15u);  // (cp.async.bulk: 16-byte aligned destination)
  const unsigned stg = smem_u32(stg_all) + warp * stages * STG_BYTES(R);
  const unsigned bars = smem_u32(stg_all) + nwarps * stages * STG_BYTES(R) + warp * stages * 8u;

  zero_bins<C>(cells, nslots, lane);
  for (unsigned t = lane; t < nslots * 3u; t += 32u) acc[t] = 0ull;
  if (lane == 0) {
    for (unsigned s = 0; s < stages; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + s * 8u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const u64 tile = 32ull * R;  // records per buffer
  const u64 full = n / tile;                // whole tiles: bulk copies; the ragged rest goes through plain loads below
  const u64 gw = (u64)blockIdx.x * nwarps + warp, GW = (u64)gridDim.x * nwarps;
  // this warp's k-th tile is tile gw + k * GW; its buffer is k % stages, used for the (k / stages)-th time
  if (lane == 0) {
    for (unsigned k = 0; k < stages; k++) {
      const u64 t = gw + (u64)k * GW;
      if (t < full) stage_fill(stg + k * STG_BYTES(R), rec + t * tile, bars + k * 8u, STG_BYTES(R));
    }
  }
  unsigned since_flush = 0, s = 0, parity = 0;
  for (u64 t = gw; t < full; t += GW) {
    stage_wait(bars + s * 8u, parity);
    uint4 r[R];
    const uint4* buf = reinterpret_cast<const uint4*>(stg_all + (size_t)(warp * stages + s) * STG_BYTES(R));
#pragma unroll
    for (int u = 0; u < R; u++) r[u] = lds128(buf + (unsigned)u * 32u + lane);
    bin_add_tile_fwd<R, C>(cells, nslots, lane, r);
    // every lane has consumed its rows (the bin updates depend on them): the buffer may be overwritten
    __syncwarp();
    const u64 nt = t + (u64)stages * GW;
    if (lane == 0 && nt < full) stage_fill(stg + s * STG_BYTES(R), rec + nt * tile, bars + s * 8u, STG_BYTES(R));
    if (++s == stages) {
      s = 0;
      parity ^= 1u;
    }
    if (++since_flush >= flush_every) {
      fold_bins<C>(cells, acc, nslots, lane, true);
      since_flush = 0;
    }
  }
  if (gw == full % GW) {  // ragged tail of the ring (< one tile): the warp whose turn it would be
#pragma unroll 1
    for (int u = 0; u < R; u++) {
      const u64 i = full * tile + (unsigned)u * 32u + lane;
      const uint4 r = i < n ? ld_stream_16(rec + i) : make_uint4(0xffffffffu, 0u, 0u, 0u);
      bin_add_tile<1, 1, C>(cells, nslots, lane, &r);
    }
  }
  fold_bins<C>(cells, acc, nslots, lane, false);
  block_epilogue<C>(smem_st, nwarps, nslots, dev_totals, ticket, page, mm, dev_mem);
}

extern "C" {

// Two instances: 32 columns (every lane its own) while eight warps with two buffers each still fit, 16 columns (lanes L and
// L+16 share a column and take turns) beyond that -- half the bins, twice the warps.  With one warp per scheduler (four warps
// per SM at 64 slots and 32 columns) instruction latency is exposed: 0.88 of the roofline; with 16 columns and eight warps
// 0.95 (profiles/r02_acct_staged_variants.jsonl).  The host picks (gh_acct.cpp).
#define STAGED_KERNEL(NAME, C, R, MAXW)                                                                                     \
  __global__ void __launch_bounds__((MAXW) * 32, 1)                                                                         \
  NAME(const uint4* __restrict__ rec, u64 n, unsigned nslots, u64* __restrict__ dev_totals, unsigned* __restrict__ ticket, \
       gemhook_totals_page* __restrict__ page, gemhook_mem_mirror mm, u64* __restrict__ dev_mem, unsigned flush_every,     \
       unsigned stages) {                                                                                                  \
    reduce_staged_body<C, R>(rec, n, nslots, dev_totals, ticket, page, mm, dev_mem, flush_every, stages);                  \
  }
STAGED_KERNEL(gemhook_acct_reduce_staged, 32u, GEMHOOK_UNROLL, GEMHOOK_MAX_WARPS_PER_BLOCK)
STAGED_KERNEL(gemhook_acct_reduce_staged_c16, 16u, GEMHOOK_UNROLL, GEMHOOK_MAX_WARPS_PER_BLOCK)
// (2 KB buffers -- R = 4 -- with nine or ten warps per SM were measured too: 0.80-0.82 of the roofline at 48-64 slots against
//  0.95-1.00 for 4 KB buffers and eight warps; the per-buffer costs -- barrier wait, copy issue -- double.)

// The live hook's regime: a flush carries a handful to a few thousand records.  ONE warp: no bin zeroing for eight
// warps, no shuffle trees, no ticket; the running totals come back from the atomics themselves, so nothing is re-read.
// dynamic shared memory: (nslots + 1) * COLS * 16 bytes.
// (Tried: pre-loading the running totals through L2 at kernel start and updating them with plain stores instead of atomics
//  with return -- 6.7-7.7 us under ncu against 5.8-6.4 us for this version on a box whose noop took 9 % longer: not kept.)
__global__ void __launch_bounds__(32, 1)
gemhook_acct_reduce_small(const uint4* __restrict__ rec, unsigned n, unsigned nslots, u64* __restrict__ dev_totals,
                          gemhook_totals_page* __restrict__ page, gemhook_mem_mirror mm, u64* __restrict__ dev_mem) {
  extern __shared__ __align__(16) unsigned char smem[];
  const unsigned lane = threadIdx.x;
  uint4* cells = reinterpret_cast<uint4*>(smem);
  zero_bins(cells, nslots, lane);
  __syncwarp();
  // n <= 512 (host, gh_acct.cpp SMALL_N): a handful of records per lane, the packed count cannot overflow
  for (unsigned base = 0; base < n; base += 32u * 8u) {
    uint4 r[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      unsigned i = base + (unsigned)u * 32u + lane;
      r[u] = i < n ? ld_stream_16(rec + i) : make_uint4(0xffffffffu, 0u, 0u, 0u);
    }
    bin_add_tile<8>(cells, nslots, lane, r);
  }
  __syncwarp();
  const u64 e = *reinterpret_cast<volatile u64*>(dev_totals + nslots * 3u) + 1ull;
  u64* dst = page ? page->buf[e & 1ull] : nullptr;
  const bool tree = nslots <= GEMHOOK_SHFL_SLOTS;  // few clients: warp tree, lane 0 owns every slot
  for (unsigned s = tree ? 0u : lane; s < nslots; s += tree ? 1u : 32u) {
    u64 ns = 0ull, la = 0ull, rc = 0ull;
    if (tree) {
      warp_tree_slot(cells, s, lane, ns, la, rc);
      if (lane != 0) continue;
    } else {
#pragma unroll 8
      for (unsigned c = 0; c < COLS; c++) {
        uint4 v = cells[s * COLS + ((c + lane) & (COLS - 1u))];
        u64 pk = ((u64)v.w << 32) | v.z;
        ns += ((u64)v.y << 32) | v.x;
        la += pk & PK_MASK;
        rc += pk >> 48;
      }
    }
    // the kernel is alone on its stream and owns dev_totals: the values the atomics return ARE the old totals
    u64 t0 = rc ? atomicAdd(dev_totals + s * 3u + 0u, ns) + ns : *reinterpret_cast<volatile u64*>(dev_totals + s * 3u + 0u);
    u64 t1 = rc ? atomicAdd(dev_totals + s * 3u + 1u, la) + la : *reinterpret_cast<volatile u64*>(dev_totals + s * 3u + 1u);
    u64 t2 = rc ? atomicAdd(dev_totals + s * 3u + 2u, rc) + rc : *reinterpret_cast<volatile u64*>(dev_totals + s * 3u + 2u);
    if (dst) {
      dst[s * 3u + 0u] = t0;
      dst[s * 3u + 1u] = t1;
      dst[s * 3u + 2u] = t2;
    }
  }
  if (lane == 0) {
    if (page) {
      page->mem[e & 1ull][0] = mm.used;
      page->mem[e & 1ull][1] = mm.limit;
    }
    if (dev_mem) {
      dev_mem[0] = mm.used;
      dev_mem[1] = mm.limit;
      dev_mem[2] = mm.slot;
      dev_mem[3] = e;
    }
  }
  __threadfence_system();
  __syncwarp();
  if (lane == 0) {
    dev_totals[nslots * 3u] = e;
    if (page) {
      page->nslots = nslots;
      page->mem_slot = mm.slot;
      *reinterpret_cast<volatile u64*>(&page->epoch) = e;
    }
  }
}

// Proof of "shared-pinned": read the pod's gpu_mem counter words straight out of the credit pool (host memory, page
// locked and device-mapped by cuMemHostRegister) from device code and leave them in device memory.  Not on any hot
// path: launched on demand by gemhook_acct_peek_pool().
__global__ void gemhook_peek_pool(const volatile u64* __restrict__ pool_words, u64* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    out[0] = pool_words[0];  // mem_used   (SlotShared, gh_pool.cpp)
    out[1] = pool_words[1];  // mem_limit mirror
    out[2] = pool_words[2];  // gpu_ns published so far
    out[3] = pool_words[3];  // launches published so far
  }
}

// Zero the running totals (stream-ordered reset).
__global__ void gemhook_acct_clear(u64* __restrict__ dev_totals, unsigned n) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dev_totals[i] = 0ull;
}

}  // extern "C"
