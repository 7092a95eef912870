// acct_kernels.cu -- sm_100a device code of the gemhook accounting path.
//
// The reference hook has no device code at all: its only GPU-time measurement is one event pair per
// token resolved on the host (reference Gemini/src/hook.cpp:456-502).  The B200-native hook stamps
// launch segments with CUDA events (csrc/acct.cpp) and reduces the resulting 16-byte launch records
// on the device, so the per-client running totals live next to the ring in HBM and only one small
// snapshot per launch crosses to the mapped pinned totals page the host gate reads.
//
// Record (16 B, one uint4, see include/gemhook.h gemhook_record):
//     x = slot            client slot in the credit pool; slot >= nslots is ignored
//     y = launches        kernel launches covered by the record
//     z,w = elapsed_ns    u64 little-endian: SM-time of the segment in nanoseconds
// Output: per slot {sum elapsed_ns, sum launches, record count}, all u64 -> integer sums are
// order-independent, so parity with the CPU oracle is bit-exact.
//
// Roofline: pure streaming read, 16 B per record, O(nslots) bytes written per block -> HBM-bound.
// Design (see DESIGN.md "acct_reduce"):
//   * every lane loads whole records with 128-bit ld.global.nc.L1::no_allocate (a warp covers
//     512 contiguous bytes per load, UNROLL independent loads in flight per lane);
//   * privatised accumulation without atomics: each warp owns a column-major bin table in shared
//     memory, bins[slot][lane], so lane L only ever touches column L (conflict-free banks, no races);
//   * block epilogue: columns are folded with __shfl_down_sync, warps are folded through shared
//     memory, and ONE atomicAdd per (slot, field) per block goes to the device-resident totals;
//   * the last block to finish (threadfence + ticket) publishes the running totals to the mapped
//     pinned totals page: double-buffered by epoch parity, one system fence, then the epoch store.
//
// Build: nvcc -cubin -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 (csrc/Makefile); the cubin
// is embedded in libgemhook.so.1 and loaded with cuModuleLoadData (no cudart dependency).

#include <stdint.h>

#define GEMHOOK_MAX_WARPS_PER_BLOCK 8  /* host picks 8 or 4 warps by shared-memory budget */
#ifndef GEMHOOK_UNROLL
#define GEMHOOK_UNROLL 16              /* independent 16-byte loads in flight per lane (host: gh_acct.cpp TILE_RECORDS) */
#endif
#ifndef GEMHOOK_MIN_BLOCKS
#define GEMHOOK_MIN_BLOCKS 1
#endif

extern "C" {

#define GEMHOOK_PAGE_MAX_SLOTS 64
struct gemhook_totals_page {   // mapped pinned page (host reads it without any CUDA call)
  unsigned long long epoch;    // number of reduce launches published; buf[epoch & 1] holds the current totals
  unsigned long long nslots;
  unsigned long long reserved[2];
  unsigned long long buf[2][GEMHOOK_PAGE_MAX_SLOTS * 3];  // [slot][3]: elapsed_ns, launches, records
};

__device__ __forceinline__ uint4 ld_stream_16(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  return v;
}

}  // extern "C" (templates below need C++ linkage)

// Privatised bins, per warp: ns[nslots][COLS] u64 | la[nslots][COLS] u64 | rc[nslots][COLS] u32 (20 B per cell).
// COLS = 32: every lane owns a column -> plain read-modify-write, no races, conflict-free banks.
// COLS = 16: lanes L and L+16 share column L and take turns (two phases per tile separated by __syncwarp):
//            half the shared memory per slot.  The host picks it for nslots > 16: with 32 columns the bins of 20+
//            slots eat so much of the 228 KB L1/shared array that too few loads can be in flight (measured on B200:
//            6.5 TB/s up to 16 slots, 5.4 at 20, 4.5 at 32, 2.5 at 64 with 32 columns).
template <unsigned COLS>
__device__ __forceinline__ void bin_add(unsigned long long* ns, unsigned long long* la, unsigned* rc,
                                        unsigned nslots, unsigned col, const uint4& r) {
  if (r.x < nslots) {
    unsigned idx = r.x * COLS + col;
    ns[idx] += ((unsigned long long)r.w << 32) | r.z;
    la[idx] += r.y;
    rc[idx] += 1u;
  }
}

// Last block of a launch (threadfence + ticket) publishes the running totals to the mapped pinned page.
__device__ __forceinline__ void publish_totals(unsigned nslots, unsigned long long* __restrict__ dev_totals,
                                               unsigned* __restrict__ ticket, gemhook_totals_page* __restrict__ page) {
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
    // Double-buffered publication: the totals go to the buffer the host is NOT reading (epoch parity), ONE
    // system-scope fence orders them before the 8-byte epoch store that flips the reader over.  (A seqlock would
    // need three fences across PCIe; the reader-side rule is in gh_acct.cpp read_page.)  The publish counter lives
    // in device memory (dev_totals[nslots*3]) so nothing is ever READ over PCIe here.
    __shared__ unsigned long long e_sh;
    __threadfence();
    if (threadIdx.x == 0) e_sh = *reinterpret_cast<volatile unsigned long long*>(dev_totals + nslots * 3u) + 1ull;
    __syncthreads();
    const unsigned long long e = e_sh;
    unsigned long long* dst = page->buf[e & 1ull];
    for (unsigned t = threadIdx.x; t < nslots * 3u; t += blockDim.x) {
      // read through L2 (the atomics above were resolved there); volatile avoids a stale L1 line
      dst[t] = *reinterpret_cast<volatile unsigned long long*>(dev_totals + t);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      dev_totals[nslots * 3u] = e;
      page->nslots = nslots;
      *reinterpret_cast<volatile unsigned long long*>(&page->epoch) = e;
    }
  }
}

// dev_totals: [nslots][3] u64 running totals + 1 u64 publish counter (device memory, persistent)
// ticket:     u32 zero-initialised, self-resetting
template <unsigned COLS>
__device__ __forceinline__ void acct_reduce_body(const uint4* __restrict__ rec, unsigned long long n, unsigned nslots,
                                                 unsigned long long* __restrict__ dev_totals,
                                                 unsigned* __restrict__ ticket, gemhook_totals_page* __restrict__ page) {
  extern __shared__ __align__(16) unsigned char smem[];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  const unsigned nwarps = blockDim.x >> 5;
  const unsigned col = lane & (COLS - 1u);
  const unsigned per_warp_bytes = nslots * COLS * 20u;
  unsigned long long* ns = reinterpret_cast<unsigned long long*>(smem + warp * per_warp_bytes);
  unsigned long long* la = ns + nslots * COLS;
  unsigned* rc = reinterpret_cast<unsigned*>(la + nslots * COLS);

  if (lane < COLS) {
    for (unsigned s = 0; s < nslots; s++) {
      ns[s * COLS + lane] = 0ull;
      la[s * COLS + lane] = 0ull;
      rc[s * COLS + lane] = 0u;
    }
  }
  __syncwarp();

  // the per-column record count is u32: a column sees at most 2 n / (32 * warps) records, far below 2^32
  const unsigned long long warps_total = (unsigned long long)gridDim.x * nwarps;
  const unsigned long long gwarp = (unsigned long long)blockIdx.x * nwarps + warp;
  const unsigned long long tile = 32ull * GEMHOOK_UNROLL;  // records per warp-iteration
  unsigned long long base = gwarp * tile;
  const unsigned long long stride = warps_total * tile;

  for (; base + tile <= n; base += stride) {
    uint4 r[GEMHOOK_UNROLL];
#pragma unroll
    for (int u = 0; u < GEMHOOK_UNROLL; u++) r[u] = ld_stream_16(rec + base + (unsigned)u * 32u + lane);
    if (COLS == 32u) {
#pragma unroll
      for (int u = 0; u < GEMHOOK_UNROLL; u++) bin_add<COLS>(ns, la, rc, nslots, col, r[u]);
    } else {
      // lanes sharing a column take turns: phase p = lanes [p*COLS, (p+1)*COLS)
#pragma unroll
      for (unsigned ph = 0; ph < 32u / COLS; ph++) {
        if (lane / COLS == ph) {
#pragma unroll
          for (int u = 0; u < GEMHOOK_UNROLL; u++) bin_add<COLS>(ns, la, rc, nslots, col, r[u]);
        }
        __syncwarp();
      }
    }
  }
  if (base < n) {  // ragged tail of this warp's last tile
#pragma unroll 1
    for (int u = 0; u < GEMHOOK_UNROLL; u++) {
      unsigned long long i = base + (unsigned)u * 32u + lane;
      uint4 r = make_uint4(0xffffffffu, 0u, 0u, 0u);  // out-of-range slot: ignored
      if (i < n) r = ld_stream_16(rec + i);
      if (COLS == 32u) {
        bin_add<COLS>(ns, la, rc, nslots, col, r);
      } else {
        for (unsigned ph = 0; ph < 32u / COLS; ph++) {
          if (lane / COLS == ph) bin_add<COLS>(ns, la, rc, nslots, col, r);
          __syncwarp();
        }
      }
    }
  }
  __syncwarp();

  // fold the columns of every slot with shuffles; lane 0 leaves the warp result in cells 0..2 of the slot's ns row
  for (unsigned s = 0; s < nslots; s++) {
    const bool own = lane < COLS;
    unsigned long long a = warp_sum_u64(own ? ns[s * COLS + lane] : 0ull);
    unsigned long long l = warp_sum_u64(own ? la[s * COLS + lane] : 0ull);
    unsigned long long k = warp_sum_u64(own ? (unsigned long long)rc[s * COLS + lane] : 0ull);
    __syncwarp();
    if (lane == 0) {
      ns[s * COLS] = a;
      ns[s * COLS + 1] = l;
      ns[s * COLS + 2] = k;
    }
  }
  __syncthreads();

  // fold warps: thread t handles (slot, field) = (t / 3, t % 3)
  for (unsigned t = threadIdx.x; t < nslots * 3u; t += blockDim.x) {
    unsigned s = t / 3u, f = t % 3u;
    unsigned long long acc = 0ull;
    for (unsigned w = 0; w < nwarps; w++) {
      const unsigned long long* wns = reinterpret_cast<const unsigned long long*>(smem + w * per_warp_bytes);
      acc += wns[s * COLS + f];
    }
    if (acc) atomicAdd(dev_totals + t, acc);
  }

  publish_totals(nslots, dev_totals, ticket, page);
}


extern "C" {

__global__ void __launch_bounds__(GEMHOOK_MAX_WARPS_PER_BLOCK * 32, GEMHOOK_MIN_BLOCKS)
gemhook_acct_reduce(const uint4* __restrict__ rec, unsigned long long n, unsigned nslots,
                    unsigned long long* __restrict__ dev_totals, unsigned* __restrict__ ticket,
                    gemhook_totals_page* __restrict__ page) {
  acct_reduce_body<32u>(rec, n, nslots, dev_totals, ticket, page);
}

// 16-column variant for nslots > 16 (see bin_add)
__global__ void __launch_bounds__(GEMHOOK_MAX_WARPS_PER_BLOCK * 32, GEMHOOK_MIN_BLOCKS)
gemhook_acct_reduce_c16(const uint4* __restrict__ rec, unsigned long long n, unsigned nslots,
                        unsigned long long* __restrict__ dev_totals, unsigned* __restrict__ ticket,
                        gemhook_totals_page* __restrict__ page) {
  acct_reduce_body<16u>(rec, n, nslots, dev_totals, ticket, page);
}

// Device-side timestamp: slot_ns_signed[slot] += (kind ? +t : -t) with t = %globaltimer, so a begin/end pair adds
// its duration.  NOT used by the hook in round 1 (segments are marked with CUDA events, gh_hook.cpp); kept as the
// building block of the stamp-kernel marking listed in DESIGN.md 7 (a launch costs 2.05 us of host time on the
// box, an event record 2.65 us plus a 2.7 us elapsed query).
__global__ void gemhook_stamp(unsigned long long* __restrict__ slot_ns_signed, unsigned slot, unsigned kind) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  if (kind) atomicAdd(slot_ns_signed + slot, t);
  else atomicAdd(slot_ns_signed + slot, 0ull - t);
}

// Zero the running totals (stream-ordered reset).
__global__ void gemhook_acct_clear(unsigned long long* __restrict__ dev_totals, unsigned n) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dev_totals[i] = 0ull;
}

}  // extern "C"
