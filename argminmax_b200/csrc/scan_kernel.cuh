// scan_kernel.cuh -- the one-launch fused argmin+argmax scan for sm_100a.
//
// B200 sibling of the reference's SIMD backend as a unit:
//   _core_argminmax (hot loop)             src/simd/generic.rs:362-412
//   _overflow_safe_core_argminmax          src/simd/generic.rs:487-543
//   _horiz_min/_horiz_max, min_index_value src/simd/generic.rs:56-87, src/simd/task.rs:265-296
//   argminmax_generic split + merge        src/simd/task.rs:4-55, 149-260
//
// Design (not a translation of the lane algorithm):
//
//  * The array is cut into TILES of blockDim.x * U vectors (VEC = 16 or 32 bytes per
//    vector, streamed with ld.global.nc.L1::no_allocate[.L2::256B]).  A persistent
//    grid (k CTAs per SM) walks the tiles grid-stride; the next tile's loads are
//    issued before the current tile is folded, so every thread keeps 2 x U vectors
//    in flight.
//  * Per element the thread only folds a VALUE-ONLY min and max (policies.cuh).
//    Per VECTOR (16 or 32 bytes) it compares the vector's (min key, max key) with its
//    running pair and remembers WHICH vector won (a 32-bit slot id; strict compare,
//    vectors visited in ascending order => earliest vector per thread).  No index
//    lanes, hence no index overflow and none of the reference's restart loop; the loop
//    is branch-free and its cost does not depend on the data order.
//  * CTA reduce (warp redux, then shared memory) and the cross-CTA step are lexicographic on
//    (key, global vector index): the earliest vector that contains the global extreme wins.
//    32-bit keys: one packed 64-bit word per direction, atomicMin / atomicMax + a ticket
//    (packed_grid_reduce); 64-bit keys: one tagged 64-byte slot per CTA, polled by CTA 0
//    (grid_reduce).
//  * One warp of the finishing CTA finishes exactly: the winning min / max / first-NaN
//    vectors (<= 32 elements each) are re-read with one element per lane and a ballot
//    picks the first match; the unaligned head and the sub-vector tail of the array
//    ride along as per-lane candidates.  That yields first-occurrence indices with the
//    reference's tie-break rule (src/simd/task.rs:273,290) and re-reads < 200 bytes.
//  * One 64-byte record (keys, global indices, first NaN, flags) is written to the
//    workspace's device slot and to a pinned host slot: one launch, no memcpy.
#pragma once
#include <stdint.h>

#include <type_traits>

#include "../../include/argminmax_b200.h"
#include "policies.cuh"
#include "record_slot.h"

namespace amm {

constexpr uint32_t TILE_NONE = 0xFFFFFFFFu;  // "no slot yet" in the per-thread running state
constexpr uint64_t POS_NONE = 0xFFFFFFFFFFFFFFFFull;
constexpr int MAX_THREADS = 512;
constexpr int DIRECT_MAX_THREADS = 512;

struct ScanParams {
  const uint8_t* base;     // caller's device pointer (element 0)
  uint64_t n;              // elements
  uint64_t n_bytes;        // n * sizeof(element)
  uint64_t global_offset;  // added to every reported index (shard offset)
  uint64_t head_elems;     // elements before the first VEC-aligned address (<= n)
  uint64_t n_vec;          // whole vectors in the aligned main region
  uint64_t tail_start;     // first element after the main region
  uint32_t n_full_tiles;   // statically assigned full tiles: n_vec / (blockDim.x * U) - dyn_tiles
  uint32_t rem_vecs;       // n_vec % (blockDim.x * U)
  uint64_t per_vecs;       // contiguous partition: vectors per CTA (a multiple of blockDim.x); 0 = grid-stride tiles
  uint32_t direct_vpt;     // latency kernel: vectors per thread (1..4)
  uint32_t rem_owner;      // grid-stride tiles: the CTA that takes the partial tile
  uint32_t dyn_tiles;      // full tiles behind the static ones that CTAs claim at run time (0 = none)
  uint32_t dyn_koff;       // slot tile ids >= dyn_koff name claimed tiles (UINT32_MAX = none)
  void* partials;          // gridDim.x x 64-byte tagged slots (grid_reduce, 64-bit keys)
  uint32_t* ticket;        // scratch block, layout WS_* below
  amm_record* rec_dev;     // device record slot
  amm_record* rec_host;    // pinned + mapped host record slot (device address) or nullptr
  uint64_t seq;
  int mode;
  int publish_fence;  // debug: also issue a system-scope fence after publishing
  int pipelined;      // 1: the streaming phase may overlap the previous grid's tail (PDL)
  // ---- fused cross-GPU exchange (world > 1): see finalize()
  int world;                   // ranks taking part in this logical call (1 = none)
  int rank;                    // this GPU's position (ascending shard order)
  uint64_t* const* peer_slots;  // device array [world]: every rank's exchange buffer, peer-mapped
  uint64_t xseq;               // exchange sequence number, identical on all ranks
  uint64_t xtimeout_ns;        // give up waiting for a peer after this long (0 = wait for ever)
  uint64_t* stamps;            // AMM_PHASE_STAMPS builds: gridDim.x x 8 %globaltimer stamps (else unused)
};

// 32-bit words of the workspace's scratch block (`ticket`, WS_SCRATCH_BYTES, zero-initialised except
// the packed words): the ticket, the packed candidate words and the claim counters are hit by up
// to 592 CTAs within a microsecond; the ticket and each packed word sit on their own 128-byte
// line (measured: no faster than sharing the ticket's line, profiles/r02_fixed_cost.md -- kept
// because it cannot hurt).
constexpr uint32_t WS_TICKET = 0;        // last-block-done counter (self-resetting)
constexpr uint32_t WS_BOUNDS_ERR = 8;    // AMM_BOUNDS_CHECK builds: out-of-range loads seen
constexpr uint32_t WS_CLAIM_RING = 64;   // 64 claim counters of the dynamic tile schedule, indexed by seq & 63
constexpr uint32_t WS_PACK_MIN = 128;    // packed candidate words (packed_grid_reduce), 64 bits each
constexpr uint32_t WS_PACK_MAX = 160;
constexpr uint32_t WS_PACK_NAN = 192;
constexpr size_t WS_SCRATCH_BYTES = 1024;

// Debug builds (-DAMM_BOUNDS_CHECK, argminmax_b200/_build.py:build_boundscheck) verify that every global load of input
// data lies inside [base, base + n_bytes); violations are counted in a scratch word and
// surface as AMM_REC_BOUNDS in the record (compute-sanitizer is not available on the pool).
struct Bounds {
#ifdef AMM_BOUNDS_CHECK
  const uint8_t* lo;
  const uint8_t* hi;
  uint32_t* err;
  __device__ __forceinline__ explicit Bounds(const ScanParams& p)
      : lo(p.base), hi(p.base + p.n_bytes), err(p.ticket + WS_BOUNDS_ERR) {}
  __device__ __forceinline__ void check(const void* ptr, uint32_t bytes) const {
    const uint8_t* a = static_cast<const uint8_t*>(ptr);
    if (a < lo || a + bytes > hi) atomicAdd(err, 1u);
  }
#else
  __device__ __forceinline__ explicit Bounds(const ScanParams&) {}
  __device__ __forceinline__ void check(const void*, uint32_t) const {}
#endif
};
constexpr uint64_t AMM_REC_BOUNDS = 8ull;  // record.flags bit (debug builds only)

// Phase stamps (-DAMM_PHASE_STAMPS, argminmax_b200/_build.py:build_stamps; tools/phase_stamps.py):
// thread 0 of every CTA records %globaltimer at the phase boundaries of the scan, so that the
// fixed cost of a call can be attributed (profiles/r02_fixed_cost.md).  Compiled out otherwise.
enum : int { ST_ENTRY = 0, ST_FIRST_FOLD = 1, ST_STREAM_DONE = 2, ST_CTA_REDUCED = 3, ST_TICKET = 4,
             ST_GRID_FOLDED = 5, ST_EXACT_DONE = 6, ST_PUBLISHED = 7 };
__device__ __forceinline__ void stamp(const ScanParams& p, int which) {
#ifdef AMM_PHASE_STAMPS
  if (p.stamps != nullptr && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.stamps[(uint64_t)blockIdx.x * 8 + which] = t;
  }
#else
  (void)p;
  (void)which;
#endif
}


// (key, position) candidates for min, max and first NaN.  Positions are global vector
// indices in the streaming phase and element indices in the exact phase.
template <typename Key>
struct Best {
  Key kmin;
  Key kmax;
  uint64_t pmin;
  uint64_t pmax;
  uint64_t pnan;
  __device__ __forceinline__ void clear() {
    kmin = KeyLim<Key>::MAXV;
    kmax = KeyLim<Key>::MINV;
    pmin = POS_NONE;
    pmax = POS_NONE;
    pnan = POS_NONE;
  }
  // lexicographic merge = the reference's min_index_value / max_index_value rule
  // `v < best || (v == best && idx < best_idx)` (src/simd/task.rs:273, 290)
  __device__ __forceinline__ void merge_min(Key k, uint64_t p) {
    if (k < kmin || (k == kmin && p < pmin)) {
      kmin = k;
      pmin = p;
    }
  }
  __device__ __forceinline__ void merge_max(Key k, uint64_t p) {
    if (k > kmax || (k == kmax && p < pmax)) {
      kmax = k;
      pmax = p;
    }
  }
  __device__ __forceinline__ void merge(const Best& o) {
    merge_min(o.kmin, o.pmin);
    merge_max(o.kmax, o.pmax);
    pnan = o.pnan < pnan ? o.pnan : pnan;
  }
};

template <typename Key>
__device__ __forceinline__ Best<Key> shfl_xor(const Best<Key>& b, int lane_mask) {
  Best<Key> o;
  o.kmin = __shfl_xor_sync(0xFFFFFFFFu, b.kmin, lane_mask);
  o.kmax = __shfl_xor_sync(0xFFFFFFFFu, b.kmax, lane_mask);
  o.pmin = __shfl_xor_sync(0xFFFFFFFFu, b.pmin, lane_mask);
  o.pmax = __shfl_xor_sync(0xFFFFFFFFu, b.pmax, lane_mask);
  o.pnan = __shfl_xor_sync(0xFFFFFFFFu, b.pnan, lane_mask);
  return o;
}

// The warp's smallest / largest key among the lanes that have one (one redux for 32-bit keys, two
// for 64-bit keys: high words, then low words among the lanes that match).  Lanes without a key
// must still call; without any key in the warp the result is the identity.
template <typename Key>
__device__ __forceinline__ Key warp_min_key(bool has, Key k) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  if constexpr (sizeof(Key) == 4) {
    return (Key)(__reduce_min_sync(FULL, has ? (uint32_t)k ^ 0x80000000u : 0xFFFFFFFFu) ^ 0x80000000u);
  } else {
    const uint64_t b = has ? (uint64_t)k ^ 0x8000000000000000ull : ~0ull;
    const uint32_t hi = (uint32_t)(b >> 32), lo = (uint32_t)b;
    const uint32_t mh = __reduce_min_sync(FULL, hi);
    const uint32_t ml = __reduce_min_sync(FULL, hi == mh ? lo : 0xFFFFFFFFu);
    return (Key)((((uint64_t)mh << 32) | ml) ^ 0x8000000000000000ull);
  }
}
template <typename Key>
__device__ __forceinline__ Key warp_max_key(bool has, Key k) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  if constexpr (sizeof(Key) == 4) {
    return (Key)(__reduce_max_sync(FULL, has ? (uint32_t)k ^ 0x80000000u : 0u) ^ 0x80000000u);
  } else {
    const uint64_t b = has ? (uint64_t)k ^ 0x8000000000000000ull : 0ull;
    const uint32_t hi = (uint32_t)(b >> 32), lo = (uint32_t)b;
    const uint32_t mh = __reduce_max_sync(FULL, hi);
    const uint32_t ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
    return (Key)((((uint64_t)mh << 32) | ml) ^ 0x8000000000000000ull);
  }
}

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
  const uint32_t mh = __reduce_min_sync(FULL, hi);
  const uint32_t ml = __reduce_min_sync(FULL, hi == mh ? lo : 0xFFFFFFFFu);
  return ((uint64_t)mh << 32) | ml;
}

// Warp-wide lexicographic reduce of (key, position) candidates; every lane gets the result.
// Integer keys: the warp integer min/max instruction (redux.sync) on 32-bit pieces -- key, then
// position among the lanes that hold that key: 8-10 instructions where five shuffle rounds over
// five 64-bit words were ~250 (the code size of a latency kernel is part of its latency).
// Floating-point keys (window kernels, value-domain compares): shuffles.
template <typename Key>
__device__ __forceinline__ void warp_reduce_best(Best<Key>& b) {
  if constexpr (std::is_integral<Key>::value) {
    const bool hmin = b.pmin != POS_NONE, hmax = b.pmax != POS_NONE;
    const Key km = warp_min_key(hmin, b.kmin), kM = warp_max_key(hmax, b.kmax);
    b.pmin = warp_min_u64((hmin && b.kmin == km) ? b.pmin : POS_NONE);
    b.pmax = warp_min_u64((hmax && b.kmax == kM) ? b.pmax : POS_NONE);
    b.kmin = km;
    b.kmax = kM;
    b.pnan = warp_min_u64(b.pnan);
  } else {
#pragma unroll 1
    for (int m = 16; m >= 1; m >>= 1) b.merge(shfl_xor(b, m));
  }
}

// All threads of the CTA must call; every thread returns the CTA-wide result.
template <typename Key>
__device__ __forceinline__ Best<Key> block_reduce(Best<Key> b, Best<Key>* smem /* [32] */) {
  warp_reduce_best(b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  __syncthreads();  // smem may still be read from a previous call
  if (lane == 0) smem[warp] = b;
  __syncthreads();
  if (warp == 0) {
    Best<Key> t;
    t.clear();
    if (lane < nwarps) t = smem[lane];
    warp_reduce_best(t);
    if (lane == 0) smem[0] = t;
  }
  __syncthreads();
  return smem[0];
}

// ------------------------------------------------------------- streaming loads
// Read-once data: non-coherent path, no L1 allocation.  HINT selects the L2 prefetch-size
// qualifier (0: none, 1: .L2::128B, 2: .L2::256B).
template <int VEC, int HINT = 2>
struct VecLoad;
#define AMM_LD_V4(QUAL)                                                              \
  asm volatile("ld.global.nc.L1::no_allocate" QUAL ".v4.u32 {%0,%1,%2,%3}, [%4];"    \
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])                      \
               : "l"(p))
#define AMM_LD_V8(QUAL)                                                                      \
  asm volatile("ld.global.nc.L1::no_allocate" QUAL ".v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" \
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]),     \
                 "=r"(w[6]), "=r"(w[7])                                                      \
               : "l"(p))
template <int HINT>
struct VecLoad<16, HINT> {
  static constexpr int NW = 4;
  static __device__ __forceinline__ void ld(const uint8_t* p, uint32_t (&w)[4]) {
    if (HINT == 0) AMM_LD_V4("");
    if (HINT == 1) AMM_LD_V4(".L2::128B");
    if (HINT == 2) AMM_LD_V4(".L2::256B");
  }
};
template <int HINT>
struct VecLoad<32, HINT> {
  static constexpr int NW = 8;
  // 256-bit load: new on sm_100 (SASS LDG.E.NA.ENL2.LTC256B.256.CONSTANT)
  static __device__ __forceinline__ void ld(const uint8_t* p, uint32_t (&w)[8]) {
    if (HINT == 0) AMM_LD_V8("");
    if (HINT == 1) AMM_LD_V8(".L2::128B");
    if (HINT == 2) AMM_LD_V8(".L2::256B");
  }
};

// Running per-thread state of the streaming phase.  A SLOT is one vector position of one tile:
// slot = tile * U + j (32-bit); together with the thread id it names one global vector:
//   vector index = tile * (blockDim * U) + j * blockDim + tid,
// and vector index order IS address order, so a lexicographic reduce on (key, vector index)
// finds the vector that holds the first occurrence of the extreme.
template <typename Key>
struct Running {
  Key kmin, kmax;
  uint32_t smin, smax, snan;
};

template <class P, int NW, int U>
__device__ __forceinline__ void fold_tile(const uint32_t (&v)[U][NW], uint32_t tile0,
                                          Running<typename P::Key>& r) {
  using Key = typename P::Key;
#pragma unroll
  for (int j = 0; j < U; ++j) {
    typename P::Acc acc;
    P::init(acc, &v[j][0]);
#pragma unroll
    for (int w = P::WPI; w < NW; w += P::WPI) P::feed(acc, &v[j][w]);
    Key kmin, kmax;
    bool valid, has_nan;
    P::finish(acc, kmin, kmax, valid, has_nan);
    // slots arrive in ascending order per thread, so "strictly better, or nothing yet" IS the
    // lexicographic (key, slot) update: earliest vector per thread, no branches.
    const uint32_t slot = tile0 * U + j;
    const bool first = (r.smin == TILE_NONE);
    const bool um = valid && (first || kmin < r.kmin);
    const bool uM = valid && (first || kmax > r.kmax);
    r.kmin = um ? kmin : r.kmin;
    r.smin = um ? slot : r.smin;
    r.kmax = uM ? kmax : r.kmax;
    r.smax = uM ? slot : r.smax;
    r.snan = (has_nan && r.snan == TILE_NONE) ? slot : r.snan;
  }
}

// Exact (key, index) scan of elements [begin, end) by the whole CTA, 4 independent loads
// in flight per thread.  ASCENDING: the caller guarantees this thread has not yet seen any
// index >= begin, so a strict compare keeps the first occurrence; otherwise ranges may come
// in any order and the full lexicographic rule is applied.
template <class P, bool ASCENDING>
__device__ __forceinline__ void exact_range(const Bounds& bd, const uint8_t* base, uint64_t begin,
                                            uint64_t end, Best<typename P::Key>& b) {
  using Elem = typename P::Elem;
  using Key = typename P::Key;
  const Elem* e = reinterpret_cast<const Elem*>(base);
  constexpr int UN = 4;
  for (uint64_t i0 = begin + threadIdx.x; i0 < end; i0 += (uint64_t)UN * blockDim.x) {
    Elem raw[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const uint64_t i = i0 + (uint64_t)u * blockDim.x;
      if (i < end) bd.check(&e[i], sizeof(Elem));
      raw[u] = (i < end) ? e[i] : Elem(0);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const uint64_t i = i0 + (uint64_t)u * blockDim.x;
      if (i < end) {
        Key k;
        bool valid, is_nan;
        P::classify(raw[u], k, valid, is_nan);
        if (valid) {
          if (ASCENDING) {
            if (k < b.kmin || b.pmin == POS_NONE) {
              b.kmin = k;
              b.pmin = i;
            }
            if (k > b.kmax || b.pmax == POS_NONE) {
              b.kmax = k;
              b.pmax = i;
            }
          } else {
            b.merge_min(k, i);
            b.merge_max(k, i);
          }
        }
        if (P::TRACK_NAN && is_nan && i < b.pnan) b.pnan = i;
      }
    }
  }
}

// Element e of a vector held as 32-bit words (static indices only: stays in registers).
template <class P, int NW>
__device__ __forceinline__ typename P::Elem vec_elem(const uint32_t (&w)[NW], int e) {
  using Elem = typename P::Elem;
  if (sizeof(Elem) == 1) return (Elem)((w[e >> 2] >> (8 * (e & 3))) & 0xFFu);
  if (sizeof(Elem) == 2) return (Elem)((w[e >> 1] >> (16 * (e & 1))) & 0xFFFFu);
  if (sizeof(Elem) == 4) return (Elem)w[e];
  return (Elem)(((uint64_t)w[2 * e + 1] << 32) | (uint64_t)w[2 * e]);
}

// Publish one record to the device slot and (if mapped) the pinned host slot.
// The host polls the pinned slot.  Instead of ordering the eight words with a system-scope
// fence (which costs microseconds on the PCIe path), a ninth word carries a checksum over
// the other eight: the host accepts the slot only when `seq` matches the call AND the
// checksum matches, so a torn / partially arrived record is simply polled again.  The
// checksum is position-sensitive (multiply-rotate per word, seeded): a slot that mixes 16-byte
// pairs of two different records does not pass by accident the way a plain XOR would when the
// stale words happen to cancel (small integer keys and indices do that).
__device__ __forceinline__ void publish_record(const ScanParams& p, const amm_record& r) {
  *p.rec_dev = r;
  if (p.rec_host != nullptr) {
    // 16-byte stores: every store to mapped host memory is its own PCIe write, and the host
    // needs the LAST one to have landed (4 x 16 B + 8 B instead of 9 x 8 B: -0.25 us per call,
    // tools/exp/latency_floor.cu cases G/H)
    uint64_t* h = reinterpret_cast<uint64_t*>(p.rec_host);
    const uint64_t w[8] = {(uint64_t)r.min_key, r.min_idx, (uint64_t)r.max_key, r.max_idx,
                           r.nan_idx,           r.flags,   r.n,                 r.seq};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(h + 2 * i), "l"(w[2 * i]),
                   "l"(w[2 * i + 1])
                   : "memory");
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(h + 8), "l"(mix_checksum(w, HOST_SLOT_SEED))
                 : "memory");
    if (p.publish_fence) __threadfence_system();
  }
}

template <typename Key>
__device__ __forceinline__ amm_record make_record(const ScanParams& p, const Best<Key>& x) {
  amm_record r;
  const bool has_value = x.pmin != POS_NONE;
  const bool has_nan = x.pnan != POS_NONE;
  r.min_key = has_value ? (int64_t)x.kmin : 0;
  r.max_key = has_value ? (int64_t)x.kmax : 0;
  r.min_idx = has_value ? x.pmin + p.global_offset : AMM_NO_INDEX;
  r.max_idx = has_value ? x.pmax + p.global_offset : AMM_NO_INDEX;
  r.nan_idx = has_nan ? x.pnan + p.global_offset : AMM_NO_INDEX;
  r.flags = (has_value ? AMM_REC_HAS_VALUE : 0ull) | (has_nan ? AMM_REC_HAS_NAN : 0ull);
#ifdef AMM_BOUNDS_CHECK
  if (atomicExch(p.ticket + WS_BOUNDS_ERR, 0u) != 0u) r.flags |= AMM_REC_BOUNDS;
#endif
  r.n = p.n;
  r.seq = p.seq;
  return r;
}

// Last step of a call, executed by WARP 0 of the finishing CTA (all 32 lanes must call; `x` is read
// from lane 0).  world == 1: publish this GPU's record.  world > 1: the compute step is FUSED
// with its collective -- instead of returning to the host for an NCCL all-gather, lane t stores
// the 64-byte record straight into peer t's exchange buffer over NVLink (peer-mapped pointers),
// then polls this GPU's own buffer until peer t's record of this call (tag == xseq, checksum
// valid) has landed; lane 0 folds the `world` records in ascending rank order with strict
// compares (= global first occurrence, the reference's merge rule src/simd/generic.rs:513-520)
// and publishes the GLOBAL record.  Two parity banks make back-to-back calls safe: a rank can
// only reach call s+2 after every peer has finished reading call s.
// A peer that does not arrive within xtimeout_ns (0 = wait for ever) makes the failure
// COLLECTIVE: the waiting rank re-pushes its own record with AMM_REC_PEER_TIMEOUT set into every
// peer's buffer, so a late rank that would otherwise have completed normally reads the flag and
// fails too; the host layer then marks the exchange unusable (libamm.cu: amm_workspace_wait).
// XCHG is a compile-time flag of the kernels: single-GPU launches use instantiations that do
// not contain this code at all (its mere presence costs the streaming loop ~3 % -- registers,
// loss of uniform-datapath loop control), multi-GPU launches use the fused instantiations.
__device__ __forceinline__ void push_peer_slot(volatile uint64_t* dst, const uint64_t* src /* [8] */,
                                               uint64_t xseq) {
#pragma unroll
  for (int i = 0; i < 8; ++i) dst[i] = src[i];
  dst[8] = mix_checksum(src, PEER_SLOT_SEED ^ xseq);
  dst[9] = xseq;
}

template <typename Key>
__device__ __forceinline__ void exchange_and_publish_warp(const ScanParams& p, const Best<Key>& x) {
  __shared__ amm_record s_mine;
  __shared__ amm_record s_all[XCHG_MAX_WORLD];
  __shared__ uint32_t s_timeout;
  const int lane = threadIdx.x & 31;
  if (lane == 0) {
    s_mine = make_record(p, x);
    s_timeout = 0;
  }
  __syncwarp();
  const uint64_t bank = (p.xseq & 1ull) * (uint64_t)p.world;
  const bool active = lane < p.world;
  if (active) {
    // push my record into peer `lane`'s buffer, slot [bank + rank]
    volatile uint64_t* dst = p.peer_slots[lane] + (bank + (uint64_t)p.rank) * XCHG_SLOT_WORDS;
    push_peer_slot(dst, reinterpret_cast<const uint64_t*>(&s_mine), p.xseq);
    __threadfence_system();
    // pull peer `lane`'s record out of MY buffer, slot [bank + lane]
    const volatile uint64_t* in = p.peer_slots[p.rank] + (bank + (uint64_t)lane) * XCHG_SLOT_WORDS;
    uint64_t w[10];
    unsigned long long t0 = 0, now = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    bool ok = false;
    uint32_t spins = 0;
    while (!ok) {
      if (in[9] == p.xseq) {
#pragma unroll
        for (int i = 0; i < 10; ++i) w[i] = in[i];
        ok = (mix_checksum(w, PEER_SLOT_SEED ^ p.xseq) == w[8]) && (w[9] == p.xseq);
      }
      if (!ok && p.xtimeout_ns != 0 && (++spins & 0x3Fu) == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > p.xtimeout_ns) {
          atomicExch(&s_timeout, 1u);
          break;
        }
      }
    }
    uint64_t* out = reinterpret_cast<uint64_t*>(&s_all[lane]);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = ok ? w[i] : 0ull;  // not arrived: no value, no NaN
  }
  __syncwarp();
  if (s_timeout != 0) {  // warp-uniform
    // collective failure: poison my slot in every peer's buffer (same call, flag set)
    if (lane == 0) s_mine.flags |= AMM_REC_PEER_TIMEOUT;
    __syncwarp();
    if (active) {
      volatile uint64_t* dst = p.peer_slots[lane] + (bank + (uint64_t)p.rank) * XCHG_SLOT_WORDS;
      push_peer_slot(dst, reinterpret_cast<const uint64_t*>(&s_mine), p.xseq);
      __threadfence_system();
    }
  }
  __syncwarp();
  if (lane == 0) {
    amm_record o;
    bool have = false;
    o.min_key = o.max_key = 0;
    o.min_idx = o.max_idx = o.nan_idx = AMM_NO_INDEX;
    o.n = 0;
    uint64_t bad = s_timeout ? AMM_REC_PEER_TIMEOUT : 0ull;
    for (int g = 0; g < p.world; ++g) {
      const amm_record r = s_all[g];
      o.n += r.n;
      bad |= r.flags & AMM_REC_PEER_TIMEOUT;  // a peer gave up on this call
      if ((r.flags & AMM_REC_HAS_NAN) && r.nan_idx < o.nan_idx) o.nan_idx = r.nan_idx;
      if (!(r.flags & AMM_REC_HAS_VALUE)) continue;
      if (!have || r.min_key < o.min_key || (r.min_key == o.min_key && r.min_idx < o.min_idx)) {
        o.min_key = r.min_key;
        o.min_idx = r.min_idx;
      }
      if (!have || r.max_key > o.max_key || (r.max_key == o.max_key && r.max_idx < o.max_idx)) {
        o.max_key = r.max_key;
        o.max_idx = r.max_idx;
      }
      have = true;
    }
    o.flags = (have ? AMM_REC_HAS_VALUE : 0ull) |
              (o.nan_idx != AMM_NO_INDEX ? AMM_REC_HAS_NAN : 0ull) | bad;
    o.seq = p.seq;
    publish_record(p, o);
  }
}

template <bool XCHG, typename Key>
__device__ __forceinline__ void finalize_warp(const ScanParams& p, const Best<Key>& x) {
  if (XCHG) {
    exchange_and_publish_warp(p, x);
  } else {
    if ((threadIdx.x & 31) == 0) publish_record(p, make_record(p, x));
  }
  stamp(p, ST_PUBLISHED);
}

// Programmatic dependent launch (PDL).  Scans are launched with the programmatic-stream-
// serialization attribute, so the NEXT scan on the stream can be scheduled while this one drains.
// Default (safe) mode: a CTA executes griddepcontrol.wait before its first global load and only
// then signals launch_dependents -- launch latency is hidden, and because the trigger comes after
// the wait a dependent can never run ahead of the kernel that produced THIS scan's input (a
// PDL-aware producer that triggers early is therefore harmless, also for a pipelined scan that
// follows a non-pipelined one, as in windows_huge).
// Pipelined mode (amm_workspace_set_pipelined, caller promises that a scan's input is not
// produced by the kernel right before it on the stream): the trigger is issued on entry and the
// wait moves to the point where the CTA first touches the shared workspace (slots, ticket,
// record slots), so the STREAMING phase of scan k+1 overlaps the fold / locate / publish tail
// of scan k.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait_prior_grid() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pdl_entry(const ScanParams& p) {
  if (p.pipelined) {
    pdl_launch_dependents();
  } else {
    pdl_wait_prior_grid();
    pdl_launch_dependents();
  }
}

// Cross-CTA step for candidates that do not fit one 64-bit word (64-bit keys): TAGGED SLOTS.
// Every CTA but CTA 0 stores its CTA-wide Best -- five 64-bit words plus an integrity word seeded
// with the call's sequence number and the slot index -- into its own 64-byte slot of `partials` and
// exits; CTA 0 polls the slots (every thread one or two at a time, all loads in flight together),
// accepts a slot when its integrity word matches, and folds them.  One one-way trip through L2 plus
// a poll, where "store partial, fence, ticket, last CTA: fence, load all partials" was three
// dependent round trips (tools/exp/latency_floor.cu, 4 MiB skeleton: 9.8 -> 8.5 us per call).  No
// fence and no atomicity assumption beyond aligned 64-bit words: a slot that is torn, or still
// holds an older call's record, fails the integrity check and is polled again.  CTA 0 holds one
// CTA slot of the GPU while it polls and every other CTA runs to completion without waiting, so
// grids larger than the GPU are fine.
// Returns true in warp 0 of CTA 0 only (all 32 lanes), with `b` = the grid-wide result.  `b` must
// be the CTA-wide result in every thread (block_reduce) on entry.
template <typename Key>
__device__ __forceinline__ bool grid_reduce(const ScanParams& p, Best<Key>& b, Best<Key>* smem) {
  pdl_wait_prior_grid();  // from here on the workspace is ours
  if (gridDim.x == 1) return threadIdx.x < 32;
  uint64_t* const slots = reinterpret_cast<uint64_t*>(p.partials);  // gridDim.x x 8 words
  if (blockIdx.x != 0) {
    if (threadIdx.x == 0) {
      const uint64_t w[5] = {(uint64_t)(int64_t)b.kmin, (uint64_t)(int64_t)b.kmax, b.pmin, b.pmax, b.pnan};
      uint64_t* dst = slots + (uint64_t)blockIdx.x * 8;
      asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(w[0]), "l"(w[1]) : "memory");
      asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst + 2), "l"(w[2]), "l"(w[3]) : "memory");
      asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst + 4), "l"(w[4]),
                   "l"(grid_slot_check(w, p.seq, blockIdx.x))
                   : "memory");
    }
    stamp(p, ST_TICKET);
    return false;
  }
  stamp(p, ST_TICKET);
  Best<Key> g = b;
  for (uint32_t s0 = 1 + threadIdx.x; s0 < gridDim.x; s0 += 2 * blockDim.x) {
    const uint32_t s1 = s0 + blockDim.x;
    const bool two = s1 < gridDim.x;
    uint64_t w0[6], w1[6];
    bool ok0 = false, ok1 = !two;
    while (!(ok0 && ok1)) {
      if (!ok0) {
        const uint64_t* src = slots + (uint64_t)s0 * 8;
#pragma unroll
        for (int i = 0; i < 3; ++i)
          asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0[2 * i]), "=l"(w0[2 * i + 1]) : "l"(src + 2 * i) : "memory");
      }
      if (!ok1) {
        const uint64_t* src = slots + (uint64_t)s1 * 8;
#pragma unroll
        for (int i = 0; i < 3; ++i)
          asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w1[2 * i]), "=l"(w1[2 * i + 1]) : "l"(src + 2 * i) : "memory");
      }
      if (!ok0) {
        const uint64_t d[5] = {w0[0], w0[1], w0[2], w0[3], w0[4]};
        ok0 = grid_slot_check(d, p.seq, s0) == w0[5];
      }
      if (!ok1) {
        const uint64_t d[5] = {w1[0], w1[1], w1[2], w1[3], w1[4]};
        ok1 = grid_slot_check(d, p.seq, s1) == w1[5];
      }
    }
    Best<Key> o;
    o.kmin = (Key)(int64_t)w0[0];
    o.kmax = (Key)(int64_t)w0[1];
    o.pmin = w0[2];
    o.pmax = w0[3];
    o.pnan = w0[4];
    g.merge(o);
    if (two) {
      o.kmin = (Key)(int64_t)w1[0];
      o.kmax = (Key)(int64_t)w1[1];
      o.pmin = w1[2];
      o.pmax = w1[3];
      o.pnan = w1[4];
      g.merge(o);
    }
  }
  b = block_reduce(g, smem);
  return threadIdx.x < 32;
}

// ---------------------------------------------------------------------------------------
// Cross-CTA step for 32-bit keys and fewer than 2^32 - 2 positions: a candidate is ONE 64-bit
// word, (biased key << 32) | position for the minimum and (biased key << 32) | ~position for the
// maximum, so "smaller key, then smaller position" is a plain unsigned min and "larger key, then
// smaller position" a plain unsigned max.  Every CTA folds its word into the workspace with one
// atomicMin / atomicMax (+ one for the first NaN position) and takes a ticket; the last CTA swaps
// the words back to their identities and carries on.  Compared with "store partials, ticket, last
// CTA loads and reduces them" this removes one global round trip and one CTA-wide reduce from the
// dependent chain (tools/exp/latency_floor.cu: 9.4 -> 8.1 us for the skeleton of a 4 MiB scan).
// Positions are element indices in the latency kernel and vector indices in the streaming
// kernel.  Words: p.ticket + WS_PACK_MIN / WS_PACK_MAX / WS_PACK_NAN.
// ---------------------------------------------------------------------------------------
constexpr unsigned long long PACK_MIN_NONE = ~0ull;
constexpr unsigned long long PACK_MAX_NONE = 0ull;
constexpr uint32_t PACK_POS_NONE = 0xFFFFFFFFu;

// One thread's / warp's / CTA's candidates in 32-bit pieces: keys BIASED (key ^ 0x80000000, so
// unsigned order is key order), 32-bit positions, PACK_POS_NONE = nothing yet.  Reduced with the
// warp-wide integer min/max instruction (redux.sync, 5 instructions per level instead of 5
// shuffle rounds over three 64-bit words).
struct PackedCand {
  uint32_t kmin, pmin, kmax, pmax, pnan;
  __device__ __forceinline__ void clear() {
    kmin = 0xFFFFFFFFu;
    kmax = 0u;
    pmin = pmax = pnan = PACK_POS_NONE;
  }
  __device__ __forceinline__ void set(int32_t key_min, uint32_t pos_min, int32_t key_max, uint32_t pos_max) {
    kmin = (uint32_t)key_min ^ 0x80000000u;
    pmin = pos_min;
    kmax = (uint32_t)key_max ^ 0x80000000u;
    pmax = pos_max;
  }
  template <bool TRACK_NAN>
  __device__ __forceinline__ void warp_reduce() {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    // a lane without a min (max) candidate holds the identity key and PACK_POS_NONE
    const uint32_t km = __reduce_min_sync(FULL, kmin);
    const uint32_t kM = __reduce_max_sync(FULL, kmax);
    const uint32_t pm = __reduce_min_sync(FULL, kmin == km ? pmin : PACK_POS_NONE);
    const uint32_t pM = __reduce_min_sync(FULL, kmax == kM ? pmax : PACK_POS_NONE);
    kmin = km;
    kmax = kM;
    pmin = pm;
    pmax = pM;
    if (TRACK_NAN) pnan = __reduce_min_sync(FULL, pnan);
  }
  // -> Best with positions widened
  __device__ __forceinline__ Best<int32_t> unpack() const {
    Best<int32_t> g;
    g.clear();
    if (pmin != PACK_POS_NONE) {
      g.kmin = (int32_t)(kmin ^ 0x80000000u);
      g.pmin = pmin;
    }
    if (pmax != PACK_POS_NONE) {
      g.kmax = (int32_t)(kmax ^ 0x80000000u);
      g.pmax = pmax;
    }
    if (pnan != PACK_POS_NONE) g.pnan = pnan;
    return g;
  }
};

// Folds every thread's candidates over the CTA and then over the grid.  Returns true in ALL lanes
// of warp 0 of the last CTA to arrive (or of the only CTA), with the grid-wide candidates in `c`;
// every other thread gets false.  All threads of the CTA must call (one __syncthreads inside).
template <class P>
__device__ __forceinline__ bool packed_grid_reduce(const ScanParams& p, PackedCand& c,
                                                   uint32_t* smem /* [5][32] */) {
  c.warp_reduce<P::TRACK_NAN>();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  if (nwarps > 1) {
    if (lane == 0) {
      smem[warp] = c.kmin;
      smem[32 + warp] = c.pmin;
      smem[64 + warp] = c.kmax;
      smem[96 + warp] = c.pmax;
      if (P::TRACK_NAN) smem[128 + warp] = c.pnan;
    }
    __syncthreads();
    if (warp != 0) return false;
    c.clear();
    if (lane < nwarps) {
      c.kmin = smem[lane];
      c.pmin = smem[32 + lane];
      c.kmax = smem[64 + lane];
      c.pmax = smem[96 + lane];
      if (P::TRACK_NAN) c.pnan = smem[128 + lane];
    }
    c.warp_reduce<P::TRACK_NAN>();
  }
  stamp(p, ST_CTA_REDUCED);
  uint32_t last = 1;
  unsigned long long wmin = PACK_MIN_NONE, wmax = PACK_MAX_NONE, wnan = POS_NONE;
  if (lane == 0) {
    pdl_wait_prior_grid();  // from here on the workspace is ours
    if (gridDim.x > 1) {
      unsigned long long* const word_min = reinterpret_cast<unsigned long long*>(p.ticket + WS_PACK_MIN);
      unsigned long long* const word_max = reinterpret_cast<unsigned long long*>(p.ticket + WS_PACK_MAX);
      unsigned long long* const word_nan = reinterpret_cast<unsigned long long*>(p.ticket + WS_PACK_NAN);
      if (c.pmin != PACK_POS_NONE) atomicMin(word_min, ((unsigned long long)c.kmin << 32) | c.pmin);
      if (c.pmax != PACK_POS_NONE) atomicMax(word_max, ((unsigned long long)c.kmax << 32) | (uint32_t)~c.pmax);
      if (P::TRACK_NAN && c.pnan != PACK_POS_NONE) atomicMin(word_nan, (unsigned long long)c.pnan);
      __threadfence();
      last = atomicInc(p.ticket + WS_TICKET, gridDim.x - 1) == gridDim.x - 1;
      if (last) {
        __threadfence();
        wmin = atomicExch(word_min, PACK_MIN_NONE);  // read the grid-wide result and re-arm
        wmax = atomicExch(word_max, PACK_MAX_NONE);
        if (P::TRACK_NAN) wnan = atomicExch(word_nan, POS_NONE);
      }
    }
  }
  stamp(p, ST_TICKET);
  last = __shfl_sync(0xFFFFFFFFu, last, 0);
  if (!last) return false;
  if (gridDim.x > 1) {
    wmin = __shfl_sync(0xFFFFFFFFu, wmin, 0);
    wmax = __shfl_sync(0xFFFFFFFFu, wmax, 0);
    wnan = __shfl_sync(0xFFFFFFFFu, wnan, 0);
    c.clear();
    if (wmin != PACK_MIN_NONE) {
      c.kmin = (uint32_t)(wmin >> 32);
      c.pmin = (uint32_t)wmin;
    }
    if (wmax != PACK_MAX_NONE) {
      c.kmax = (uint32_t)(wmax >> 32);
      c.pmax = ~(uint32_t)wmax;
    }
    c.pnan = (wnan == POS_NONE) ? PACK_POS_NONE : (uint32_t)wnan;
  }
  return true;
}

// Shared-memory scratch of the CTA reduce: Best<Key>[32] (block_reduce), 5 x 32 words for the
// packed flow.
template <typename Key>
struct alignas(16) ReduceScratch {
  static constexpr size_t BYTES = sizeof(Best<Key>) * 32 > 640 ? sizeof(Best<Key>) * 32 : 640;
  unsigned char raw[BYTES];
  __device__ __forceinline__ Best<Key>* best() { return reinterpret_cast<Best<Key>*>(raw); }
  __device__ __forceinline__ uint32_t* words() { return reinterpret_cast<uint32_t*>(raw); }
};

// Index (within a 16-byte vector held in registers) of the first element whose key equals `want`
// -- the key of a non-NaN element known to be in the vector -- and of the first NaN.
template <class P, int NW>
__device__ __forceinline__ int locate_key(const uint32_t (&w)[NW], typename P::Key want) {
  constexpr int VE = NW * 4 / (int)sizeof(typename P::Elem);
  if constexpr (P::SIMD_EQ) {  // 8-bit ints, f16: packed equality compare + find-first-set
    return vec_first_eq<P>(w, want);
  } else {
    int found = 0;
#pragma unroll
    for (int e = VE - 1; e >= 0; --e) {
      typename P::Key k;
      bool valid, nan;
      P::classify(vec_elem<P, NW>(w, e), k, valid, nan);
      found = (valid && k == want) ? e : found;
    }
    return found;
  }
}
template <class P, int NW>
__device__ __forceinline__ int locate_nan(const uint32_t (&w)[NW]) {
  constexpr int VE = NW * 4 / (int)sizeof(typename P::Elem);
  int found = 0;
#pragma unroll
  for (int e = VE - 1; e >= 0; --e) {
    typename P::Key k;
    bool valid, nan;
    P::classify(vec_elem<P, NW>(w, e), k, valid, nan);
    found = nan ? e : found;
  }
  return found;
}

// out = w[u] for a run-time u (selects over the register array: no local memory, no branches)
template <int UN, int NW>
__device__ __forceinline__ void select_vec(const uint32_t (&w)[UN][NW], int u, uint32_t (&out)[NW]) {
#pragma unroll
  for (int q = 0; q < NW; ++q) {
    uint32_t v = w[0][q];
#pragma unroll
    for (int k = 1; k < UN; ++k) v = (u == k) ? w[k][q] : v;
    out[q] = v;
  }
}

// ---------------------------------------------------------------------------------------
// Small inputs (up to a few MB): latency, not bandwidth, is what a caller sees.  ONE trip: every
// thread loads its 1..4 vectors (16 B each, `direct_vpt` per thread, all in flight at once) and
// keeps them in registers; value-only fold per vector exactly as in the streaming kernel, the
// best vector per direction is remembered by its register slot, and the first matching element is
// then located IN REGISTERS -- no index is carried per element and nothing is re-read -- and
// LAZILY: only by the one or two lanes of a warp that hold the warp's best key (all 32 lanes
// walking their four vectors cost 0.3-0.9 us per call).  The dependent chain is
// load -> fold -> warp redux -> locate -> CTA reduce -> grid step -> publish.
// ---------------------------------------------------------------------------------------
template <class P, bool XCHG, bool PACKED = false>
__global__ void __launch_bounds__(DIRECT_MAX_THREADS) argminmax_direct_kernel(const ScanParams p) {
  using Key = typename P::Key;
  using Elem = typename P::Elem;
  constexpr int VEC = 16, NW = 4, UN = 4;
  constexpr int VE = VEC / sizeof(Elem);
  __shared__ ReduceScratch<Key> scratch;
  pdl_entry(p);
  stamp(p, ST_ENTRY);
  const uint8_t* main_base = p.base + p.head_elems * sizeof(Elem);
  const Bounds bd(p);
  const uint32_t vpt = p.direct_vpt;  // vectors per thread, 1..UN
  const uint64_t v0 = (uint64_t)blockIdx.x * blockDim.x * vpt + threadIdx.x;
  uint32_t w[UN][NW];
  bool have[UN];
#pragma unroll
  for (int u = 0; u < UN; ++u) {
    const uint64_t vi = v0 + (uint64_t)u * blockDim.x;
    have[u] = (uint32_t)u < vpt && vi < p.n_vec;
    if (have[u]) {
      bd.check(main_base + vi * VEC, VEC);
      VecLoad<VEC>::ld(main_base + vi * VEC, w[u]);
    } else {
#pragma unroll
      for (int q = 0; q < NW; ++q) w[u][q] = 0u;  // never selected (select_vec), but defined
    }
  }
  // value-only fold per vector; vectors ascend with u, so a strict compare keeps the first
  Key bkmin = KeyLim<Key>::MAXV, bkmax = KeyLim<Key>::MINV;
  int umin = -1, umax = -1, unan = -1;
#pragma unroll
  for (int u = 0; u < UN; ++u) {
    if (have[u]) {
      typename P::Acc acc;
      P::init(acc, &w[u][0]);
#pragma unroll
      for (int q = P::WPI; q < NW; q += P::WPI) P::feed(acc, &w[u][q]);
      Key kmin, kmax;
      bool valid, has_nan;
      P::finish(acc, kmin, kmax, valid, has_nan);
      const bool um = valid && (umin < 0 || kmin < bkmin);
      const bool uM = valid && (umax < 0 || kmax > bkmax);
      bkmin = um ? kmin : bkmin;
      umin = um ? u : umin;
      bkmax = uM ? kmax : bkmax;
      umax = uM ? u : umax;
      if (P::TRACK_NAN) unan = (has_nan && unan < 0) ? u : unan;
    }
  }
  stamp(p, ST_FIRST_FOLD);
  // Locate the first matching element inside the winning vectors (registers only) -> (key, element
  // index) candidates in `x`.  One vector per thread (the usual latency case): one fused pass over
  // its elements.  Several: the winning vector is picked out of the register array with selects, so
  // there is one copy of the locate code per direction (code size is latency here); with three or
  // four vectors the warp first agrees on its best keys (one or two redux each) and only the lanes
  // that hold one go on (LAZY: 1 Mi f32 -0.4..-0.6 us per call); with two, every lane locates in
  // its own winners (the redux round trips in front would cost more than they save).
  Best<Key> x;
  x.clear();
  {
    const uint64_t i0 = p.head_elems + v0 * VE;  // element index of vector u: i0 + u * blockDim * VE
    const bool hv = umin >= 0;                   // umin and umax are set together
    const uint32_t vnan = unan >= 0 ? (uint32_t)v0 + (uint32_t)unan * blockDim.x : PACK_POS_NONE;
    if (vpt == 1) {
      // the usual latency case, one vector per thread: one fused pass over its elements
      if (hv || (P::TRACK_NAN && unan >= 0)) {
        int emin = 0, emax = 0, enan = 0;
        if constexpr (P::SIMD_EQ) {
          emin = vec_first_eq<P>(w[0], bkmin);
          emax = vec_first_eq<P>(w[0], bkmax);
          if (P::TRACK_NAN) enan = locate_nan<P, NW>(w[0]);
        } else {
#pragma unroll
          for (int e = VE - 1; e >= 0; --e) {
            Key k;
            bool valid, nan;
            P::classify(vec_elem<P, NW>(w[0], e), k, valid, nan);
            emin = (valid && k == bkmin) ? e : emin;
            emax = (valid && k == bkmax) ? e : emax;
            if (P::TRACK_NAN) enan = nan ? e : enan;
          }
        }
        if (hv) {
          x.kmin = bkmin;
          x.pmin = i0 + (uint64_t)emin;
          x.kmax = bkmax;
          x.pmax = i0 + (uint64_t)emax;
        }
        if (P::TRACK_NAN && unan >= 0) x.pnan = i0 + (uint64_t)enan;
      }
    } else {
      Key wkmin = bkmin, wkmax = bkmax;
      uint32_t vfirst = vnan;
      if (vpt >= 3) {
        wkmin = warp_min_key(hv, bkmin);
        wkmax = warp_max_key(hv, bkmax);
        // vectors ascend with the address: the lane whose NaN vector comes first holds the first NaN
        if (P::TRACK_NAN) vfirst = __reduce_min_sync(0xFFFFFFFFu, vnan);
      }
      uint32_t sel[NW];
      if (hv && bkmin == wkmin) {
        select_vec<UN, NW>(w, umin, sel);
        x.kmin = bkmin;
        x.pmin = i0 + (uint64_t)umin * blockDim.x * VE + (uint64_t)locate_key<P, NW>(sel, bkmin);
      }
      if (hv && bkmax == wkmax) {
        select_vec<UN, NW>(w, umax, sel);
        x.kmax = bkmax;
        x.pmax = i0 + (uint64_t)umax * blockDim.x * VE + (uint64_t)locate_key<P, NW>(sel, bkmax);
      }
      if (P::TRACK_NAN && unan >= 0 && vnan == vfirst) {
        select_vec<UN, NW>(w, unan, sel);
        x.pnan = i0 + (uint64_t)unan * blockDim.x * VE + (uint64_t)locate_nan<P, NW>(sel);
      }
    }
  }
  // unaligned head / sub-vector tail (< 16 B each, first / last CTA): one element per thread,
  // lexicographic merges
  {
    const bool in_head = blockIdx.x == 0 && threadIdx.x < p.head_elems;
    const uint64_t it = p.tail_start + threadIdx.x;
    const bool in_tail = blockIdx.x == gridDim.x - 1 && it < p.n;
#pragma unroll 1
    for (int part = 0; part < 2; ++part) {
      const bool on = part == 0 ? in_head : in_tail;
      const uint64_t i = part == 0 ? (uint64_t)threadIdx.x : it;
      if (on) {
        const Elem* e = reinterpret_cast<const Elem*>(p.base);
        bd.check(&e[i], sizeof(Elem));
        Key k;
        bool valid, is_nan;
        P::classify(e[i], k, valid, is_nan);
        if (valid) {
          x.merge_min(k, i);
          x.merge_max(k, i);
        }
        if (P::TRACK_NAN && is_nan && i < x.pnan) x.pnan = i;
      }
    }
  }
  stamp(p, ST_STREAM_DONE);
  if constexpr (PACKED) {
    static_assert(sizeof(Key) == 4, "packed candidates need 32-bit keys");
    PackedCand c;
    c.clear();
    if (x.pmin != POS_NONE) c.kmin = (uint32_t)x.kmin ^ 0x80000000u, c.pmin = (uint32_t)x.pmin;
    if (x.pmax != POS_NONE) c.kmax = (uint32_t)x.kmax ^ 0x80000000u, c.pmax = (uint32_t)x.pmax;
    if (P::TRACK_NAN && x.pnan != POS_NONE) c.pnan = (uint32_t)x.pnan;
    if (!packed_grid_reduce<P>(p, c, scratch.words())) return;
    stamp(p, ST_GRID_FOLDED);
    stamp(p, ST_EXACT_DONE);
    finalize_warp<XCHG>(p, c.unpack());
  } else {
    x = block_reduce(x, scratch.best());
    stamp(p, ST_CTA_REDUCED);
    if (!grid_reduce(p, x, scratch.best())) return;
    stamp(p, ST_GRID_FOLDED);
    stamp(p, ST_EXACT_DONE);
    finalize_warp<XCHG>(p, x);
  }
}

// ---------------------------------------------------------------------------------------
// Exact finish, executed by WARP 0 of the finishing CTA (all 32 lanes; the result is valid in lane 0):
// the winning min / max / first-NaN vectors (`b` holds their VECTOR indices; <= 32 elements each)
// are re-read with one element per lane and a ballot picks the first match; the unaligned head
// and the sub-vector tail (< one vector each) ride along as exact per-lane candidates.  All five
// loads are issued before the first one is consumed: one memory round trip, not five.
// ---------------------------------------------------------------------------------------
template <class P, uint32_t VEC_ELEMS>
__device__ __forceinline__ Best<typename P::Key> exact_finish_warp(const ScanParams& p, const Bounds& bd,
                                                                   const Best<typename P::Key>& b) {
  using Elem = typename P::Elem;
  using Key = typename P::Key;
  static_assert(VEC_ELEMS <= 32, "one element per lane");
  const Elem* e = reinterpret_cast<const Elem*>(p.base);
  const uint32_t lane = threadIdx.x & 31;
  constexpr int NC = P::TRACK_NAN ? 3 : 2;
  const uint64_t cand[3] = {b.pmin, b.pmax, b.pnan};
  Elem raw[5];
  bool have[5];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    have[c] = cand[c] != POS_NONE && lane < VEC_ELEMS;
    const uint64_t i = p.head_elems + cand[c] * VEC_ELEMS + lane;
    if (have[c]) bd.check(&e[i], sizeof(Elem));
    raw[c] = have[c] ? e[i] : Elem(0);
  }
  // head: elements [0, head_elems); tail: [tail_start, n); both shorter than one vector
  const uint64_t ih = lane, it = p.tail_start + lane;
  have[3] = ih < p.head_elems;
  have[4] = it < p.n;
  if (have[3]) bd.check(&e[ih], sizeof(Elem));
  if (have[4]) bd.check(&e[it], sizeof(Elem));
  raw[3] = have[3] ? e[ih] : Elem(0);
  raw[4] = have[4] ? e[it] : Elem(0);

  Best<Key> x;
  x.clear();
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    Key k;
    bool valid, is_nan;
    P::classify(raw[c], k, valid, is_nan);
    const bool hit = have[c] && ((c == 0) ? (valid && k == b.kmin) : (c == 1) ? (valid && k == b.kmax) : is_nan);
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
    if (m != 0 && lane == 0) {
      const uint64_t idx = p.head_elems + cand[c] * VEC_ELEMS + (uint64_t)(__ffs(m) - 1);
      if (c == 0) {
        x.kmin = b.kmin;
        x.pmin = idx;
      } else if (c == 1) {
        x.kmax = b.kmax;
        x.pmax = idx;
      } else {
        x.pnan = idx;
      }
    }
  }
  if (p.head_elems != 0 || p.tail_start != p.n) {  // warp-uniform; the usual aligned call skips it
#pragma unroll
    for (int part = 3; part < 5; ++part) {
      const uint64_t i = part == 3 ? ih : it;
      if (have[part]) {
        Key k;
        bool valid, is_nan;
        P::classify(raw[part], k, valid, is_nan);
        if (valid) {
          x.merge_min(k, i);
          x.merge_max(k, i);
        }
        if (P::TRACK_NAN && is_nan && i < x.pnan) x.pnan = i;
      }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) x.merge(shfl_xor(x, m));
  }
  return x;
}

// ---------------------------------------------------------------------------------------
// The streaming kernel.  A CTA owns a sequence of TILES (blockDim x U vectors each):
//   grid-stride (per_vecs == 0, large inputs): tiles b, b + G, b + 2G, ... of the whole array;
//   contiguous (per_vecs != 0, small / medium inputs): its own range of per_vecs vectors, cut
//     into tiles -- every CTA gets the same amount to within one row of blockDim vectors, where
//     grid-stride hands out whole 16-32 KiB tiles (4 vs 5 tiles per CTA at 40 MB = 20 % idle).
// Local tile k lives at vector  cta_vec_base + k * kstride_vecs;  a SLOT id is k * U + j.
// The partial tile at the END of the CTA's sequence is loaded FIRST (so that its latency hides
// behind tile 0's loads) and folded into its own running state, merged last with "main wins
// ties" -- every main-loop vector of a thread precedes its remainder vectors.
// ---------------------------------------------------------------------------------------
template <typename Key>
__device__ __forceinline__ void running_clear(Running<Key>& r) {
  r.kmin = KeyLim<Key>::MAXV;
  r.kmax = KeyLim<Key>::MINV;
  r.smin = r.smax = r.snan = TILE_NONE;
}

template <class P, int VEC, int U, bool PREFETCH, int HINT, bool XCHG, bool PACKED = false>
__global__ void __launch_bounds__(MAX_THREADS) argminmax_scan_kernel(const ScanParams p) {
  using VL = VecLoad<VEC, HINT>;
  constexpr uint32_t VEC_ELEMS = VEC / sizeof(typename P::Elem);
  using Key = typename P::Key;
  constexpr int NW = VL::NW;
  static_assert(!PACKED || sizeof(Key) == 4, "packed candidates need 32-bit keys");
  __shared__ ReduceScratch<Key> scratch;
  pdl_entry(p);
  stamp(p, ST_ENTRY);

  const uint32_t tile_vecs = blockDim.x * U;
  const uint64_t tile_bytes = (uint64_t)tile_vecs * VEC;
  const uint8_t* main_base = p.base + p.head_elems * sizeof(typename P::Elem);
  const Bounds bd(p);

  // this CTA's tile sequence: n_full static tiles, then (dyn_tiles != 0) claimed ones, and at
  // most one partial tile of `rem` vectors at rem_vec_base
  uint64_t cta_vec_base, kstride_vecs, rem_vec_base;
  uint32_t n_full, rem;
  if (p.per_vecs != 0) {
    const uint64_t vb = (uint64_t)blockIdx.x * p.per_vecs;
    const uint64_t ve = vb + p.per_vecs < p.n_vec ? vb + p.per_vecs : p.n_vec;
    const uint64_t len = ve > vb ? ve - vb : 0;
    n_full = (uint32_t)(len / tile_vecs);
    rem = (uint32_t)(len % tile_vecs);
    cta_vec_base = vb;
    kstride_vecs = tile_vecs;
    rem_vec_base = vb + (uint64_t)n_full * tile_vecs;
  } else {
    const uint32_t nft = p.n_full_tiles;  // static tiles only
    n_full = blockIdx.x < nft ? (nft - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    rem = (p.rem_vecs != 0 && blockIdx.x == p.rem_owner) ? p.rem_vecs : 0;
    cta_vec_base = (uint64_t)blockIdx.x * tile_vecs;
    kstride_vecs = (uint64_t)gridDim.x * tile_vecs;
    rem_vec_base = (uint64_t)(nft + p.dyn_tiles) * tile_vecs;
  }
  const uint8_t* lane_base = main_base + (cta_vec_base + threadIdx.x) * VEC;
  const uint64_t kstride_bytes = kstride_vecs * VEC;
  const uint64_t jstride = (uint64_t)blockDim.x * VEC;

  // Dynamic tail (see launch_packed): the last dyn_tiles full tiles of the array are not
  // pre-assigned; a CTA that has finished its static share claims them one at a time from a
  // counter in the workspace, so the CTAs finish together although they progress at very
  // different rates (equal static shares finish 2.4x apart at 40 MB, 5 % apart at 4 GiB).
  // Claims are issued one tile ahead by thread 0 and broadcast through shared memory; claimed
  // tiles ascend per CTA and lie behind every static tile, so slot order stays address order.
  // The counter is one of a ring of 64, picked by the call's sequence number, and is re-armed by
  // the call's own last CTA.  Pipelined scans overlap, but call s + 64 cannot start claiming
  // before call s has completed: a pipelined CTA cannot exit before the previous grid is complete
  // (griddepcontrol.wait in the grid step), a grid only starts once every CTA of the one before
  // it has started, and 63 whole grids of at least one CTA per SM do not fit on the GPU (32 CTAs
  // per SM at most).
  __shared__ uint32_t s_claim[2];
  const uint32_t dyn_tiles = p.dyn_tiles;
  uint32_t* const dyn_counter = p.ticket + WS_CLAIM_RING + (uint32_t)(p.seq & 63ull);
  uint32_t pending = 0;
  if (dyn_tiles != 0 && threadIdx.x == 0) pending = atomicAdd(dyn_counter, 1u);
  auto next_tile = [&](uint32_t i, const uint8_t*& q, uint32_t& t) -> bool {
    if (i < n_full) {
      q = lane_base + (uint64_t)i * kstride_bytes;
      t = i;
      return true;
    }
    if (dyn_tiles == 0) return false;
    const uint32_t c = i - n_full;  // this CTA's c-th claim
    if (threadIdx.x == 0) s_claim[c & 1] = pending;
    __syncthreads();
    const uint32_t d = s_claim[c & 1];
    if (d >= dyn_tiles) return false;
    if (threadIdx.x == 0) pending = atomicAdd(dyn_counter, 1u);
    q = main_base + ((uint64_t)(p.n_full_tiles + d) * tile_vecs + threadIdx.x) * VEC;
    t = p.dyn_koff + d;
    return true;
  };

  Running<Key> run, run_rem;
  running_clear(run);
  running_clear(run_rem);

  {
    uint32_t a[U][NW];
    // ---------------- partial tile: its loads go out first
    uint32_t cur[U][NW];
    const bool mine = threadIdx.x < rem;  // rem != 0 and this thread owns at least row 0 of it
    if (mine) {
      const uint8_t* q = main_base + (rem_vec_base + threadIdx.x) * VEC;
      bd.check(q, VEC);
      VL::ld(q, cur[0]);
#pragma unroll
      for (int j = 1; j < U; ++j) {
        if ((uint32_t)j * blockDim.x + threadIdx.x < rem) {
          bd.check(q + j * jstride, VEC);
          VL::ld(q + j * jstride, cur[j]);
        }
      }
    }
    // ---------------- streaming phase: full tiles, two register buffers in ping-pong: the loads
    // of the next tile are in flight while the current one is folded (PREFETCH)
    uint32_t i = 0;
    const uint8_t* qa = nullptr;
    uint32_t ta = 0;
    bool va = false;
    if (PREFETCH) {
      va = next_tile(0, qa, ta);
      if (va) {
#pragma unroll
        for (int j = 0; j < U; ++j) {
          bd.check(qa + j * jstride, VEC);
          VL::ld(qa + j * jstride, a[j]);
        }
      }
    }
    if (mine) {
#pragma unroll
      for (int j = 1; j < U; ++j) {
        if (!((uint32_t)j * blockDim.x + threadIdx.x < rem)) {  // duplicates are neutral for a value-only min/max
#pragma unroll
          for (int w = 0; w < NW; ++w) cur[j][w] = cur[0][w];
        }
      }
      fold_tile<P, NW, U>(cur, 0, run_rem);
    }
    if (PREFETCH) {
      uint32_t b[U][NW];
      while (va) {
        const uint8_t* qb = nullptr;
        uint32_t tb = 0;
        const bool vb = next_tile(i + 1, qb, tb);
        if (vb) {
#pragma unroll
          for (int j = 0; j < U; ++j) {
            bd.check(qb + j * jstride, VEC);
            VL::ld(qb + j * jstride, b[j]);
          }
        }
        fold_tile<P, NW, U>(a, ta, run);
#ifdef AMM_PHASE_STAMPS
        if (i == 0) stamp(p, ST_FIRST_FOLD);
#endif
        if (!vb) break;
        va = next_tile(i + 2, qa, ta);
        if (va) {
#pragma unroll
          for (int j = 0; j < U; ++j) {
            bd.check(qa + j * jstride, VEC);
            VL::ld(qa + j * jstride, a[j]);
          }
        }
        fold_tile<P, NW, U>(b, tb, run);
        i += 2;
      }
    } else {
      while (next_tile(i, qa, ta)) {
#pragma unroll
        for (int j = 0; j < U; ++j) {
          bd.check(qa + j * jstride, VEC);
          VL::ld(qa + j * jstride, a[j]);
        }
        fold_tile<P, NW, U>(a, ta, run);
        ++i;
      }
    }
  }
  stamp(p, ST_STREAM_DONE);

  // ------------------- slot ids -> global vector indices; the partial tile holds this thread's
  // LAST vectors, so the main state wins ties
  auto slot_to_vec = [&](uint32_t slot) -> uint64_t {
    if (slot == TILE_NONE) return POS_NONE;
    const uint32_t tl = slot / U, j = slot % U;
    const uint64_t tile_base = tl < p.dyn_koff ? cta_vec_base + (uint64_t)tl * kstride_vecs
                                               : (uint64_t)(p.n_full_tiles + (tl - p.dyn_koff)) * tile_vecs;
    return tile_base + (uint64_t)j * blockDim.x + threadIdx.x;
  };
  auto rem_slot_to_vec = [&](uint32_t slot) -> uint64_t {
    return slot == TILE_NONE ? POS_NONE : rem_vec_base + (uint64_t)slot * blockDim.x + threadIdx.x;
  };
  Key kmin = run.kmin, kmax = run.kmax;
  uint64_t vmin = slot_to_vec(run.smin), vmax = slot_to_vec(run.smax), vnan = slot_to_vec(run.snan);
  {
    const bool um = run_rem.smin != TILE_NONE && (run.smin == TILE_NONE || run_rem.kmin < run.kmin);
    const bool uM = run_rem.smax != TILE_NONE && (run.smax == TILE_NONE || run_rem.kmax > run.kmax);
    kmin = um ? run_rem.kmin : kmin;
    vmin = um ? rem_slot_to_vec(run_rem.smin) : vmin;
    kmax = uM ? run_rem.kmax : kmax;
    vmax = uM ? rem_slot_to_vec(run_rem.smax) : vmax;
    vnan = run.snan == TILE_NONE ? rem_slot_to_vec(run_rem.snan) : vnan;
  }

  // ------------------- CTA + grid reduce on (key, global vector index); only warp 0 of the
  // finishing CTA (last ticket / CTA 0) goes on, with the grid-wide candidates in `b`
  Best<Key> b;
  if constexpr (PACKED) {
    // 32-bit pieces (launch_one guarantees fewer than 2^32 - 2 vectors)
    PackedCand c;
    c.clear();
    if (vmin != POS_NONE)  // min and max are set together (fold_tile)
      c.set(kmin, (uint32_t)vmin, kmax, (uint32_t)vmax);
    if (P::TRACK_NAN && vnan != POS_NONE) c.pnan = (uint32_t)vnan;
    if (!packed_grid_reduce<P>(p, c, scratch.words())) return;
    b = c.unpack();
  } else {
    b.kmin = kmin;
    b.kmax = kmax;
    b.pmin = vmin;
    b.pmax = vmax;
    b.pnan = vnan;
    b = block_reduce(b, scratch.best());
    stamp(p, ST_CTA_REDUCED);
    if (!grid_reduce(p, b, scratch.best())) return;
  }
  // every CTA has passed the grid step, hence made its last claim: re-arm the claim counter
  if (dyn_tiles != 0 && threadIdx.x == 0) *dyn_counter = 0u;
  stamp(p, ST_GRID_FOLDED);
  const Best<Key> x = exact_finish_warp<P, VEC_ELEMS>(p, bd, b);
  stamp(p, ST_EXACT_DONE);
  finalize_warp<XCHG>(p, x);
}

}  // namespace amm
