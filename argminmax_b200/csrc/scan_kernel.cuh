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
//  * CTA reduce (warp shuffles, then shared memory) and the cross-CTA reduce
//    (last-block-done ticket) are lexicographic on (key, global vector index): the
//    earliest vector that contains the global extreme wins.
//  * One warp of the last CTA finishes exactly: the winning min / max / first-NaN
//    vectors (<= 32 elements each) are re-read with one element per lane and a ballot
//    picks the first match; the unaligned head and the sub-vector tail of the array
//    ride along as per-lane candidates.  That yields first-occurrence indices with the
//    reference's tie-break rule (src/simd/task.rs:273,290) and re-reads < 200 bytes.
//  * One 64-byte record (keys, global indices, first NaN, flags) is written to the
//    workspace's device slot and to a pinned host slot: one launch, no memcpy.
#pragma once
#include <stdint.h>

#include "../../include/argminmax_b200.h"
#include "policies.cuh"

namespace amm {

constexpr uint32_t TILE_NONE = 0xFFFFFFFFu;  // "no slot yet" in the per-thread running state
constexpr uint64_t POS_NONE = 0xFFFFFFFFFFFFFFFFull;
constexpr int MAX_THREADS = 512;

struct ScanParams {
  const uint8_t* base;     // caller's device pointer (element 0)
  uint64_t n;              // elements
  uint64_t n_bytes;        // n * sizeof(element)
  uint64_t global_offset;  // added to every reported index (shard offset)
  uint64_t head_elems;     // elements before the first VEC-aligned address (<= n)
  uint64_t n_vec;          // whole vectors in the aligned main region
  uint64_t tail_start;     // first element after the main region
  uint32_t n_full_tiles;   // n_vec / (blockDim.x * U)
  uint32_t rem_vecs;       // n_vec % (blockDim.x * U)
  void* partials;          // gridDim.x x Best<Key>
  uint32_t* ticket;        // zero-initialised, self-resetting (atomicInc wrap)
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
};

// Debug builds (-DAMM_BOUNDS_CHECK, argminmax_b200/_build.py:build_boundscheck) verify that every global load of input
// data lies inside [base, base + n_bytes); violations are counted in a scratch word and
// surface as AMM_REC_BOUNDS in the record (compute-sanitizer is not available on the pool).
struct Bounds {
#ifdef AMM_BOUNDS_CHECK
  const uint8_t* lo;
  const uint8_t* hi;
  uint32_t* err;
  __device__ __forceinline__ explicit Bounds(const ScanParams& p)
      : lo(p.base), hi(p.base + p.n_bytes), err(p.ticket + 8) {}
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

constexpr int XCHG_SLOT_WORDS = 16;     // 128 B per (parity, rank) slot: 8 record words,
constexpr int XCHG_MAX_WORLD = 16;      // checksum, tag; 2 parities per buffer
constexpr uint64_t AMM_REC_PEER_TIMEOUT = 4ull;  // record.flags bit: a peer never showed up

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

// All threads of the CTA must call; every thread returns the CTA-wide result.
template <typename Key>
__device__ __forceinline__ Best<Key> block_reduce(Best<Key> b, Best<Key>* smem /* [32] */) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) b.merge(shfl_xor(b, m));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  __syncthreads();  // smem may still be read from a previous call
  if (lane == 0) smem[warp] = b;
  __syncthreads();
  if (warp == 0) {
    Best<Key> t;
    t.clear();
    if (lane < nwarps) t = smem[lane];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) t.merge(shfl_xor(t, m));
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
// checksum matches, so a torn / partially arrived record is simply polled again.
__device__ __forceinline__ uint64_t record_checksum(const amm_record& r) {
  return r.min_key ^ r.min_idx ^ r.max_key ^ r.max_idx ^ r.nan_idx ^ r.flags ^ r.n ^ r.seq ^
         0xA5A5C3C3F00DBEEFull;
}

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
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(h + 8), "l"(record_checksum(r)) : "memory");
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
  if (atomicExch(p.ticket + 8, 0u) != 0u) r.flags |= AMM_REC_BOUNDS;
#endif
  r.n = p.n;
  r.seq = p.seq;
  return r;
}

// Last step of the last CTA.  world == 1: publish this GPU's record.  world > 1: the
// compute step is FUSED with its collective -- instead of returning to the host for an NCCL
// all-gather, the last CTA stores its 64-byte record straight into every peer's exchange
// buffer over NVLink (peer-mapped pointers, one thread per peer), then polls its own buffer
// until all `world` records of this call (tag == xseq, checksum valid) have landed, folds them
// in ascending rank order with strict compares (= global first occurrence, the reference's
// merge rule src/simd/generic.rs:513-520) and publishes the GLOBAL record.  Two parity banks
// make back-to-back calls safe: a rank can only reach call s+2 after every peer has finished
// reading call s.  A peer that never arrives trips a ~4 s timeout and is reported in flags.
// XCHG is a compile-time flag of the kernels: single-GPU launches use instantiations that do
// not contain this code at all (its mere presence costs the streaming loop ~3 % -- registers,
// loss of uniform-datapath loop control), multi-GPU launches use the fused instantiations.
template <typename Key>
__device__ __forceinline__ void exchange_and_publish(const ScanParams& p, const Best<Key>& x) {
  __shared__ amm_record s_mine;
  __shared__ amm_record s_all[XCHG_MAX_WORLD];
  __shared__ uint32_t s_timeout;
  if (threadIdx.x == 0) {
    s_mine = make_record(p, x);
    s_timeout = 0;
  }
  __syncthreads();
  const uint64_t bank = (p.xseq & 1ull) * (uint64_t)p.world;
  if ((int)threadIdx.x < p.world) {
    // push my record into peer `t`'s buffer, slot [bank + rank]
    const int t = threadIdx.x;
    volatile uint64_t* dst = p.peer_slots[t] + (bank + (uint64_t)p.rank) * XCHG_SLOT_WORDS;
    const uint64_t* src = reinterpret_cast<const uint64_t*>(&s_mine);
    uint64_t c = 0x5EEDFACE0DDBA11ull ^ p.xseq;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dst[i] = src[i];
      c ^= src[i];
    }
    dst[8] = c;
    dst[9] = p.xseq;
    __threadfence_system();
    // pull peer `t`'s record out of MY buffer, slot [bank + t]
    const volatile uint64_t* in = p.peer_slots[p.rank] + (bank + (uint64_t)t) * XCHG_SLOT_WORDS;
    uint64_t w[10];
    unsigned long long t0 = 0, now = 0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    bool ok = false;
    while (!ok) {
      if (in[9] == p.xseq) {
#pragma unroll
        for (int i = 0; i < 10; ++i) w[i] = in[i];
        uint64_t cc = 0x5EEDFACE0DDBA11ull ^ p.xseq;
#pragma unroll
        for (int i = 0; i < 8; ++i) cc ^= w[i];
        ok = (cc == w[8]) && (w[9] == p.xseq);
      }
      if (!ok) {
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (now - t0 > 4000000000ull) {  // ns
          atomicExch(&s_timeout, 1u);
          break;
        }
      }
    }
    uint64_t* out = reinterpret_cast<uint64_t*>(&s_all[t]);
    if (ok) {
#pragma unroll
      for (int i = 0; i < 8; ++i) out[i] = w[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) out[i] = 0;  // no value, no NaN
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    amm_record o;
    bool have = false;
    o.min_key = o.max_key = 0;
    o.min_idx = o.max_idx = o.nan_idx = AMM_NO_INDEX;
    o.n = 0;
    for (int g = 0; g < p.world; ++g) {
      const amm_record r = s_all[g];
      o.n += r.n;
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
              (o.nan_idx != AMM_NO_INDEX ? AMM_REC_HAS_NAN : 0ull) |
              (s_timeout ? AMM_REC_PEER_TIMEOUT : 0ull);
    o.seq = p.seq;
    publish_record(p, o);
  }
}

template <bool XCHG, typename Key>
__device__ __forceinline__ void finalize(const ScanParams& p, const Best<Key>& x) {
  if (XCHG) {
    exchange_and_publish(p, x);
  } else {
    if (threadIdx.x == 0) publish_record(p, make_record(p, x));
  }
}

// Programmatic dependent launch (PDL).  Scans are launched with the programmatic-stream-
// serialization attribute and every CTA signals launch_dependents on entry, so the NEXT scan
// on the stream can be scheduled while this one drains.  Default (safe) mode: a CTA executes
// griddepcontrol.wait before its first global load, i.e. only launch latency is hidden.
// Pipelined mode (amm_workspace_set_pipelined, caller promises that a scan's input is not
// produced by the kernel right before it on the stream): the wait moves to the point where
// the CTA first touches the shared workspace (partials, ticket, record slots), so the
// STREAMING phase of scan k+1 overlaps the fold / locate / publish tail of scan k.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait_prior_grid() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Cross-CTA reduce: every CTA deposits its Best, the last one to arrive (ticket) folds them.
// Returns true in the last CTA only, with `b` = grid-wide result.  gridDim.x == 1 skips the
// round trip through global memory.
template <typename Key>
__device__ __forceinline__ bool grid_reduce(const ScanParams& p, Best<Key>& b, Best<Key>* smem,
                                            uint32_t* s_is_last) {
  pdl_wait_prior_grid();  // from here on the workspace is ours
  if (gridDim.x == 1) return true;
  Best<Key>* partials = reinterpret_cast<Best<Key>*>(p.partials);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = b;
    __threadfence();
    const uint32_t ticket = atomicInc(p.ticket, gridDim.x - 1);  // wraps to 0: reusable
    *s_is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!*s_is_last) return false;
  __threadfence();
  Best<Key> g;
  g.clear();
  for (uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
    Best<Key> o;
    const Best<Key>* src = partials + i;
    o.kmin = __ldcg(&src->kmin);
    o.kmax = __ldcg(&src->kmax);
    o.pmin = __ldcg(&src->pmin);
    o.pmax = __ldcg(&src->pmax);
    o.pnan = __ldcg(&src->pnan);
    g.merge(o);
  }
  b = block_reduce(g, smem);
  return true;
}

// ---------------------------------------------------------------------------------------
// Small inputs (a few MB): latency, not bandwidth, is what a caller sees.  One pass with
// exact per-element (key, index) tracking -- no tile bookkeeping, no second phase: each CTA
// takes a contiguous slice, CTA reduce, ticket, the last CTA folds the partials and
// publishes.  The dependent chain is load -> reduce -> ticket -> fold -> publish.
// ---------------------------------------------------------------------------------------
// Cross-CTA step of the latency kernel for 32-bit keys and fewer than 2^32 - 2 elements: a
// candidate is ONE 64-bit word, (biased key << 32) | index for the minimum and
// (biased key << 32) | ~index for the maximum, so "smaller key, then smaller index" is a plain
// unsigned min and "larger key, then smaller index" a plain unsigned max.  Every CTA folds its
// word into the workspace with one atomicMin / atomicMax (+ one for the first NaN index) and takes
// a ticket; the last CTA swaps the words back to their identities and publishes.  Compared with
// "store partials, ticket, last CTA loads and reduces them" this removes one global round trip
// and one CTA-wide reduce from the dependent chain (tools/exp/latency_floor.cu: 9.4 -> 8.1 us
// for the skeleton of a 4 MiB scan).  Words: p.ticket + 16 (min), + 18 (max), + 20 (NaN).
constexpr unsigned long long PACK_MIN_NONE = ~0ull;
constexpr unsigned long long PACK_MAX_NONE = 0ull;
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
  return __shfl_xor_sync(0xFFFFFFFFu, v, m);
}
struct PackedWords {
  unsigned long long wmin, wmax, wnan;
  __device__ __forceinline__ void clear() {
    wmin = PACK_MIN_NONE;
    wmax = PACK_MAX_NONE;
    wnan = POS_NONE;
  }
  __device__ __forceinline__ void set_min(int32_t key, uint32_t pos) {
    wmin = ((unsigned long long)((uint32_t)key ^ 0x80000000u) << 32) | pos;
  }
  __device__ __forceinline__ void set_max(int32_t key, uint32_t pos) {
    wmax = ((unsigned long long)((uint32_t)key ^ 0x80000000u) << 32) | (uint32_t)~pos;
  }
  // -> Best with 32-bit positions widened
  __device__ __forceinline__ Best<int32_t> unpack() const {
    Best<int32_t> g;
    g.clear();
    if (wmin != PACK_MIN_NONE) {
      g.kmin = (int32_t)((uint32_t)(wmin >> 32) ^ 0x80000000u);
      g.pmin = (uint32_t)wmin;
    }
    if (wmax != PACK_MAX_NONE) {
      g.kmax = (int32_t)((uint32_t)(wmax >> 32) ^ 0x80000000u);
      g.pmax = (uint32_t)~(uint32_t)wmax;
    }
    g.pnan = wnan;
    return g;
  }
};

// Folds every thread's words over the CTA and then over the grid.  Returns true in ALL lanes of
// warp 0 of the last CTA to arrive (or of the only CTA), with the grid-wide words in `w`; every
// other thread gets false.  All threads of the CTA must call (one __syncthreads inside).
template <class P>
__device__ __forceinline__ bool packed_grid_reduce(const ScanParams& p, PackedWords& w,
                                                   unsigned long long* smem /* [3][32] */) {
  auto warp_fold = [&]() {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      const unsigned long long a = shfl_xor_u64(w.wmin, m), b = shfl_xor_u64(w.wmax, m);
      w.wmin = a < w.wmin ? a : w.wmin;
      w.wmax = b > w.wmax ? b : w.wmax;
      if (P::TRACK_NAN) {
        const unsigned long long c = shfl_xor_u64(w.wnan, m);
        w.wnan = c < w.wnan ? c : w.wnan;
      }
    }
  };
  warp_fold();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  if (lane == 0) {
    smem[warp] = w.wmin;
    smem[32 + warp] = w.wmax;
    if (P::TRACK_NAN) smem[64 + warp] = w.wnan;
  }
  __syncthreads();
  if (warp != 0) return false;
  w.wmin = lane < nwarps ? smem[lane] : PACK_MIN_NONE;
  w.wmax = lane < nwarps ? smem[32 + lane] : PACK_MAX_NONE;
  w.wnan = (P::TRACK_NAN && lane < nwarps) ? smem[64 + lane] : POS_NONE;
  warp_fold();
  uint32_t last = 1;
  if (lane == 0) {
    pdl_wait_prior_grid();  // from here on the workspace is ours
    if (gridDim.x > 1) {
      unsigned long long* words = reinterpret_cast<unsigned long long*>(p.ticket + 16);
      if (w.wmin != PACK_MIN_NONE) atomicMin(&words[0], w.wmin);
      if (w.wmax != PACK_MAX_NONE) atomicMax(&words[1], w.wmax);
      if (P::TRACK_NAN && w.wnan != POS_NONE) atomicMin(&words[2], w.wnan);
      __threadfence();
      last = atomicInc(p.ticket, gridDim.x - 1) == gridDim.x - 1;
      if (last) {
        __threadfence();
        w.wmin = atomicExch(&words[0], PACK_MIN_NONE);  // read the grid-wide result and re-arm
        w.wmax = atomicExch(&words[1], PACK_MAX_NONE);
        if (P::TRACK_NAN) w.wnan = atomicExch(&words[2], POS_NONE);
      }
    }
  }
  last = __shfl_sync(0xFFFFFFFFu, last, 0);
  if (!last) return false;
  w.wmin = __shfl_sync(0xFFFFFFFFu, w.wmin, 0);
  w.wmax = __shfl_sync(0xFFFFFFFFu, w.wmax, 0);
  w.wnan = __shfl_sync(0xFFFFFFFFu, w.wnan, 0);
  return true;
}

template <class P, bool XCHG, bool PACKED = false>
__global__ void __launch_bounds__(MAX_THREADS) argminmax_direct_kernel(const ScanParams p) {
  using Key = typename P::Key;
  constexpr int VEC = 16, NW = 4, UN = 4;
  constexpr int VE = VEC / sizeof(typename P::Elem);
  __shared__ Best<Key> smem[32];
  __shared__ uint32_t s_is_last;
  pdl_launch_dependents();
  if (!p.pipelined) pdl_wait_prior_grid();
  const uint8_t* main_base = p.base + p.head_elems * sizeof(typename P::Elem);
  const Bounds bd(p);
  // contiguous slice of whole vectors per CTA, coalesced 128-bit loads, UN in flight
  const uint64_t per = (p.n_vec + gridDim.x - 1) / gridDim.x;
  const uint64_t vb = per * blockIdx.x;
  uint64_t ve = vb + per;
  ve = ve < p.n_vec ? ve : p.n_vec;
  Best<Key> x;
  x.clear();
  for (uint64_t v0 = vb + threadIdx.x; v0 < ve; v0 += (uint64_t)UN * blockDim.x) {
    uint32_t w[UN][NW];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const uint64_t vi = v0 + (uint64_t)u * blockDim.x;
      if (vi < ve) {
        bd.check(main_base + vi * VEC, VEC);
        VecLoad<VEC>::ld(main_base + vi * VEC, w[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const uint64_t vi = v0 + (uint64_t)u * blockDim.x;
      if (vi < ve) {
        const uint64_t i0 = p.head_elems + vi * VE;
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          Key k;
          bool valid, is_nan;
          P::classify(vec_elem<P, NW>(w[u], e), k, valid, is_nan);
          // indices ascend within a thread: a strict compare keeps the first occurrence
          if (valid && (k < x.kmin || x.pmin == POS_NONE)) {
            x.kmin = k;
            x.pmin = i0 + e;
          }
          if (valid && (k > x.kmax || x.pmax == POS_NONE)) {
            x.kmax = k;
            x.pmax = i0 + e;
          }
          if (P::TRACK_NAN && is_nan && x.pnan == POS_NONE) x.pnan = i0 + e;
        }
      }
    }
  }
  // unaligned head / sub-vector tail (< 16 B each): lexicographic merges, any order
  if (blockIdx.x == 0) exact_range<P, false>(bd, p.base, 0, p.head_elems, x);
  if (blockIdx.x == gridDim.x - 1) exact_range<P, false>(bd, p.base, p.tail_start, p.n, x);
  if constexpr (PACKED) {
    static_assert(sizeof(Key) == 4 && !XCHG, "packed candidates need 32-bit keys");
    PackedWords w;
    w.clear();
    if (x.pmin != POS_NONE) w.set_min(x.kmin, (uint32_t)x.pmin);
    if (x.pmax != POS_NONE) w.set_max(x.kmax, (uint32_t)x.pmax);
    if (P::TRACK_NAN) w.wnan = x.pnan;
    if (!packed_grid_reduce<P>(p, w, reinterpret_cast<unsigned long long*>(smem))) return;
    if (threadIdx.x == 0) publish_record(p, make_record(p, w.unpack()));
  } else {
    x = block_reduce(x, smem);
    if (!grid_reduce(p, x, smem, &s_is_last)) return;
    finalize<XCHG>(p, x);
  }
}

template <class P, int VEC, int U, bool PREFETCH, int HINT, bool XCHG>
__global__ void __launch_bounds__(MAX_THREADS) argminmax_scan_kernel(const ScanParams p) {
  using VL = VecLoad<VEC, HINT>;
  constexpr uint32_t VEC_ELEMS = VEC / sizeof(typename P::Elem);
  using Key = typename P::Key;
  constexpr int NW = VL::NW;
  __shared__ Best<Key> smem[32];
  __shared__ uint32_t s_is_last;
  pdl_launch_dependents();
  if (!p.pipelined) pdl_wait_prior_grid();

  const uint32_t tile_vecs = blockDim.x * U;
  const uint64_t tile_bytes = (uint64_t)tile_vecs * VEC;
  const uint8_t* main_base = p.base + p.head_elems * sizeof(typename P::Elem);
  const Bounds bd(p);

  Running<Key> run;
  run.kmin = KeyLim<Key>::MAXV;
  run.kmax = KeyLim<Key>::MINV;
  run.smin = run.smax = run.snan = TILE_NONE;

  // ------------------------------------------------ streaming phase: full tiles
  // Two register buffers in ping-pong: the loads of the next tile are in flight
  // while the current one is folded (PREFETCH), without register-to-register moves.
  {
    const uint32_t nft = p.n_full_tiles;
    const uint8_t* lane_base = main_base + (uint64_t)threadIdx.x * VEC;
    const uint64_t jstride = (uint64_t)blockDim.x * VEC;
    uint32_t t = blockIdx.x;
    uint32_t a[U][NW];
    if (PREFETCH) {
      uint32_t b[U][NW];
      if (t < nft) {
        const uint8_t* q = lane_base + (uint64_t)t * tile_bytes;
#pragma unroll
        for (int j = 0; j < U; ++j) {
          bd.check(q + j * jstride, VEC);
          VL::ld(q + j * jstride, a[j]);
        }
      }
      while (t < nft) {
        const uint32_t t1 = t + gridDim.x;
        if (t1 < nft) {
          const uint8_t* q = lane_base + (uint64_t)t1 * tile_bytes;
#pragma unroll
          for (int j = 0; j < U; ++j) {
            bd.check(q + j * jstride, VEC);
            VL::ld(q + j * jstride, b[j]);
          }
        }
        fold_tile<P, NW, U>(a, t, run);
        if (t1 >= nft) break;
        const uint32_t t2 = t1 + gridDim.x;
        if (t2 < nft) {
          const uint8_t* q = lane_base + (uint64_t)t2 * tile_bytes;
#pragma unroll
          for (int j = 0; j < U; ++j) {
          bd.check(q + j * jstride, VEC);
          VL::ld(q + j * jstride, a[j]);
        }
        }
        fold_tile<P, NW, U>(b, t1, run);
        t = t2;
      }
    } else {
      for (; t < nft; t += gridDim.x) {
        const uint8_t* q = lane_base + (uint64_t)t * tile_bytes;
#pragma unroll
        for (int j = 0; j < U; ++j) {
          bd.check(q + j * jstride, VEC);
          VL::ld(q + j * jstride, a[j]);
        }
        fold_tile<P, NW, U>(a, t, run);
      }
    }
  }
  // ------------------------------- the one partial tile (id n_full_tiles + 1)
  if (p.rem_vecs != 0 && blockIdx.x == p.n_full_tiles % gridDim.x && threadIdx.x < p.rem_vecs) {
    uint32_t cur[U][NW];
    const uint8_t* q =
        main_base + (uint64_t)p.n_full_tiles * tile_bytes + (uint64_t)threadIdx.x * VEC;
    bd.check(q, VEC);
    VL::ld(q, cur[0]);
#pragma unroll
    for (int j = 1; j < U; ++j) {
      if ((uint32_t)j * blockDim.x + threadIdx.x < p.rem_vecs) {
        bd.check(q + (uint64_t)j * blockDim.x * VEC, VEC);
        VL::ld(q + (uint64_t)j * blockDim.x * VEC, cur[j]);
      } else {  // duplicates are neutral for a value-only min/max
#pragma unroll
        for (int w = 0; w < NW; ++w) cur[j][w] = cur[0][w];
      }
    }
    fold_tile<P, NW, U>(cur, p.n_full_tiles, run);
  }

  // ------------------------------- CTA reduce on (key, global vector index)
  auto slot_to_vec = [&](uint32_t slot) -> uint64_t {
    return slot == TILE_NONE ? POS_NONE
                             : (uint64_t)(slot / U) * tile_vecs + (uint64_t)(slot % U) * blockDim.x +
                                   threadIdx.x;
  };
  Best<Key> b;
  b.kmin = run.kmin;
  b.kmax = run.kmax;
  b.pmin = slot_to_vec(run.smin);
  b.pmax = slot_to_vec(run.smax);
  b.pnan = slot_to_vec(run.snan);
  b = block_reduce(b, smem);

  // ------------------------------------------- last-block-done grid finalise
  if (!grid_reduce(p, b, smem, &s_is_last)) return;

  // ------- exact finish by ONE warp: the winning vectors (<= 32 elements each) are re-read
  // with one element per lane and a ballot picks the first match; the unaligned head and the
  // sub-vector tail (< one vector each) ride along as exact per-lane candidates.
  __shared__ Best<Key> s_final;
  if (threadIdx.x < 32) {
    using Elem = typename P::Elem;
    const Elem* e = reinterpret_cast<const Elem*>(p.base);
    const uint32_t lane = threadIdx.x;
    Best<Key> x;
    x.clear();
    const uint64_t cand[3] = {b.pmin, b.pmax, b.pnan};
#pragma unroll
    for (int c = 0; c < (P::TRACK_NAN ? 3 : 2); ++c) {
      if (cand[c] == POS_NONE) continue;  // warp-uniform
      const uint64_t first = p.head_elems + cand[c] * VEC_ELEMS;
      bool hit = false;
      if (lane < VEC_ELEMS) {
        Key k;
        bool valid, is_nan;
        bd.check(&e[first + lane], sizeof(Elem));
        P::classify(e[first + lane], k, valid, is_nan);
        hit = (c == 0) ? (valid && k == b.kmin) : (c == 1) ? (valid && k == b.kmax) : is_nan;
      }
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
      if (m != 0 && lane == 0) {
        const uint64_t idx = first + (uint64_t)(__ffs(m) - 1);
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
    // head: elements [0, head_elems); tail: [tail_start, n); both shorter than one vector
#pragma unroll
    for (int part = 0; part < 2; ++part) {
      const uint64_t begin = part == 0 ? 0 : p.tail_start;
      const uint64_t end = part == 0 ? p.head_elems : p.n;
      const uint64_t i = begin + lane;
      if (i < end) {
        Key k;
        bool valid, is_nan;
        bd.check(&e[i], sizeof(Elem));
        P::classify(e[i], k, valid, is_nan);
        if (valid) {
          x.merge_min(k, i);
          x.merge_max(k, i);
        }
        if (P::TRACK_NAN && is_nan && i < x.pnan) x.pnan = i;
      }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) x.merge(shfl_xor(x, m));
    if (lane == 0) s_final = x;
  }
  __syncthreads();
  const Best<Key> x = s_final;

  finalize<XCHG>(p, x);
}

}  // namespace amm
