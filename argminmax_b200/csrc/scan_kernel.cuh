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
//    Per TILE it compares the tile's (min key, max key) with its running pair and
//    remembers the TILE ID of the winner (strict compare, tiles visited in
//    ascending order => earliest tile per thread).  No index lanes, hence no index
//    overflow and none of the reference's restart loop; the loop is branch-free and
//    its cost does not depend on the data order.
//  * CTA reduce (warp shuffles, then shared memory) and the cross-CTA reduce
//    (last-block-done ticket) are lexicographic on (key, tile): the earliest tile
//    that contains the global extreme wins.
//  * The last CTA then runs an EXACT lexicographic (key, index) scan over at most
//    five small element ranges: the winning min tile, the winning max tile, the
//    first tile that holds a NaN (return-NaN mode), and the unaligned head and the
//    sub-vector tail of the array.  That yields first-occurrence indices with the
//    reference's tie-break rule (src/simd/task.rs:273,290) and re-reads at most
//    3 tiles (<= ~200 KB, L2-resident) of an arbitrarily large input.
//  * One 64-byte record (keys, global indices, first NaN, flags) is written to the
//    workspace's device slot and to a pinned host slot: one launch, no memcpy.
#pragma once
#include <stdint.h>

#include "../../include/argminmax_b200.h"
#include "policies.cuh"

namespace amm {

constexpr uint32_t TILE_NONE = 0xFFFFFFFFu;
constexpr uint64_t POS_NONE = 0xFFFFFFFFFFFFFFFFull;
constexpr int MAX_THREADS = 512;

struct ScanParams {
  const uint8_t* base;     // caller's device pointer (element 0)
  uint64_t n;              // elements
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
};

// (key, position) candidates for min, max and first NaN.  Positions are tile ids in
// the streaming phase and element indices in the exact phase.
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
template <int VEC>
struct VecLoad;
template <>
struct VecLoad<16> {
  static constexpr int NW = 4;
  static __device__ __forceinline__ void ld(const uint8_t* p, uint32_t (&w)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                 : "l"(p));
  }
};
template <>
struct VecLoad<32> {
  static constexpr int NW = 8;
  // 256-bit load: new on sm_100 (SASS LDG.E.NA.ENL2.LTC256B.256.CONSTANT)
  static __device__ __forceinline__ void ld(const uint8_t* p, uint32_t (&w)[8]) {
    asm volatile(
        "ld.global.nc.L1::no_allocate.L2::256B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]),
          "=r"(w[7])
        : "l"(p));
  }
};

// Running per-thread state of the streaming phase (tile ids are 32-bit here).
template <typename Key>
struct Running {
  Key kmin, kmax;
  uint32_t tmin, tmax, tnan;
};

template <class P, int NW, int U>
__device__ __forceinline__ void fold_tile(const uint32_t (&v)[U][NW], uint32_t tile_id,
                                          Running<typename P::Key>& r) {
  using Key = typename P::Key;
  typename P::Acc acc;
  P::init(acc, &v[0][0]);
#pragma unroll
  for (int j = 0; j < U; ++j) {
#pragma unroll
    for (int w = 0; w < NW; w += P::WPI) {
      if (j == 0 && w == 0) continue;
      P::feed(acc, &v[j][w]);
    }
  }
  Key kmin, kmax;
  bool valid, has_nan;
  P::finish(acc, kmin, kmax, valid, has_nan);
  // tiles arrive in ascending order per thread, so "strictly better, or nothing yet"
  // IS the lexicographic (key, tile) update.
  const bool first = (r.tmin == TILE_NONE);
  const bool um = valid && (first || kmin < r.kmin);
  const bool uM = valid && (first || kmax > r.kmax);
  r.kmin = um ? kmin : r.kmin;
  r.tmin = um ? tile_id : r.tmin;
  r.kmax = uM ? kmax : r.kmax;
  r.tmax = uM ? tile_id : r.tmax;
  r.tnan = (has_nan && r.tnan == TILE_NONE) ? tile_id : r.tnan;
}

// Exact lexicographic (key, index) scan of elements [begin, end).
template <class P>
__device__ __forceinline__ void exact_range(const uint8_t* base, uint64_t begin, uint64_t end,
                                            Best<typename P::Key>& b) {
  using Elem = typename P::Elem;
  using Key = typename P::Key;
  const Elem* e = reinterpret_cast<const Elem*>(base);
  for (uint64_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
    Key k;
    bool valid, is_nan;
    P::classify(e[i], k, valid, is_nan);
    if (valid) {
      b.merge_min(k, i);
      b.merge_max(k, i);
    }
    if (P::TRACK_NAN && is_nan && i < b.pnan) b.pnan = i;
  }
}

template <class P, int VEC, int U, bool PREFETCH>
__global__ void __launch_bounds__(MAX_THREADS) argminmax_scan_kernel(const ScanParams p) {
  using Key = typename P::Key;
  constexpr int NW = VecLoad<VEC>::NW;
  constexpr uint32_t VEC_ELEMS = VEC / sizeof(typename P::Elem);
  __shared__ Best<Key> smem[32];
  __shared__ uint32_t s_is_last;

  const uint32_t tile_vecs = blockDim.x * U;
  const uint64_t tile_bytes = (uint64_t)tile_vecs * VEC;
  const uint8_t* main_base = p.base + p.head_elems * sizeof(typename P::Elem);

  Running<Key> run;
  run.kmin = KeyLim<Key>::MAXV;
  run.kmax = KeyLim<Key>::MINV;
  run.tmin = run.tmax = run.tnan = TILE_NONE;

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
        for (int j = 0; j < U; ++j) VecLoad<VEC>::ld(q + j * jstride, a[j]);
      }
      while (t < nft) {
        const uint32_t t1 = t + gridDim.x;
        if (t1 < nft) {
          const uint8_t* q = lane_base + (uint64_t)t1 * tile_bytes;
#pragma unroll
          for (int j = 0; j < U; ++j) VecLoad<VEC>::ld(q + j * jstride, b[j]);
        }
        fold_tile<P, NW, U>(a, t + 1, run);
        if (t1 >= nft) break;
        const uint32_t t2 = t1 + gridDim.x;
        if (t2 < nft) {
          const uint8_t* q = lane_base + (uint64_t)t2 * tile_bytes;
#pragma unroll
          for (int j = 0; j < U; ++j) VecLoad<VEC>::ld(q + j * jstride, a[j]);
        }
        fold_tile<P, NW, U>(b, t1 + 1, run);
        t = t2;
      }
    } else {
      for (; t < nft; t += gridDim.x) {
        const uint8_t* q = lane_base + (uint64_t)t * tile_bytes;
#pragma unroll
        for (int j = 0; j < U; ++j) VecLoad<VEC>::ld(q + j * jstride, a[j]);
        fold_tile<P, NW, U>(a, t + 1, run);
      }
    }
  }
  // ------------------------------- the one partial tile (id n_full_tiles + 1)
  if (p.rem_vecs != 0 && blockIdx.x == p.n_full_tiles % gridDim.x && threadIdx.x < p.rem_vecs) {
    uint32_t cur[U][NW];
    const uint8_t* q =
        main_base + (uint64_t)p.n_full_tiles * tile_bytes + (uint64_t)threadIdx.x * VEC;
    VecLoad<VEC>::ld(q, cur[0]);
#pragma unroll
    for (int j = 1; j < U; ++j) {
      if ((uint32_t)j * blockDim.x + threadIdx.x < p.rem_vecs) {
        VecLoad<VEC>::ld(q + (uint64_t)j * blockDim.x * VEC, cur[j]);
      } else {  // duplicates are neutral for a value-only min/max
#pragma unroll
        for (int w = 0; w < NW; ++w) cur[j][w] = cur[0][w];
      }
    }
    fold_tile<P, NW, U>(cur, p.n_full_tiles + 1, run);
  }

  // ------------------------------------------- CTA reduce on (key, tile id)
  Best<Key> b;
  b.kmin = run.kmin;
  b.kmax = run.kmax;
  b.pmin = run.tmin == TILE_NONE ? POS_NONE : (uint64_t)run.tmin;
  b.pmax = run.tmax == TILE_NONE ? POS_NONE : (uint64_t)run.tmax;
  b.pnan = run.tnan == TILE_NONE ? POS_NONE : (uint64_t)run.tnan;
  b = block_reduce(b, smem);

  // ------------------------------------------- last-block-done grid finalise
  Best<Key>* partials = reinterpret_cast<Best<Key>*>(p.partials);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = b;
    __threadfence();
    const uint32_t ticket = atomicInc(p.ticket, gridDim.x - 1);  // wraps to 0: reusable
    s_is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_is_last) return;
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
  g = block_reduce(g, smem);

  // ------------------- exact (key, index) scan of the few candidate ranges
  const uint64_t tile_elems = (uint64_t)tile_vecs * VEC_ELEMS;
  Best<Key> x;
  x.clear();
  exact_range<P>(p.base, 0, p.head_elems, x);
  exact_range<P>(p.base, p.tail_start, p.n, x);
  const uint64_t cand[3] = {g.pmin, g.pmax, g.pnan};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const uint64_t tid = cand[c];
    if (tid == POS_NONE) continue;
    if (c == 1 && tid == cand[0]) continue;
    if (c == 2 && (tid == cand[0] || tid == cand[1])) continue;
    const uint64_t begin = p.head_elems + (tid - 1) * tile_elems;
    uint64_t end = begin + tile_elems;
    end = end < p.tail_start ? end : p.tail_start;
    exact_range<P>(p.base, begin, end, x);
  }
  x = block_reduce(x, smem);

  if (threadIdx.x == 0) {
    amm_record r;
    const bool has_value = x.pmin != POS_NONE;
    const bool has_nan = x.pnan != POS_NONE;
    r.min_key = has_value ? (int64_t)x.kmin : 0;
    r.max_key = has_value ? (int64_t)x.kmax : 0;
    r.min_idx = has_value ? x.pmin + p.global_offset : AMM_NO_INDEX;
    r.max_idx = has_value ? x.pmax + p.global_offset : AMM_NO_INDEX;
    r.nan_idx = has_nan ? x.pnan + p.global_offset : AMM_NO_INDEX;
    r.flags = (has_value ? AMM_REC_HAS_VALUE : 0ull) | (has_nan ? AMM_REC_HAS_NAN : 0ull);
    r.n = p.n;
    r.seq = p.seq;
    *p.rec_dev = r;
    if (p.rec_host != nullptr) {
      volatile amm_record* h = p.rec_host;
      h->min_key = r.min_key;
      h->min_idx = r.min_idx;
      h->max_key = r.max_key;
      h->max_idx = r.max_idx;
      h->nan_idx = r.nan_idx;
      h->flags = r.flags;
      h->n = r.n;
      __threadfence_system();
      h->seq = r.seq;
    }
  }
}

}  // namespace amm
