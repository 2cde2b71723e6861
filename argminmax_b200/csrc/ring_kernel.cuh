// ring_kernel.cuh -- the HBM-sized scan fed through a TMA ring (sm_100a).
//
// Same algorithm as argminmax_scan_kernel (value-only fold per 16-byte vector, slot ids, packed /
// tagged-slot grid step, warp-level exact finish, fused exchange), different data path:
//
//   * one PRODUCER thread per CTA (lane 0 of an extra warp) claims 32 KiB tiles from the
//     workspace's claim counter -- every tile of the array, not just a tail: the claim costs the
//     consumers nothing -- and fetches each with ONE cp.async.bulk (global -> shared, UBLKCP,
//     completion on an mbarrier) into a ring of 3 stages; the tile id travels with the stage;
//   * 256 CONSUMER threads wait for a stage, pull their 128 bytes out of shared memory (8 x 16
//     bytes, lane i reads bytes [16 i, 16 i + 16) of every 4 KiB row; 8-byte dtypes 4 x 32
//     bytes), fold them, and hand the stage back (mbarrier with 256 arrivals).
//
// Bytes in flight are bounded by shared memory (2 CTAs x 96 KiB per SM) instead of registers
// (4 CTAs x 256 threads x 128 B = 128 KiB), the memory system sees 32 KiB sequential requests,
// and the schedule is fully dynamic without a CTA barrier.  Stand-alone harness
// (tools/exp/tma_ring.cu, profiles/r02_tma_ring.md): 7.54 TB/s against 7.42-7.45 for the register
// ping-pong on the same 4 GiB -- but only WITH the producer-side claims (a static ring: 7.0).
// In the product (same file): f32 x 2^30 serialised 7.39 -> 7.46 TB/s, pipelined 7.43 -> 7.50.
// Used for inputs of at least 2 GiB (launch_variant); smaller inputs keep the register path
// (level at 1 GiB, lower fixed cost below).  With 16-byte vectors the 8-byte dtypes lost against
// the register path's 32-byte vectors (f64 -2.9 %): they get 32-byte vectors here too (RingGeom).
#pragma once
#include "scan_kernel.cuh"
#include "windows_kernel.cuh"  // mbarrier / bulk-copy helpers

namespace amm {

constexpr int RING_CONSUMERS = 256;
constexpr int RING_THREADS = RING_CONSUMERS + 32;  // + the producer warp
constexpr int RING_STAGES = 3;
constexpr uint32_t RING_TILE_BYTES = 32768;  // 32 KiB
constexpr uint32_t RING_END = 0xFFFFFFFFu;
// A consumer's vector is 16 bytes (one LDS.128; rows of 4 KiB, conflict-free) for 1/2/4-byte
// elements and 32 bytes (two LDS.128 at a 32-byte lane stride: 2-way bank conflict, far from the
// shared-memory limit) for 8-byte elements, whose per-vector bookkeeping -- 64-bit keys -- is
// then paid per four elements instead of per two.
template <class P>
struct RingGeom {
  static constexpr int VEC = sizeof(typename P::Elem) == 8 ? 32 : 16;
  static constexpr int U = RING_TILE_BYTES / (RING_CONSUMERS * VEC);  // vectors per consumer per tile: 8 / 4
  static constexpr uint32_t TILE_VECS = RING_CONSUMERS * U;           // 2048 / 1024
};

template <class P, bool XCHG, bool PACKED>
__global__ void __launch_bounds__(RING_THREADS) argminmax_ring_kernel(const ScanParams p) {
  using Key = typename P::Key;
  constexpr int VEC = RingGeom<P>::VEC, NW = VEC / 4, U = RingGeom<P>::U;
  constexpr uint32_t RING_TILE_VECS = RingGeom<P>::TILE_VECS;
  constexpr uint32_t VEC_ELEMS = VEC / sizeof(typename P::Elem);
  static_assert(!PACKED || sizeof(Key) == 4, "packed candidates need 32-bit keys");
  extern __shared__ __align__(128) uint8_t ring[];  // RING_STAGES x RING_TILE_BYTES
  __shared__ ReduceScratch<Key> scratch;
  __shared__ __align__(8) uint64_t s_full[RING_STAGES], s_empty[RING_STAGES];
  __shared__ uint32_t s_tile[RING_STAGES];
  pdl_entry(p);
  stamp(p, ST_ENTRY);
  const uint8_t* main_base = p.base + p.head_elems * sizeof(typename P::Elem);
  const Bounds bd(p);
  const uint32_t n_tiles = p.n_full_tiles;  // whole 32 KiB tiles, all claimed at run time
  uint32_t* const claim = p.ticket + WS_CLAIM_RING + (uint32_t)(p.seq & 63ull);  // ring of counters, see scan_kernel.cuh

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < RING_STAGES; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], RING_CONSUMERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  Running<Key> run, run_rem;
  running_clear(run);
  running_clear(run_rem);
  const bool consumer = threadIdx.x < RING_CONSUMERS;

  if (!consumer) {
    // ------------------------------------------------------------------ producer
    if (threadIdx.x == RING_CONSUMERS) {
      for (uint32_t i = 0;; ++i) {
        const uint32_t t = atomicAdd(claim, 1u);
        const uint32_t s = i % RING_STAGES;
        if (i >= RING_STAGES) mbar_wait(&s_empty[s], ((i / RING_STAGES) - 1) & 1);  // stage handed back
        if (t >= n_tiles) {
          s_tile[s] = RING_END;
          mbar_arrive(&s_full[s]);
          break;
        }
        s_tile[s] = t;
        const uint8_t* src = main_base + (uint64_t)t * RING_TILE_BYTES;
        bd.check(src, RING_TILE_BYTES);
        mbar_arrive_expect_tx(&s_full[s], RING_TILE_BYTES);
        tma_bulk_g2s(ring + (size_t)s * RING_TILE_BYTES, src, RING_TILE_BYTES, &s_full[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ consumers
    // the partial tile behind the last whole one (CTA rem_owner, ordinary global loads): its
    // loads go out before the ring is touched; own running state, merged with "main wins ties"
    const uint32_t rem = blockIdx.x == p.rem_owner ? p.rem_vecs : 0;
    const bool mine = threadIdx.x < rem;
    uint32_t cur[U][NW];
    if (mine) {
      const uint8_t* q = main_base + ((uint64_t)n_tiles * RING_TILE_VECS + threadIdx.x) * VEC;
#pragma unroll
      for (int j = 0; j < U; ++j) {
        if ((uint32_t)j * RING_CONSUMERS + threadIdx.x < rem) {
          bd.check(q + (size_t)j * RING_CONSUMERS * VEC, VEC);
          VecLoad<VEC>::ld(q + (size_t)j * RING_CONSUMERS * VEC, cur[j]);
        }
      }
#pragma unroll
      for (int j = 1; j < U; ++j) {
        if (!((uint32_t)j * RING_CONSUMERS + threadIdx.x < rem)) {  // duplicates are neutral for a value-only min/max
#pragma unroll
          for (int w = 0; w < NW; ++w) cur[j][w] = cur[0][w];
        }
      }
      fold_tile<P, NW, U>(cur, 0, run_rem);
    }
    const uint32_t my = smem_u32(ring) + threadIdx.x * VEC;
    for (uint32_t c = 0;; ++c) {
      const uint32_t s = c % RING_STAGES;
      mbar_wait(&s_full[s], (c / RING_STAGES) & 1);
      const uint32_t t = s_tile[s];
      if (t == RING_END) break;
      uint32_t a[U][NW];
      const uint32_t sb = my + s * RING_TILE_BYTES;
#pragma unroll
      for (int j = 0; j < U; ++j)
#pragma unroll
        for (int h = 0; h < NW; h += 4)
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(a[j][h]), "=r"(a[j][h + 1]), "=r"(a[j][h + 2]), "=r"(a[j][h + 3])
                       : "r"(sb + j * (RING_CONSUMERS * VEC) + h * 4));
      mbar_arrive(&s_empty[s]);  // the registers hold the data: the stage may be refilled
      fold_tile<P, NW, U>(a, t, run);
#ifdef AMM_PHASE_STAMPS
      if (c == 0) stamp(p, ST_FIRST_FOLD);
#endif
    }
  }
  stamp(p, ST_STREAM_DONE);

  // ------------------- slot ids -> global vector indices (tile t, row j, consumer tid)
  auto slot_to_vec = [&](uint32_t slot, uint64_t tile0) -> uint64_t {
    return slot == TILE_NONE ? POS_NONE
                             : (tile0 + slot / U) * RING_TILE_VECS + (uint64_t)(slot % U) * RING_CONSUMERS + threadIdx.x;
  };
  Key kmin = run.kmin, kmax = run.kmax;
  uint64_t vmin = slot_to_vec(run.smin, 0), vmax = slot_to_vec(run.smax, 0), vnan = slot_to_vec(run.snan, 0);
  {
    const bool um = run_rem.smin != TILE_NONE && (run.smin == TILE_NONE || run_rem.kmin < run.kmin);
    const bool uM = run_rem.smax != TILE_NONE && (run.smax == TILE_NONE || run_rem.kmax > run.kmax);
    kmin = um ? run_rem.kmin : kmin;
    vmin = um ? slot_to_vec(run_rem.smin, n_tiles) : vmin;
    kmax = uM ? run_rem.kmax : kmax;
    vmax = uM ? slot_to_vec(run_rem.smax, n_tiles) : vmax;
    vnan = run.snan == TILE_NONE ? slot_to_vec(run_rem.snan, n_tiles) : vnan;
  }

  // ------------------- CTA + grid reduce (the producer warp takes part with nothing to offer)
  Best<Key> b;
  if constexpr (PACKED) {
    PackedCand c;
    c.clear();
    if (vmin != POS_NONE) c.set(kmin, (uint32_t)vmin, kmax, (uint32_t)vmax);
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
  if (threadIdx.x == 0) *claim = 0u;
  stamp(p, ST_GRID_FOLDED);
  const Best<Key> x = exact_finish_warp<P, VEC_ELEMS>(p, bd, b);
  stamp(p, ST_EXACT_DONE);
  finalize_warp<XCHG>(p, x);
}

}  // namespace amm
