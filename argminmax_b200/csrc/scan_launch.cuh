// scan_launch.cuh -- host-side launch geometry for the scan kernels (included by the
// per-policy-group translation units, scan_inst.cu).
#pragma once
#include <stdlib.h>
#include <string.h>

#include "amm_internal.h"
#include "scan_kernel.cuh"
#include "windows_kernel.cuh"
#include "ring_kernel.cuh"

namespace amm {

// Load shapes compiled for every policy.  {VEC bytes, U vectors per thread per tile,
// software prefetch}.  Tile = threads * U * VEC bytes.
struct VariantDesc {
  int vec, u, prefetch, hint;
  const char* name;
};
static const VariantDesc kVariants[] __attribute__((unused)) = {
    {16, 4, 1, 2, "v16x4_pf"},  // 0: 128-bit loads, 64 B/thread/tile, 2 tiles in flight (auto, < 512 MiB)
    {32, 2, 1, 2, "v32x2_pf"},  // 1: 256-bit loads, 64 B/thread/tile, 2 tiles in flight (auto, >= 512 MiB)
    {32, 4, 1, 2, "v32x4_pf"},  // 2: 256-bit loads, 128 B/thread/tile, 2 tiles in flight
    {16, 8, 0, 2, "v16x8"},     // 3: 128-bit loads, 128 B/thread/tile, no software prefetch
    {16, 2, 1, 2, "v16x2_pf"},  // 4: small tiles
    {16, 4, 1, 0, "v16x4_pf_nohint"},  // 5: as 0 without the .L2::256B prefetch-size hint
    {16, 8, 0, 0, "tma_ring_32k_x3"},  // 6: ring_kernel.cuh: 32 KiB tiles by TMA bulk copy, 3 stages, all tiles claimed
};
constexpr int kNumVariants = (int)(sizeof(kVariants) / sizeof(kVariants[0]));

// Launch with the programmatic-stream-serialization attribute (see pdl_* in scan_kernel.cuh).
template <typename Kernel>
static cudaError_t launch_scan(Kernel kernel, unsigned grid, unsigned threads, cudaStream_t stream,
                               const ScanParams& p, bool pdl, size_t dyn_smem = 0) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = dyn_smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, p);
}

// ---------------------------------------------------------------------------------------
// Tile plan of the streaming kernel: pure host arithmetic (unit-tested without a GPU through
// amm_debug_plan_scan, tests/test_scan_plan.py, which replays the kernel's index formulas and
// checks that every vector is visited exactly once).
// ---------------------------------------------------------------------------------------
struct ScanPlan {
  uint64_t head_elems, n_vec, tail_start;  // unaligned head | whole vectors | sub-vector tail
  uint64_t per_vecs;                       // CONTIGUOUS: vectors per CTA (multiple of threads), else 0
  uint64_t grid;
  uint32_t n_full_tiles;  // statically assigned full tiles
  uint32_t rem_vecs;      // vectors of the one partial tile (behind all full tiles)
  uint32_t rem_owner;     // GRID-STRIDE / DYNAMIC: the CTA that takes it
  uint32_t dyn_tiles;     // DYNAMIC: full tiles claimed at run time (behind the static ones)
  uint32_t dyn_koff;      // slot tile ids >= dyn_koff name claimed tiles (UINT32_MAX: none)
  int kind;               // 0 grid-stride, 1 contiguous, 2 dynamic tail
  int status;             // AMM_OK or AMM_ERR_TOO_LARGE
};

struct ScanPlanKnobs {
  int sm_count;
  int scan_contig;         // -1 automatic, 0 never, 1 always
  int scan_dyn_pct;        // -1 automatic, 0 none, 1..90
  int scan_dyn_pipelined;  // claimed tail also for pipelined scans (HBM-sized inputs only)
};

static ScanPlan plan_scan(uint64_t addr, uint64_t n, uint32_t es, uint32_t vec, uint32_t u, uint32_t threads,
                          uint64_t grid_cap, bool pipelined, const ScanPlanKnobs& k) {
  ScanPlan pl;
  memset(&pl, 0, sizeof(pl));
  pl.dyn_koff = 0xFFFFFFFFu;
  const uint64_t vec_elems = vec / es;
  const uint64_t tile_vecs = (uint64_t)threads * u;
  uint64_t head = ((vec - (addr % vec)) % vec) / es;
  if (head > n) head = n;
  pl.head_elems = head;
  pl.n_vec = (n - head) / vec_elems;
  pl.tail_start = head + pl.n_vec * vec_elems;
  const uint64_t full = pl.n_vec / tile_vecs;
  if ((2 * full + 4) * (uint64_t)u >= 0xFFFFFFFFull) {  // 32-bit slot ids
    pl.status = AMM_ERR_TOO_LARGE;
    return pl;
  }
  pl.n_full_tiles = (uint32_t)full;
  pl.rem_vecs = (uint32_t)(pl.n_vec % tile_vecs);
  const uint64_t tiles = full + (pl.rem_vecs ? 1 : 0);
  uint64_t grid = grid_cap < 1 ? 1 : grid_cap;
  // How the tiles are handed out (profiles/r02_fixed_cost.md).  CTAs given equal shares finish
  // far apart -- the memory system serves them at very different rates (2.4x between the first
  // and the last CTA at 40 MB, +-2.6 % at 4 GiB) and the stragglers then stream alone, latency-
  // bound.  So, with at least two tiles per CTA:
  //   DYNAMIC TAIL  the first tiles grid-stride, the last AMM_SCAN_DYN_PCT % claimed at run time
  //                 (8 % on HBM-sized inputs -- more costs a barrier per tile for nothing --
  //                 30 % below 0.5 GiB);
  // otherwise
  //   CONTIGUOUS    fewer than 64 tiles per CTA: equal ranges of whole rows (blockDim vectors);
  //   GRID-STRIDE   else: whole tiles b, b + G, ... (the +-1 tile imbalance is noise there).
  const bool big = n * es >= ((uint64_t)512 << 20);
  const int dyn_pct = k.scan_dyn_pct >= 0 ? k.scan_dyn_pct : (big ? 8 : 30);
  // (pipelined scans: only on HBM-sized inputs, +0.9 % at 4 GiB; below that the next scan's
  // streaming already fills the gaps the stragglers leave and the claims only cost: 100 M u8
  // 14.1 -> 15.4 us per call in a stream of calls)
  const bool dynamic = dyn_pct > 0 && (!pipelined || (k.scan_dyn_pipelined && big)) && k.scan_contig != 1 &&
                       full >= 2 * grid && grid >= (uint64_t)k.sm_count;
  const bool contig = !dynamic && (k.scan_contig < 0 ? tiles < 64 * grid : k.scan_contig != 0);
  if (dynamic) {
    uint64_t dyn = full * (uint64_t)(dyn_pct > 90 ? 90 : dyn_pct) / 100;
    if (dyn < grid) dyn = grid;  // one claim per CTA goes out on entry
    pl.dyn_tiles = (uint32_t)dyn;
    pl.n_full_tiles = (uint32_t)(full - dyn);
    pl.dyn_koff = (uint32_t)((full - dyn + grid - 1) / grid + 1);
    pl.kind = 2;
  } else if (contig) {
    const uint64_t rows = (pl.n_vec + (uint64_t)threads - 1) / (uint64_t)threads;
    uint64_t rows_per_cta = (rows + grid - 1) / grid;
    if (rows_per_cta < 1) rows_per_cta = 1;
    grid = (rows + rows_per_cta - 1) / rows_per_cta;  // drops the CTAs that would get nothing
    pl.per_vecs = rows_per_cta * (uint64_t)threads;
    pl.kind = 1;
  } else {
    if (grid > tiles) grid = tiles;
    pl.rem_owner = grid ? (uint32_t)(full % grid) : 0;
    pl.kind = 0;
  }
  pl.grid = grid < 1 ? 1 : grid;
  return pl;
}

template <class P, int VEC, int U, bool PF, int HINT, bool XCHG, bool PACKED>
static int launch_packed(amm_workspace* ws, ScanParams& p, int dtype_slot, int variant,
                         cudaStream_t stream) {
  auto kernel = argminmax_scan_kernel<P, VEC, U, PF, HINT, XCHG, PACKED>;
  const int threads = ws->eff_threads;
  // fused / packed instantiations cache their occupancy separately
  const int vslot = variant + (XCHG ? kNumVariants : 0) + (PACKED ? 2 * kNumVariants : 0);
  static_assert(4 * kNumVariants <= AMM_MAX_VARIANTS, "occupancy cache too small");
  int& occ = ws->occupancy[dtype_slot][vslot];
  if (occ == 0 || ws->occupancy_threads[dtype_slot][vslot] != threads) {
    int o = 0;
    AMM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kernel, threads, 0));
    occ = o < 1 ? 1 : o;
    ws->occupancy_threads[dtype_slot][vslot] = threads;
  }
  int per_sm = ws->eff_ctas_per_sm < occ ? ws->eff_ctas_per_sm : occ;
  uint64_t grid_cap = (uint64_t)ws->sm_count * (uint64_t)per_sm;
  if (grid_cap > (uint64_t)ws->max_grid) grid_cap = (uint64_t)ws->max_grid;
  const ScanPlanKnobs knobs = {ws->sm_count, ws->scan_contig, ws->scan_dyn_pct, ws->scan_dyn_pipelined};
  const ScanPlan pl = plan_scan((uint64_t)(uintptr_t)p.base, p.n, (uint32_t)sizeof(typename P::Elem), VEC, U,
                                (uint32_t)threads, grid_cap, p.pipelined != 0, knobs);
  if (pl.status) return pl.status;
  p.head_elems = pl.head_elems;
  p.n_vec = pl.n_vec;
  p.tail_start = pl.tail_start;
  p.n_full_tiles = pl.n_full_tiles;
  p.rem_vecs = pl.rem_vecs;
  p.rem_owner = pl.rem_owner;
  p.per_vecs = pl.per_vecs;
  p.dyn_tiles = pl.dyn_tiles;
  p.dyn_koff = pl.dyn_koff;
  const uint64_t grid = pl.grid;
  // With programmatic dependent launch the CTAs of the NEXT scan on the stream become resident
  // while this one runs.  If an SM can hold more than twice the CTAs a scan puts on it, the
  // block scheduler packs the next grid onto the free slots of fewer SMs (3 per SM instead of
  // 2) and that scan then streams from part of the GPU only (measured with a 47-register
  // build of this kernel and 2 CTAs per SM: +10-20 us on a 200 MB scan).  Unused dynamic shared
  // memory caps the co-residency at exactly 2 x per_sm CTAs per SM.  Not in pipelined mode,
  // where consecutive scans stream concurrently and extra resident CTAs help.
  size_t dyn = 0;
  if (ws->pdl && !p.pipelined && occ > 2 * per_sm && ws->smem_per_sm > 0) {
    size_t& stat = ws->static_smem[dtype_slot][vslot];
    if (stat == 0) {
      cudaFuncAttributes fa;
      AMM_CUDA(cudaFuncGetAttributes(&fa, kernel));
      stat = fa.sharedSizeBytes + 1;  // +1: "queried" marker for kernels without static smem
    }
    const size_t per_cta = ws->smem_per_sm / (size_t)(2 * per_sm);
    const size_t fixed = (stat - 1) + ws->smem_reserved_per_cta;
    if (per_cta > fixed + 1024) {
      dyn = (per_cta - fixed) / 128 * 128;
      if (dyn > ws->smem_optin_per_cta - (stat - 1)) dyn = (ws->smem_optin_per_cta - (stat - 1)) / 128 * 128;
      size_t& have = ws->dyn_smem_set[dtype_slot][vslot];
      if (have < dyn) {
        AMM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        have = dyn;
      }
    }
  }
  AMM_CUDA(launch_scan(kernel, (unsigned)grid, (unsigned)threads, stream, p, ws->pdl != 0, dyn));
  ws->last_grid = (int)grid;
  ws->last_threads = threads;
  ws->launches += 1;
  return AMM_OK;
}

// 32-bit keys with fewer than 2^32 - 2 vectors: the cross-CTA step travels as packed 64-bit
// words (see packed_grid_reduce); everything else goes through the tagged slots (grid_reduce).
template <class P, int VEC, int U, bool PF, int HINT, bool XCHG>
static int launch_one(amm_workspace* ws, ScanParams& p, int dtype_slot, int variant,
                      cudaStream_t stream) {
  if constexpr (sizeof(typename P::Key) == 4) {
    const uint64_t vec_elems = VEC / sizeof(typename P::Elem);
    if (ws->scan_packed && p.n / vec_elems < 0xFFFFFFFEull)
      return launch_packed<P, VEC, U, PF, HINT, XCHG, true>(ws, p, dtype_slot, variant, stream);
  }
  return launch_packed<P, VEC, U, PF, HINT, XCHG, false>(ws, p, dtype_slot, variant, stream);
}

// HBM-sized inputs: the TMA-ring kernel (ring_kernel.cuh), every tile claimed at run time.
template <class P, bool XCHG, bool PACKED>
static int launch_ring_packed(amm_workspace* ws, ScanParams& p, int dtype_slot, cudaStream_t stream) {
  auto kernel = argminmax_ring_kernel<P, XCHG, PACKED>;
  constexpr uint64_t VEC = RingGeom<P>::VEC;
  constexpr uint64_t TILE_VECS = RingGeom<P>::TILE_VECS;
  const uint64_t vec_elems = VEC / sizeof(typename P::Elem);
  const uint64_t addr = (uint64_t)(uintptr_t)p.base;
  uint64_t head = ((VEC - (addr % VEC)) % VEC) / sizeof(typename P::Elem);
  if (head > p.n) head = p.n;
  p.head_elems = head;
  p.n_vec = (p.n - head) / vec_elems;
  p.tail_start = head + p.n_vec * vec_elems;
  const uint64_t full = p.n_vec / TILE_VECS;
  if ((full + 2) * (uint64_t)RingGeom<P>::U >= 0xFFFFFFFFull) return AMM_ERR_TOO_LARGE;  // 32-bit slot ids
  p.n_full_tiles = (uint32_t)full;
  p.rem_vecs = (uint32_t)(p.n_vec % TILE_VECS);
  p.rem_owner = 0;
  p.per_vecs = 0;
  p.dyn_tiles = 0;
  p.dyn_koff = 0xFFFFFFFFu;
  const size_t dyn = (size_t)RING_STAGES * RING_TILE_BYTES;
  int& ready = ws->ring_ready[dtype_slot][(XCHG ? 1 : 0) + (PACKED ? 2 : 0)];
  if (!ready) {
    AMM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    ready = 1;
  }
  uint64_t grid = (uint64_t)ws->sm_count * (uint64_t)ws->ring_ctas_per_sm;
  if (grid > (uint64_t)ws->max_grid) grid = (uint64_t)ws->max_grid;
  if (grid > full + 1) grid = full + 1;
  if (grid < 1) grid = 1;
  // the claim-counter ring is safe for overlapping scans only when a grid has at least one CTA
  // per SM (scan_kernel.cuh); smaller grids (the ring forced onto a small input) run serialised
  if (grid < (uint64_t)ws->sm_count) p.pipelined = 0;
  AMM_CUDA(launch_scan(kernel, (unsigned)grid, (unsigned)RING_THREADS, stream, p, ws->pdl != 0, dyn));
  ws->last_grid = (int)grid;
  ws->last_threads = RING_THREADS;
  ws->eff_threads = RING_THREADS;
  ws->eff_ctas_per_sm = ws->ring_ctas_per_sm;
  ws->eff_variant = kNumVariants - 1;
  ws->launches += 1;
  return AMM_OK;
}

template <class P, bool XCHG>
static int launch_ring(amm_workspace* ws, ScanParams& p, int dtype_slot, cudaStream_t stream) {
  if constexpr (sizeof(typename P::Key) == 4) {
    const uint64_t vec_elems = RingGeom<P>::VEC / sizeof(typename P::Elem);
    if (ws->scan_packed && p.n / vec_elems < 0xFFFFFFFEull)
      return launch_ring_packed<P, XCHG, true>(ws, p, dtype_slot, stream);
  }
  return launch_ring_packed<P, XCHG, false>(ws, p, dtype_slot, stream);
}

// Latency path for small inputs: one trip, see argminmax_direct_kernel.
// Shape of the launch, pure host logic (tests/test_direct_plan.py replays it through
// amm_debug_plan_direct).  Latency, not throughput: as long as CTAs are available every thread
// gets ONE vector (more CTAs are free latency); beyond sm_count x ctas_per_sm CTAs threads take
// 2..4 vectors, all in flight at once.  Small inputs take ONE CTA -- `threads` (256) threads up to
// 4 x threads vectors, 512 threads up to one_cta_vecs (2048 vectors = 32 KiB) and 16 Ki elements:
// a second CTA brings in the cross-CTA round trip (~1 us; in-process A/B,
// profiles/r02g_latency_ab_inproc.jsonl: 8 Ki f32 9.65 -> 8.6 us, 16 Ki f16 9.85 -> 8.9, 4 Ki i64
// 10.5 -> 8.4).  Few deep threads beat many shallow ones (4 Ki f32: 8.6 us as 256 x 4, 9.4 as
// 1024 x 1), but one SM folding 64 KiB is no faster than 16 SMs plus the round trip (u8 x 64 Ki in
// one CTA: 11.1 us against 9.3).
struct DirectPlanKnobs {
  int sm_count, ctas_per_sm, threads, min_vpt, one_cta_vecs;
};
struct DirectPlan {
  uint64_t head_elems, n_vec, tail_start, grid;
  uint32_t threads, vpt;
  bool fits;  // false: more than 4 vectors per thread at the CTA cap -> streaming kernel
};
static DirectPlan plan_direct(uint64_t addr, uint64_t n, uint32_t es, const DirectPlanKnobs& k) {
  constexpr uint64_t VEC = 16;
  DirectPlan pl;
  const uint64_t vec_elems = VEC / es;
  uint64_t head = ((VEC - (addr % VEC)) % VEC) / es;
  if (head > n) head = n;
  pl.head_elems = head;
  pl.n_vec = (n - head) / vec_elems;
  pl.tail_start = head + pl.n_vec * vec_elems;
  uint64_t threads = (uint64_t)k.threads;
  const uint64_t cap = (uint64_t)k.sm_count * (uint64_t)k.ctas_per_sm;
  pl.fits = n / vec_elems <= cap * threads * 4;
  uint64_t vpt = (pl.n_vec + cap * threads - 1) / (cap * threads);
  if (vpt < (uint64_t)k.min_vpt) vpt = (uint64_t)k.min_vpt;
  if (pl.n_vec > threads * 4 && pl.n_vec <= (uint64_t)k.one_cta_vecs &&
      pl.n_vec <= (uint64_t)DIRECT_MAX_THREADS * 4 && n <= 16384)
    threads = DIRECT_MAX_THREADS;
  if (pl.n_vec <= threads * 4) vpt = (pl.n_vec + threads - 1) / threads;
  if (vpt < 1) vpt = 1;
  if (vpt > 4) pl.fits = false;
  pl.threads = (uint32_t)threads;
  pl.vpt = (uint32_t)vpt;
  pl.grid = (pl.n_vec + threads * vpt - 1) / (threads * vpt);
  if (pl.grid < 1) pl.grid = 1;
  return pl;
}
static DirectPlanKnobs direct_knobs(const amm_workspace* ws) {
  return {ws->sm_count, ws->direct_ctas_per_sm, ws->direct_threads, ws->direct_vecs_per_thread, ws->direct_one_cta_vecs};
}

template <class P>
static bool direct_fits(const amm_workspace* ws, uint64_t n) {
  return plan_direct(0, n, (uint32_t)sizeof(typename P::Elem), direct_knobs(ws)).fits;
}

template <class P>
static int launch_direct(amm_workspace* ws, ScanParams& p, cudaStream_t stream) {
  const DirectPlan pl = plan_direct((uint64_t)(uintptr_t)p.base, p.n, (uint32_t)sizeof(typename P::Elem), direct_knobs(ws));
  if (!pl.fits) return AMM_ERR_ARG;  // launch_variant checks direct_fits() first
  p.head_elems = pl.head_elems;
  p.n_vec = pl.n_vec;
  p.tail_start = pl.tail_start;
  p.direct_vpt = pl.vpt;
  const uint64_t grid = pl.grid, threads = pl.threads;
  const bool packed = sizeof(typename P::Key) == 4 && p.n < 0xFFFFFFFEull && ws->direct_packed;
  if (p.world > 1) {
    if constexpr (sizeof(typename P::Key) == 4) {
      if (packed) {
        AMM_CUDA(launch_scan(argminmax_direct_kernel<P, true, true>, (unsigned)grid, (unsigned)threads,
                             stream, p, ws->pdl != 0));
      }
    }
    if (!packed)
      AMM_CUDA(launch_scan(argminmax_direct_kernel<P, true>, (unsigned)grid, (unsigned)threads, stream,
                           p, ws->pdl != 0));
  } else if (packed) {
    // 32-bit keys: candidates travel as packed words (see packed_grid_reduce)
    if constexpr (sizeof(typename P::Key) == 4)
      AMM_CUDA(launch_scan(argminmax_direct_kernel<P, false, true>, (unsigned)grid,
                           (unsigned)threads, stream, p, ws->pdl != 0));
  } else {
    AMM_CUDA(launch_scan(argminmax_direct_kernel<P, false>, (unsigned)grid, (unsigned)threads,
                         stream, p, ws->pdl != 0));
  }
  ws->last_grid = (int)grid;
  ws->last_threads = (int)threads;
  ws->launches += 1;
  return AMM_OK;
}

template <class P>
static int launch_variant(amm_workspace* ws, ScanParams& p, int dtype_slot, cudaStream_t stream) {
  const uint64_t bytes = p.n * sizeof(typename P::Elem);
  // Latency kernel below the measured crossover with the streaming kernel (tools/latency_c.py),
  // per key width, and only while the input fits its single trip; an explicit threshold
  // (amm_workspace_set_direct_threshold / AMM_DIRECT_BYTES) replaces the size rule.
  const bool direct = (ws->direct_auto
                           ? p.n <= (sizeof(typename P::Key) == 4 ? ws->direct_elems32 : ws->direct_elems64)
                           : bytes <= ws->direct_bytes) &&
                      direct_fits<P>(ws, p.n);
  if (direct) return launch_direct<P>(ws, p, stream);
  // Automatic geometry (B200 measurements, profiles/r02_fixed_cost.md): 256-bit loads, 256-thread
  // CTAs (16 KiB tiles), as many CTAs per SM as fit (4-5) on HBM-sized inputs and 4 below 0.5 GiB
  // (128 KiB in flight per SM either way; with 2 CTAs per SM a 400 MB scan streams at 5.7 TB/s
  // instead of 6.9).
  const bool big = bytes >= ((uint64_t)512 << 20);
  // Inputs of 2 GiB and more take the TMA-ring kernel: +0.8-1.2 % at 4 GiB (f32 7.39 -> 7.46 TB/s
  // serialised, 7.43 -> 7.50 pipelined), +2 % at 64 GiB, level at 1 GiB.
  // AMM_SCAN_RING=0: never, 1: for every streaming scan.
  const bool ring = ws->variant == kNumVariants - 1 ||
                    (ws->variant < 0 && (ws->scan_ring < 0 ? bytes >= ((uint64_t)2 << 30) : ws->scan_ring != 0));
  if (ring) return p.world > 1 ? launch_ring<P, true>(ws, p, dtype_slot, stream)
                               : launch_ring<P, false>(ws, p, dtype_slot, stream);
  // Pipelined scans below 128 MiB keep round 1's shape (128-bit loads, 2 CTAs per SM): the point of
  // pipelining is that the NEXT scan's CTAs are resident and streaming while this one's tail runs,
  // and 4 CTAs x 256 threads x 58-64 registers fill the SM (10 M f32 in a stream of calls: 6.6 us
  // with 2 CTAs per SM, 7.3-7.5 with 4).
  const bool small_pipelined = p.pipelined && bytes < ((uint64_t)128 << 20);
  ws->eff_variant = ws->variant >= 0 ? ws->variant : (small_pipelined ? 0 : 1);
  ws->eff_threads = ws->threads > 0 ? ws->threads : 256;
  ws->eff_ctas_per_sm = ws->ctas_per_sm > 0 ? ws->ctas_per_sm : (big ? 8 : (small_pipelined ? 2 : 4));
  if (p.world > 1) {
    // fused scan + peer exchange: instantiated for the two automatic shapes only
    if (ws->eff_variant != 0) ws->eff_variant = 1;
    return ws->eff_variant == 0 ? launch_one<P, 16, 4, true, 2, true>(ws, p, dtype_slot, 0, stream)
                                : launch_one<P, 32, 2, true, 2, true>(ws, p, dtype_slot, 1, stream);
  }
  switch (ws->eff_variant) {
    case 0: return launch_one<P, 16, 4, true, 2, false>(ws, p, dtype_slot, 0, stream);
    case 1: return launch_one<P, 32, 2, true, 2, false>(ws, p, dtype_slot, 1, stream);
    case 2: return launch_one<P, 32, 4, true, 2, false>(ws, p, dtype_slot, 2, stream);
    case 3: return launch_one<P, 16, 8, false, 2, false>(ws, p, dtype_slot, 3, stream);
    case 4: return launch_one<P, 16, 2, true, 2, false>(ws, p, dtype_slot, 4, stream);
    case 5: return launch_one<P, 16, 4, true, 0, false>(ws, p, dtype_slot, 5, stream);
    default: return AMM_ERR_ARG;
  }
}

// Windowed form.  Regimes (B200 measurements, tools/windows_bench.py):
//   <= 1 KB per window            TMA-staged blocks of windows, 1/2/4/8 lanes per window
//   1 - 2 KB per window           8 lanes per window (4 windows per warp), loads from global
//   2 - 4 KB per window           TMA-staged, one warp per window
//   >= 4096 windows               one warp per window (enough warps to fill the machine)
//   >= 4 x SMs windows            one 256-thread CTA per window, 16 KB in flight per CTA
//   fewer                         one 512-thread CTA per window, 64 KB in flight per CTA
//                                 (or one streaming scan per window, see windows_huge)
template <class P, int L>
static int launch_windows_staged(amm_workspace* ws, const WindowParams& p, uint64_t mean_bytes,
                                 cudaStream_t stream) {
  constexpr int THREADS = 128;
  constexpr uint64_t WPB = THREADS / L;
  constexpr uint64_t ES = sizeof(typename P::Elem);
  auto grid_for = [&](uint64_t n_windows) {
    const uint64_t n_blocks = (n_windows + WPB - 1) / WPB;
    uint64_t grid = (uint64_t)ws->sm_count * 6;
    grid = n_blocks < grid ? n_blocks : grid;
    return (unsigned)(grid < 1 ? 1 : grid);
  };
  WindowParams rest = p;  // what the general kernel still has to do
  // Fixed windows of a multiple of 16 bytes on a 16-byte aligned array: all FULL windows take
  // the geometry-free instantiation; a last, shorter window (if any) goes to the general one.
  if (p.offsets == nullptr && p.w_first == 0 && (p.window * ES) % 16 == 0 &&
      ((uintptr_t)p.base % 16) == 0 && ws->win_fast) {
    const uint64_t n_full = p.n / p.window;
    if (n_full > 0) {
      WindowParams pf = p;
      pf.n_windows = n_full;
      // tiny windows: several passes per block so that a block (one bulk copy, one barrier) is
      // ~8 KiB or more; <= 128 lanes x 128 B = 16 KiB per stage by the dispatch rule
      const uint64_t pass_bytes = WPB * p.window * ES;
      uint64_t kwin = pass_bytes >= 8192 ? 1 : 8192 / pass_bytes;
      if (kwin > 8) kwin = 8;
      const uint64_t stage = pass_bytes * kwin;
      const uint64_t n_blocks = (n_full + WPB * kwin - 1) / (WPB * kwin);
      uint64_t grid = (uint64_t)ws->sm_count * 6;
      grid = n_blocks < grid ? n_blocks : grid;
      argminmax_windows_staged_kernel<P, L, THREADS, true>
          <<<(unsigned)(grid < 1 ? 1 : grid), THREADS, (size_t)(2 * stage), stream>>>(pf, (uint32_t)stage, (uint32_t)kwin);
      AMM_CUDA(cudaGetLastError());
      ws->launches += 1;
    }
    if (n_full >= p.n_windows) return AMM_OK;
    rest.w_first = n_full;
    rest.n_windows = p.n_windows - n_full;
  }
  // Fixed windows that are NOT a multiple of 16 bytes (16-byte aligned array): the leading full
  // windows whose 16-byte-rounded end stays inside the array take the 32-bit-geometry kernel;
  // the last one or two windows go to the general form.
  if (p.offsets == nullptr && p.w_first == 0 && (p.window * ES) % 16 != 0 && p.window * ES >= 16 &&
      ((uintptr_t)p.base % 16) == 0 && ws->win_fixed) {
    uint64_t n_fix = p.n / p.window;  // full windows
    const uint64_t n_bytes = p.n * ES;
    while (n_fix > 0 && ((n_fix * p.window * ES + 15) & ~(uint64_t)15) > n_bytes) --n_fix;
    if (n_fix > 0) {
      WindowParams pf = p;
      pf.n_windows = n_fix;
      uint64_t stage = WPB * p.window * ES + 32;
      stage = (stage + 127) / 128 * 128;
      argminmax_windows_fixed_kernel<P, L, THREADS>
          <<<grid_for(n_fix), THREADS, (size_t)(2 * stage), stream>>>(pf, (uint32_t)stage);
      AMM_CUDA(cudaGetLastError());
      ws->launches += 1;
    }
    if (n_fix >= p.n_windows) return AMM_OK;
    rest.w_first = n_fix;
    rest.n_windows = p.n_windows - n_fix;
  }
  // stage = the bytes of one block of windows (+ alignment slack); ragged windows get 2x the
  // mean (what does not fit is read from global memory by the kernel)
  const uint64_t fixed_bytes = (p.window < p.n ? p.window : p.n) * ES;
  uint64_t stage = (p.offsets ? 2 * mean_bytes : fixed_bytes) * WPB + 64;
  stage = (stage + 127) / 128 * 128;
  if (stage < 1024) stage = 1024;
  if (stage > 20480) stage = 20480;  // 2 stages + static smem stay below the 48 KB default limit
  argminmax_windows_staged_kernel<P, L, THREADS>
      <<<grid_for(rest.n_windows), THREADS, (size_t)(2 * stage), stream>>>(rest, (uint32_t)stage, 1u);
  AMM_CUDA(cudaGetLastError());
  ws->launches += 1;
  return AMM_OK;
}

template <class P>
static int launch_windows(amm_workspace* ws, const WindowParams& p, uint64_t mean_window,
                          cudaStream_t stream) {
  const uint64_t cap = (uint64_t)ws->sm_count * 8;
  uint64_t grid;
  // vector unit of the warp- / CTA-per-window kernels: 32 bytes for 8-byte dtypes
  constexpr int WVEC = sizeof(typename P::Elem) == 8 ? 32 : 16;
  if (p.chunk_first != nullptr) {  // combine step of the chunked form
    grid = (p.n_windows + 7) / 8;
    grid = grid > cap ? cap : (grid < 1 ? 1 : grid);
    argminmax_windows_combine_kernel<P><<<(unsigned)grid, 256, 0, stream>>>(p);
    AMM_CUDA(cudaGetLastError());
    ws->launches += 1;
    return AMM_OK;
  }
  const uint64_t mean_bytes = mean_window * sizeof(typename P::Elem);
  const uint64_t lane_bytes = (uint64_t)ws->win_lane_bytes;  // target bytes per lane
  // measured (tools/windows_bench.py): 8 lanes from global memory still win between 1 and 2 KB
  if (ws->win_staged && mean_bytes <= (uint64_t)ws->win_staged_max &&
      (mean_bytes <= 8 * lane_bytes || (mean_bytes > 16 * lane_bytes && mean_bytes <= 32 * lane_bytes))) {
    if (mean_bytes <= lane_bytes) return launch_windows_staged<P, 1>(ws, p, mean_bytes, stream);
    if (mean_bytes <= 2 * lane_bytes) return launch_windows_staged<P, 2>(ws, p, mean_bytes, stream);
    if (mean_bytes <= 4 * lane_bytes) return launch_windows_staged<P, 4>(ws, p, mean_bytes, stream);
    if (mean_bytes <= 8 * lane_bytes) return launch_windows_staged<P, 8>(ws, p, mean_bytes, stream);
    return launch_windows_staged<P, 32>(ws, p, mean_bytes, stream);
  }
  if (mean_window * sizeof(typename P::Elem) <= 2048) {
    grid = (p.n_windows + 31) / 32;
    grid = grid > cap ? cap : (grid < 1 ? 1 : grid);
    argminmax_windows_warp_kernel<P, 8><<<(unsigned)grid, 256, 0, stream>>>(p);
  } else if (p.n_windows >= 4096) {
    grid = (p.n_windows + 7) / 8;  // 8 warps per CTA, one window per warp
    grid = grid > cap ? cap : (grid < 1 ? 1 : grid);
    argminmax_windows_warp_kernel<P, 32, WVEC><<<(unsigned)grid, 256, 0, stream>>>(p);
  } else if (p.n_windows >= (uint64_t)ws->sm_count * 4) {
    grid = p.n_windows < cap ? p.n_windows : cap;
    argminmax_windows_cta_kernel<P, 256, 4, WVEC><<<(unsigned)grid, 256, 0, stream>>>(p);
  } else {
    grid = p.n_windows < cap ? p.n_windows : cap;
    grid = grid < 1 ? 1 : grid;
    argminmax_windows_cta_kernel<P, 512, (WVEC == 32 ? 4 : 8), WVEC><<<(unsigned)grid, 512, 0, stream>>>(p);
  }
  AMM_CUDA(cudaGetLastError());
  ws->launches += 1;
  return AMM_OK;
}

}  // namespace amm
