// windows_kernel.cuh -- segmented / windowed argmin+argmax: many sub-slices of one array in
// ONE launch.  There is no reference entry point for this; the
// semantics are composed from the reference's: for every window w
//     (imin, imax) = arr[begin_w .. end_w].argminmax()  (or .nanargminmax()),  + begin_w
// i.e. exactly what the crate's downstream users (MinMax / MinMaxLTTB down-samplers) compute by
// calling `argminmax` once per bin (src/lib.rs:134-244 semantics per bin), but without one
// kernel launch per bin.  An empty window yields (AMM_NO_INDEX, AMM_NO_INDEX).
//
// One WARP per window (windows up to a few thousand elements: the down-sampling case) or one
// CTA per window (large windows).
#pragma once
#include "scan_kernel.cuh"

namespace amm {

struct WindowParams {
  const uint8_t* base;
  uint64_t n;
  const uint64_t* offsets;  // n_windows + 1 ascending element offsets, or nullptr ...
  uint64_t window;          // ... for fixed-size windows of `window` elements (last may be short)
  uint64_t n_windows;
  uint64_t* out;            // 2 * n_windows: (argmin, argmax) per window, absolute indices
  int mode;
};

template <class P>
__device__ __forceinline__ void window_bounds(const WindowParams& p, uint64_t w, uint64_t& b,
                                              uint64_t& e) {
  if (p.offsets != nullptr) {
    b = p.offsets[w];
    e = p.offsets[w + 1];
  } else {
    b = w * p.window;
    e = b + p.window;
  }
  e = e < p.n ? e : p.n;
  b = b < e ? b : e;
}

// ---------------------------------------------------------------------------------------
// A window is cut into UNITS: unit 0 = the unaligned head (< one 16-byte vector, one element
// per lane), units 1..nv = the aligned 16-byte vectors, unit nv+1 = the sub-vector tail.  The
// fold is value-only per unit (policies.cuh, same instructions as the streaming kernel) and a
// thread only remembers WHICH unit gave its running min / max (strict compare, ascending units
// => earliest unit).  After the lexicographic (key, unit) reduce the winning unit -- at most 16
// elements -- is re-read with one element per lane and a ballot picks the first match.  Cost
// per element is that of the streaming kernel; the per-window overhead is one reduce + three
// single-element loads.
// ---------------------------------------------------------------------------------------
template <class P>
struct WindowGeom {
  static constexpr uint32_t VE = 16 / sizeof(typename P::Elem);
  uint64_t begin, end;  // elements
  uint64_t a0;          // first element of unit 1 (16-byte aligned address)
  uint32_t hc, tc;      // head / tail element counts (< VE)
  uint64_t nv;          // aligned vectors
  __device__ __forceinline__ WindowGeom(const uint8_t* base, uint64_t b, uint64_t e)
      : begin(b), end(e) {
    const uint64_t addr = (uint64_t)(uintptr_t)base + b * sizeof(typename P::Elem);
    uint64_t head = ((16 - (addr & 15)) & 15) / sizeof(typename P::Elem);
    const uint64_t len = e - b;
    head = head < len ? head : len;
    hc = (uint32_t)head;
    a0 = b + head;
    nv = (len - head) / VE;
    tc = (uint32_t)(len - head - nv * VE);
  }
  // element range of unit u
  __device__ __forceinline__ void unit_range(uint64_t u, uint64_t& first, uint32_t& count) const {
    if (u == 0) {
      first = begin;
      count = hc;
    } else if (u <= nv) {
      first = a0 + (u - 1) * VE;
      count = VE;
    } else {
      first = a0 + nv * VE;
      count = tc;
    }
  }
};

// value-only fold of this thread's share of the window; units are visited in ascending order
template <class P, int STRIDE, int UN = 4>
__device__ __forceinline__ void window_fold(const uint8_t* base, const WindowGeom<P>& g,
                                            uint32_t t, Best<typename P::Key>& x) {
  using Elem = typename P::Elem;
  using Key = typename P::Key;
  const Elem* e = reinterpret_cast<const Elem*>(base);
  auto consider = [&](Key kmin, Key kmax, bool valid, bool has_nan, uint64_t unit) {
    if (valid && (kmin < x.kmin || x.pmin == POS_NONE)) {
      x.kmin = kmin;
      x.pmin = unit;
    }
    if (valid && (kmax > x.kmax || x.pmax == POS_NONE)) {
      x.kmax = kmax;
      x.pmax = unit;
    }
    if (has_nan && x.pnan == POS_NONE) x.pnan = unit;
  };
  for (uint32_t h = t; h < g.hc; h += STRIDE) {  // unit 0: the head, one element at a time
    Key k;
    bool valid, is_nan;
    P::classify(e[g.begin + h], k, valid, is_nan);
    consider(k, k, valid, P::TRACK_NAN && is_nan, 0);
  }
  const uint8_t* vbase = base + g.a0 * sizeof(Elem);
  for (uint64_t v0 = t; v0 < g.nv; v0 += (uint64_t)UN * STRIDE) {
    uint32_t w[UN][4];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const uint64_t v = v0 + (uint64_t)u * STRIDE;
      if (v < g.nv) VecLoad<16>::ld(vbase + v * 16, w[u]);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const uint64_t v = v0 + (uint64_t)u * STRIDE;
      if (v < g.nv) {
        typename P::Acc acc;
        P::init(acc, &w[u][0]);
#pragma unroll
        for (int q = P::WPI; q < 4; q += P::WPI) P::feed(acc, &w[u][q]);
        Key kmin, kmax;
        bool valid, has_nan;
        P::finish(acc, kmin, kmax, valid, has_nan);
        consider(kmin, kmax, valid, has_nan, v + 1);
      }
    }
  }
  for (uint32_t q = t; q < g.tc; q += STRIDE) {  // unit nv+1: the tail
    Key k;
    bool valid, is_nan;
    P::classify(e[g.a0 + g.nv * WindowGeom<P>::VE + q], k, valid, is_nan);
    consider(k, k, valid, P::TRACK_NAN && is_nan, g.nv + 1);
  }
}

// A group of G lanes (G = 32: the whole warp; G = 8: four windows per warp, for windows of a few
// hundred bytes) re-reads the winning units, G elements per step, and a ballot picks the first
// match; lane 0 of the group writes the window's answer.  Every lane of the warp executes
// the same ballots (inactive groups contribute no hits).
template <class P, int G>
__device__ __forceinline__ void window_locate_store(const WindowParams& p, uint64_t w, bool active,
                                                    const WindowGeom<P>& g,
                                                    const Best<typename P::Key>& x, uint32_t lane) {
  using Elem = typename P::Elem;
  using Key = typename P::Key;
  const Elem* e = reinterpret_cast<const Elem*>(p.base);
  const uint32_t sl = lane % G, shift = (lane / G) * G;
  const uint32_t gmask = (G == 32) ? 0xFFFFFFFFu : ((1u << G) - 1u);
  uint64_t res[3] = {POS_NONE, POS_NONE, POS_NONE};  // min, max, first NaN
  const uint64_t units[3] = {x.pmin, x.pmax, x.pnan};
#pragma unroll
  for (int c = 0; c < (P::TRACK_NAN ? 3 : 2); ++c) {
    uint64_t first = 0;
    uint32_t count = 0;
    if (active && units[c] != POS_NONE) g.unit_range(units[c], first, count);
    constexpr uint32_t STEPS = (WindowGeom<P>::VE + G - 1) / G;  // units hold <= VE elements
#pragma unroll
    for (uint32_t st = 0; st < STEPS; ++st) {
      const uint32_t k_in_unit = st * G + sl;
      bool hit = false;
      if (k_in_unit < count) {
        Key k;
        bool valid, is_nan;
        P::classify(e[first + k_in_unit], k, valid, is_nan);
        hit = (c == 0) ? (valid && k == x.kmin) : (c == 1) ? (valid && k == x.kmax) : is_nan;
      }
      const uint32_t m = (__ballot_sync(0xFFFFFFFFu, hit) >> shift) & gmask;
      if (m != 0 && res[c] == POS_NONE) res[c] = first + st * G + (uint64_t)(__ffs(m) - 1);
    }
  }
  if (active && sl == 0) {
    uint64_t imin, imax;
    if (g.begin == g.end) {
      imin = imax = AMM_NO_INDEX;
    } else if (p.mode == MODE_RETURN && res[2] != POS_NONE) {
      imin = imax = res[2];  // first NaN of the window for both
    } else if (res[0] == POS_NONE) {
      imin = imax = g.begin;  // all NaN, ignore mode: the sub-slice answer is (0, 0)
    } else {
      imin = res[0];
      imax = res[1];
    }
    p.out[2 * w] = imin;
    p.out[2 * w + 1] = imax;
  }
}

// All-reduce of the (key, unit) candidates inside a group of G lanes, with 32-bit unit ids (a
// window has fewer than 2^32 units) and without the NaN field for policies that do not track
// NaNs: per-window cost matters when windows are a few hundred elements.
template <class P, int G>
__device__ __forceinline__ void window_group_reduce(Best<typename P::Key>& x) {
  using Key = typename P::Key;
  Key kmin = x.kmin, kmax = x.kmax;
  uint32_t umin = (uint32_t)x.pmin, umax = (uint32_t)x.pmax, unan = (uint32_t)x.pnan;  // NONE -> ~0u
#pragma unroll
  for (int m = G / 2; m >= 1; m >>= 1) {
    const Key okmin = __shfl_xor_sync(0xFFFFFFFFu, kmin, m);
    const Key okmax = __shfl_xor_sync(0xFFFFFFFFu, kmax, m);
    const uint32_t oumin = __shfl_xor_sync(0xFFFFFFFFu, umin, m);
    const uint32_t oumax = __shfl_xor_sync(0xFFFFFFFFu, umax, m);
    if (okmin < kmin || (okmin == kmin && oumin < umin)) {
      kmin = okmin;
      umin = oumin;
    }
    if (okmax > kmax || (okmax == kmax && oumax < umax)) {
      kmax = okmax;
      umax = oumax;
    }
    if (P::TRACK_NAN) {
      const uint32_t ounan = __shfl_xor_sync(0xFFFFFFFFu, unan, m);
      unan = ounan < unan ? ounan : unan;
    }
  }
  x.kmin = kmin;
  x.kmax = kmax;
  x.pmin = umin == 0xFFFFFFFFu ? POS_NONE : (uint64_t)umin;
  x.pmax = umax == 0xFFFFFFFFu ? POS_NONE : (uint64_t)umax;
  x.pnan = (!P::TRACK_NAN || unan == 0xFFFFFFFFu) ? POS_NONE : (uint64_t)unan;
}

// one group of G lanes per window (32 / G windows per warp), grid-stride over windows
template <class P, int G>
__global__ void __launch_bounds__(256) argminmax_windows_warp_kernel(const WindowParams p) {
  using Key = typename P::Key;
  constexpr uint32_t PER_WARP = 32 / G;
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  const uint64_t n_rounds = (p.n_windows + PER_WARP - 1) / PER_WARP;
  for (uint64_t r = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rounds;
       r += warps) {
    const uint64_t w = r * PER_WARP + lane / G;
    const bool active = w < p.n_windows;
    uint64_t begin = 0, end = 0;
    if (active) window_bounds<P>(p, w, begin, end);
    const WindowGeom<P> g(p.base, begin, end);
    Best<Key> x;
    x.clear();
    window_fold<P, G>(p.base, g, lane % G, x);
    window_group_reduce<P, G>(x);
    window_locate_store<P, G>(p, w, active, g, x, lane);
  }
}

// one CTA per window (large windows)
template <class P, int THREADS, int UN>
__global__ void __launch_bounds__(THREADS) argminmax_windows_cta_kernel(const WindowParams p) {
  using Key = typename P::Key;
  __shared__ Best<Key> smem[32];
  for (uint64_t w = blockIdx.x; w < p.n_windows; w += gridDim.x) {
    uint64_t begin, end;
    window_bounds<P>(p, w, begin, end);
    const WindowGeom<P> g(p.base, begin, end);
    Best<Key> x;
    x.clear();
    window_fold<P, THREADS, UN>(p.base, g, threadIdx.x, x);  // THREADS * UN * 16 B in flight
    x = block_reduce(x, smem);
    if (threadIdx.x < 32) window_locate_store<P, 32>(p, w, true, g, x, threadIdx.x);
  }
}

}  // namespace amm
