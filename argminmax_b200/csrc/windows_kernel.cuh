// windows_kernel.cuh -- segmented / windowed argmin+argmax: many sub-slices of one array in
// ONE launch.  There is no reference entry point for this; the
// semantics are composed from the reference's: for every window w
//     (imin, imax) = arr[begin_w .. end_w].argminmax()  (or .nanargminmax()),  + begin_w
// i.e. exactly what the crate's downstream users (MinMax / MinMaxLTTB down-samplers) compute by
// calling `argminmax` once per bin (src/lib.rs:134-244 semantics per bin), but without one
// kernel launch per bin.  An empty window yields (AMM_NO_INDEX, AMM_NO_INDEX).
//
// One WARP per window (windows up to a few thousand elements: the down-sampling case) or one
// CTA per window (large windows); exact per-element (key, index) tracking with the policies'
// classify(), first occurrence by strict compare (indices ascend per lane) and a
// lexicographic (key, index) shuffle reduce -- the same tie-break rule as everywhere else.
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

template <typename Key>
__device__ __forceinline__ void window_store(const WindowParams& p, uint64_t w, uint64_t begin,
                                             uint64_t end, const Best<Key>& x, int mode) {
  uint64_t imin, imax;
  if (begin == end) {
    imin = imax = AMM_NO_INDEX;
  } else if (mode == MODE_RETURN && x.pnan != POS_NONE) {
    imin = imax = x.pnan;  // first NaN of the window for both
  } else if (x.pmin == POS_NONE) {
    imin = imax = begin;  // all NaN, ignore mode: the sub-slice answer is (0, 0)
  } else {
    imin = x.pmin;
    imax = x.pmax;
  }
  p.out[2 * w] = imin;
  p.out[2 * w + 1] = imax;
}

template <class P, int STRIDE>
__device__ __forceinline__ void window_scan(const WindowParams& p, uint64_t begin, uint64_t end,
                                            uint32_t lane, Best<typename P::Key>& x) {
  using Elem = typename P::Elem;
  using Key = typename P::Key;
  const Elem* e = reinterpret_cast<const Elem*>(p.base);
  constexpr int UN = 4;
  for (uint64_t i0 = begin + lane; i0 < end; i0 += (uint64_t)UN * STRIDE) {
    Elem raw[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const uint64_t i = i0 + (uint64_t)u * STRIDE;
      raw[u] = (i < end) ? e[i] : Elem(0);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const uint64_t i = i0 + (uint64_t)u * STRIDE;
      if (i < end) {
        Key k;
        bool valid, is_nan;
        P::classify(raw[u], k, valid, is_nan);
        if (valid && (k < x.kmin || x.pmin == POS_NONE)) {
          x.kmin = k;
          x.pmin = i;
        }
        if (valid && (k > x.kmax || x.pmax == POS_NONE)) {
          x.kmax = k;
          x.pmax = i;
        }
        if (P::TRACK_NAN && is_nan && x.pnan == POS_NONE) x.pnan = i;
      }
    }
  }
}

// one warp per window, grid-stride over windows
template <class P>
__global__ void __launch_bounds__(256) argminmax_windows_warp_kernel(const WindowParams p) {
  using Key = typename P::Key;
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  for (uint64_t w = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < p.n_windows;
       w += warps) {
    uint64_t begin, end;
    window_bounds<P>(p, w, begin, end);
    Best<Key> x;
    x.clear();
    window_scan<P, 32>(p, begin, end, lane, x);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) x.merge(shfl_xor(x, m));
    if (lane == 0) window_store(p, w, begin, end, x, p.mode);
  }
}

// one CTA per window (large windows)
template <class P>
__global__ void __launch_bounds__(256) argminmax_windows_cta_kernel(const WindowParams p) {
  using Key = typename P::Key;
  __shared__ Best<Key> smem[32];
  for (uint64_t w = blockIdx.x; w < p.n_windows; w += gridDim.x) {
    uint64_t begin, end;
    window_bounds<P>(p, w, begin, end);
    Best<Key> x;
    x.clear();
    window_scan<P, 256>(p, begin, end, threadIdx.x, x);
    x = block_reduce(x, smem);
    if (threadIdx.x == 0) window_store(p, w, begin, end, x, p.mode);
  }
}

}  // namespace amm
