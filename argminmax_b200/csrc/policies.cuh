// policies.cuh -- per (dtype, NaN-mode) scan policies for the sm_100a argminmax kernel.
//
// A policy answers three questions for one of the 14 instantiations
// (8 ints + {f16,f32,f64} x {ignore-NaN, return-NaN}; reference: the 14
// src/simd/simd_<dtype>[_ignore_nan|_return_nan].rs backends):
//
//   1. Acc / init / feed : how a thread folds 32-bit words of raw input into a
//      VALUE-ONLY running min and max.  This is the whole hot loop -- no index is
//      carried per element, unlike the reference's per-lane (value, index) blends
//      (src/simd/generic.rs:390-406).  Cost per element on B200:
//        f32      FMNMX / FMNMX3                       (~1 op per direction)
//        f16      HMNMX2 on two packed halves          (0.5 op per direction)
//        i16/u16  VIMNMX.{S,U}16x2                     (0.5 op per direction)
//        i8/u8    VIMNMX.16x2 on w and (w << 8): the HIGH byte of every 16-bit
//                 lane of a 16-bit min/max is the min/max of the high bytes, so
//                 odd bytes need no unpacking and even bytes one shift
//                                                      (~0.3 op per direction)
//        i32/u32  VIMNMX / VIMNMX3
//        64-bit   ISETP pair + 2 SEL (ints) / DSETP.MIN/MAX (f64; + one DSETP.NAN in return mode)
//   2. finish : collapse the accumulator to one canonical ordered key per
//      direction, plus "this tile holds a non-NaN value" and "this tile holds a
//      NaN".  Canonical keys are signed integers whose order IS the reference's
//      order (the reference's key transforms: unsigned `v ^ SIGNED_MIN`
//      src/simd/simd_u32.rs:56; float -> ord `((v >> W-1) & MAX) ^ v`
//      src/simd/simd_f32_return_nan.rs:5-9, src/scalar/scalar_f16.rs:10-13).
//      f32/f64 keys canonicalise -0.0 to +0.0 first because the reference's scalar
//      path compares them as IEEE values (src/scalar/generic.rs:208-214); f16 keys
//      do not, because the reference orders f16 through the i16 ord key
//      (src/scalar/scalar_f16.rs:119-131), where -0.0 < +0.0.
//   3. classify : the same key for ONE raw element (exact scan of the few
//      candidate ranges at the end, see scan_kernel.cuh).
//
// NaN handling needs no per-element work: ignore-NaN uses the hardware's
// NaN-dropping min/max (FMNMX / HMNMX2 / DSETP.MIN return the non-NaN operand);
// return-NaN uses the NaN-propagating form (FMNMX.NAN / HMNMX2.NAN) or, for f64,
// ord keys in which every NaN sorts outside [-inf, +inf].
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace amm {

enum : int { MODE_IGNORE = 0, MODE_RETURN = 1 };

template <typename K>
struct KeyLim;
template <>
struct KeyLim<int32_t> {
  static constexpr int32_t MAXV = INT32_MAX;
  static constexpr int32_t MINV = INT32_MIN;
};
template <>
struct KeyLim<int64_t> {
  static constexpr int64_t MAXV = INT64_MAX;
  static constexpr int64_t MINV = INT64_MIN;
};

// compare domains of the window kernels for f32 / f64 (windows_kernel.cuh: WinCmp): +-infinity
template <>
struct KeyLim<float> {
  static constexpr float MAXV = __builtin_huge_valf();
  static constexpr float MINV = -__builtin_huge_valf();
};
template <>
struct KeyLim<double> {
  static constexpr double MAXV = __builtin_huge_val();
  static constexpr double MINV = -__builtin_huge_val();
};

// --------------------------------------------------------------- 32-bit ints
template <bool SIGNED>
struct PolInt32 {
  using Elem = uint32_t;
  using Key = int32_t;
  static constexpr int WPI = 1;  // 32-bit words consumed per feed()
  static constexpr bool TRACK_NAN = false;
  static constexpr bool SIMD_EQ = false;  // no packed equality compare for this element width
  struct Acc {
    uint32_t mn, mx;
  };
  static __device__ __forceinline__ void init(Acc& a, const uint32_t* w) { a.mn = a.mx = w[0]; }
  static __device__ __forceinline__ void feed(Acc& a, const uint32_t* w) {
    if (SIGNED) {
      a.mn = (uint32_t)min((int32_t)a.mn, (int32_t)w[0]);
      a.mx = (uint32_t)max((int32_t)a.mx, (int32_t)w[0]);
    } else {
      a.mn = min(a.mn, w[0]);
      a.mx = max(a.mx, w[0]);
    }
  }
  static __device__ __forceinline__ Key key(uint32_t v) {
    return SIGNED ? (int32_t)v : (int32_t)(v ^ 0x80000000u);  // simd_u32.rs:56
  }
  static __device__ __forceinline__ void finish(const Acc& a, Key& kmin, Key& kmax, bool& valid,
                                                bool& has_nan) {
    kmin = key(a.mn);
    kmax = key(a.mx);
    valid = true;
    has_nan = false;
  }
  static __device__ __forceinline__ void classify(Elem raw, Key& k, bool& valid, bool& is_nan) {
    k = key(raw);
    valid = true;
    is_nan = false;
  }
};

// --------------------------------------------------------------- 16-bit ints
template <bool SIGNED>
struct PolInt16 {
  using Elem = uint16_t;
  using Key = int32_t;
  static constexpr int WPI = 1;
  static constexpr bool TRACK_NAN = false;
  // no packed equality locate: with 8 trivially-keyed elements per vector the scalar compare chain
  // is shorter than __vcmpeq2 + find-first-set per word (measured: i16 windows 8-13 % slower with it)
  static constexpr bool SIMD_EQ = false;
  struct Acc {
    uint32_t mn, mx;  // two packed 16-bit lanes each
  };
  static __device__ __forceinline__ uint32_t vmin(uint32_t a, uint32_t b) {
    return SIGNED ? __vmins2(a, b) : __vminu2(a, b);
  }
  static __device__ __forceinline__ uint32_t vmax(uint32_t a, uint32_t b) {
    return SIGNED ? __vmaxs2(a, b) : __vmaxu2(a, b);
  }
  static __device__ __forceinline__ void init(Acc& a, const uint32_t* w) { a.mn = a.mx = w[0]; }
  static __device__ __forceinline__ void feed(Acc& a, const uint32_t* w) {
    a.mn = vmin(a.mn, w[0]);
    a.mx = vmax(a.mx, w[0]);
  }
  static __device__ __forceinline__ Key key(uint16_t v) {
    return SIGNED ? (int32_t)(int16_t)v : (int32_t)v;
  }
  static __device__ __forceinline__ void finish(const Acc& a, Key& kmin, Key& kmax, bool& valid,
                                                bool& has_nan) {
    kmin = min(key((uint16_t)(a.mn & 0xFFFFu)), key((uint16_t)(a.mn >> 16)));
    kmax = max(key((uint16_t)(a.mx & 0xFFFFu)), key((uint16_t)(a.mx >> 16)));
    valid = true;
    has_nan = false;
  }
  static __device__ __forceinline__ void classify(Elem raw, Key& k, bool& valid, bool& is_nan) {
    k = key(raw);
    valid = true;
    is_nan = false;
  }
};

// ---------------------------------------------------------------- 8-bit ints
// Value-only min/max of bytes with 16-bit packed ops: for 16-bit lanes
// L = (hi << 8) | lo, min16(L1, L2) has hi byte min(hi1, hi2) whatever the low
// bytes are (signed lanes: the lane's sign is hi's sign).  So `w` tracks the odd
// bytes in place and `w << 8` tracks the even bytes; low bytes are don't-care.
template <bool SIGNED>
struct PolInt8 {
  using Elem = uint8_t;
  using Key = int32_t;
  static constexpr int WPI = 1;
  static constexpr bool TRACK_NAN = false;
  // first element of a 32-bit word whose raw bits equal the (bijective) key's: packed compare
  // (__vcmpeq4) + find-first-set instead of an extract + classify + compare per element (16 per
  // vector here; u8 / f16 windows of 16-1000 elements +4-23 %)
  static constexpr bool SIMD_EQ = true;
  static constexpr int PER_WORD = 4;
  static __device__ __forceinline__ uint32_t raw_bcast(Key k) { return (uint32_t)(uint8_t)k * 0x01010101u; }
  static __device__ __forceinline__ int first_eq(uint32_t w, uint32_t bc) {
    const uint32_t m = __vcmpeq4(w, bc);
    return m ? (__ffs(m) - 1) >> 3 : -1;
  }
  struct Acc {
    uint32_t mnA, mnB, mxA, mxB;  // A: odd bytes, B: even bytes (both in lane-high position)
  };
  static __device__ __forceinline__ uint32_t vmin(uint32_t a, uint32_t b) {
    return SIGNED ? __vmins2(a, b) : __vminu2(a, b);
  }
  static __device__ __forceinline__ uint32_t vmax(uint32_t a, uint32_t b) {
    return SIGNED ? __vmaxs2(a, b) : __vmaxu2(a, b);
  }
  static __device__ __forceinline__ void init(Acc& a, const uint32_t* w) {
    a.mnA = a.mxA = w[0];
    a.mnB = a.mxB = w[0] << 8;
  }
  static __device__ __forceinline__ void feed(Acc& a, const uint32_t* w) {
    const uint32_t s = w[0] << 8;
    a.mnA = vmin(a.mnA, w[0]);
    a.mnB = vmin(a.mnB, s);
    a.mxA = vmax(a.mxA, w[0]);
    a.mxB = vmax(a.mxB, s);
  }
  static __device__ __forceinline__ Key key(uint8_t v) {
    return SIGNED ? (int32_t)(int8_t)v : (int32_t)v;
  }
  static __device__ __forceinline__ void finish(const Acc& a, Key& kmin, Key& kmax, bool& valid,
                                                bool& has_nan) {
    const uint32_t m = vmin(a.mnA, a.mnB), M = vmax(a.mxA, a.mxB);
    kmin = min(key((uint8_t)(m >> 8)), key((uint8_t)(m >> 24)));
    kmax = max(key((uint8_t)(M >> 8)), key((uint8_t)(M >> 24)));
    valid = true;
    has_nan = false;
  }
  static __device__ __forceinline__ void classify(Elem raw, Key& k, bool& valid, bool& is_nan) {
    k = key(raw);
    valid = true;
    is_nan = false;
  }
};

// --------------------------------------------------------------- 64-bit ints
template <bool SIGNED>
struct PolInt64 {
  using Elem = uint64_t;
  using Key = int64_t;
  static constexpr int WPI = 2;
  static constexpr bool TRACK_NAN = false;
  static constexpr bool SIMD_EQ = false;  // no packed equality compare for this element width
  struct Acc {
    uint64_t mn, mx;
  };
  static __device__ __forceinline__ uint64_t pack(const uint32_t* w) {
    return ((uint64_t)w[1] << 32) | (uint64_t)w[0];
  }
  static __device__ __forceinline__ void init(Acc& a, const uint32_t* w) { a.mn = a.mx = pack(w); }
  static __device__ __forceinline__ void feed(Acc& a, const uint32_t* w) {
    const uint64_t v = pack(w);
    if (SIGNED) {
      a.mn = (uint64_t)min((long long)a.mn, (long long)v);
      a.mx = (uint64_t)max((long long)a.mx, (long long)v);
    } else {
      a.mn = min((unsigned long long)a.mn, (unsigned long long)v);
      a.mx = max((unsigned long long)a.mx, (unsigned long long)v);
    }
  }
  static __device__ __forceinline__ Key key(uint64_t v) {
    return SIGNED ? (int64_t)v : (int64_t)(v ^ 0x8000000000000000ull);  // simd_u64.rs:52
  }
  static __device__ __forceinline__ void finish(const Acc& a, Key& kmin, Key& kmax, bool& valid,
                                                bool& has_nan) {
    kmin = key(a.mn);
    kmax = key(a.mx);
    valid = true;
    has_nan = false;
  }
  static __device__ __forceinline__ void classify(Elem raw, Key& k, bool& valid, bool& is_nan) {
    k = key(raw);
    valid = true;
    is_nan = false;
  }
};

// ----------------------------------------------------------------------- f32
template <int MODE>
struct PolF32 {
  using Elem = uint32_t;  // raw IEEE bits
  using Key = int32_t;
  static constexpr int WPI = 1;
  static constexpr bool TRACK_NAN = (MODE == MODE_RETURN);
  static constexpr bool SIMD_EQ = false;  // no packed equality compare for this element width
  struct Acc {
    float mn, mx;
  };
  static __device__ __forceinline__ float min_nan(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));  // FMNMX.NAN
    return r;
  }
  static __device__ __forceinline__ void init(Acc& a, const uint32_t* w) {
    a.mn = a.mx = __uint_as_float(w[0]);
  }
  static __device__ __forceinline__ void feed(Acc& a, const uint32_t* w) {
    const float x = __uint_as_float(w[0]);
    a.mn = (MODE == MODE_RETURN) ? min_nan(a.mn, x) : fminf(a.mn, x);
    a.mx = fmaxf(a.mx, x);
  }
  static __device__ __forceinline__ bool nan_bits(uint32_t b) {
    return (b & 0x7FFFFFFFu) > 0x7F800000u;
  }
  // -0.0 -> +0.0, then the reference's ord transform (simd_f32_return_nan.rs:5-9)
  static __device__ __forceinline__ Key key(uint32_t b) {
    b = (b == 0x80000000u) ? 0u : b;
    const int32_t s = (int32_t)b;
    return s ^ ((s >> 31) & 0x7FFFFFFF);
  }
  static __device__ __forceinline__ void finish(const Acc& a, Key& kmin, Key& kmax, bool& valid,
                                                bool& has_nan) {
    const uint32_t bmn = __float_as_uint(a.mn);
    const bool mn_nan = nan_bits(bmn);  // ignore: every element NaN; return: any element NaN
    has_nan = (MODE == MODE_RETURN) && mn_nan;
    valid = !mn_nan;
    kmin = key(bmn);
    kmax = key(__float_as_uint(a.mx));
  }
  static __device__ __forceinline__ void classify(Elem raw, Key& k, bool& valid, bool& is_nan) {
    is_nan = nan_bits(raw);
    valid = !is_nan;
    k = key(raw);
  }
};

// ----------------------------------------------------------------------- f16
template <int MODE>
struct PolF16 {
  using Elem = uint16_t;  // raw binary16 bits
  using Key = int32_t;
  static constexpr int WPI = 1;
  static constexpr bool TRACK_NAN = (MODE == MODE_RETURN);
  static constexpr bool SIMD_EQ = true;  // see PolInt8; the ord transform is its own inverse
  static constexpr int PER_WORD = 2;
  static __device__ __forceinline__ uint32_t raw_bcast(Key k) {
    const int16_t s = (int16_t)k;
    const uint32_t r = (uint32_t)(uint16_t)(s ^ ((int16_t)(s >> 15) & 0x7FFF));
    return r | (r << 16);
  }
  static __device__ __forceinline__ int first_eq(uint32_t w, uint32_t bc) {
    const uint32_t m = __vcmpeq2(w, bc);
    return m ? (__ffs(m) - 1) >> 4 : -1;
  }
  struct Acc {
    __half2 mn, mx;
  };
  static __device__ __forceinline__ __half2 h2(uint32_t w) {
    return *reinterpret_cast<const __half2*>(&w);
  }
  static __device__ __forceinline__ void init(Acc& a, const uint32_t* w) { a.mn = a.mx = h2(w[0]); }
  static __device__ __forceinline__ void feed(Acc& a, const uint32_t* w) {
    const __half2 x = h2(w[0]);
    a.mn = (MODE == MODE_RETURN) ? __hmin2_nan(a.mn, x) : __hmin2(a.mn, x);
    a.mx = __hmax2(a.mx, x);
  }
  static __device__ __forceinline__ bool nan_bits(uint16_t b) { return (b & 0x7FFFu) > 0x7C00u; }
  // f16_to_i16ord, scalar_f16.rs:10-13 (no zero canonicalisation: -0.0 < +0.0 for f16)
  static __device__ __forceinline__ Key key(uint16_t b) {
    const int16_t s = (int16_t)b;
    return (int32_t)(int16_t)(s ^ ((int16_t)(s >> 15) & 0x7FFF));
  }
  static __device__ __forceinline__ void finish(const Acc& a, Key& kmin, Key& kmax, bool& valid,
                                                bool& has_nan) {
    const __half lo = __low2half(a.mn), hi = __high2half(a.mn);
    const __half m = (MODE == MODE_RETURN) ? __hmin_nan(lo, hi) : __hmin(lo, hi);
    const __half M = __hmax(__low2half(a.mx), __high2half(a.mx));
    const uint16_t bmn = __half_as_ushort(m);
    const bool mn_nan = nan_bits(bmn);
    has_nan = (MODE == MODE_RETURN) && mn_nan;
    valid = !mn_nan;
    kmin = key(bmn);
    kmax = key(__half_as_ushort(M));
  }
  static __device__ __forceinline__ void classify(Elem raw, Key& k, bool& valid, bool& is_nan) {
    is_nan = nan_bits(raw);
    valid = !is_nan;
    k = key(raw);
  }
};

// ----------------------------------------------------------------------- f64
static __device__ __forceinline__ bool nan_bits64(uint64_t b) {
  return (b & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull;
}
// -0.0 -> +0.0, then ord64 (simd_f64_return_nan.rs:70-77)
static __device__ __forceinline__ int64_t key_f64(uint64_t b) {
  b = (b == 0x8000000000000000ull) ? 0ull : b;
  const int64_t s = (int64_t)b;
  return s ^ ((s >> 63) & 0x7FFFFFFFFFFFFFFFll);
}

template <int MODE>
struct PolF64;

// ignore-NaN: native double min/max (DSETP.MIN/MAX drop NaNs)
template <>
struct PolF64<MODE_IGNORE> {
  using Elem = uint64_t;
  using Key = int64_t;
  static constexpr int WPI = 2;
  static constexpr bool TRACK_NAN = false;
  static constexpr bool SIMD_EQ = false;  // no packed equality compare for this element width
  struct Acc {
    double mn, mx;
  };
  static __device__ __forceinline__ double pack(const uint32_t* w) {
    return __hiloint2double((int)w[1], (int)w[0]);
  }
  static __device__ __forceinline__ void init(Acc& a, const uint32_t* w) { a.mn = a.mx = pack(w); }
  static __device__ __forceinline__ void feed(Acc& a, const uint32_t* w) {
    const double x = pack(w);
    a.mn = fmin(a.mn, x);
    a.mx = fmax(a.mx, x);
  }
  static __device__ __forceinline__ void finish(const Acc& a, Key& kmin, Key& kmax, bool& valid,
                                                bool& has_nan) {
    const uint64_t bmn = (uint64_t)__double_as_longlong(a.mn);
    valid = !nan_bits64(bmn);
    has_nan = false;
    kmin = key_f64(bmn);
    kmax = key_f64((uint64_t)__double_as_longlong(a.mx));
  }
  static __device__ __forceinline__ void classify(Elem raw, Key& k, bool& valid, bool& is_nan) {
    is_nan = nan_bits64(raw);
    valid = !is_nan;
    k = key_f64(raw);
  }
};

// return-NaN: the same native double min/max as ignore-NaN (they DROP NaNs, there is no
// min.NaN.f64) plus one unordered self-compare per element folded into a flag -- 2 DSETP.MIN/MAX
// + 1 DSETP.NAN per element instead of an ord-key transform and two 64-bit integer min/max
// (ISETP pairs + selects, ~11 integer instructions per element: 6.8 TB/s at 4 GiB where the
// other dtypes run at 7.3-7.5).
template <>
struct PolF64<MODE_RETURN> {
  using Elem = uint64_t;
  using Key = int64_t;
  static constexpr int WPI = 2;
  static constexpr bool TRACK_NAN = true;
  static constexpr bool SIMD_EQ = false;  // no packed equality compare for this element width
  struct Acc {
    double mn, mx;
    bool nan;
  };
  static __device__ __forceinline__ double pack(const uint32_t* w) {
    return __hiloint2double((int)w[1], (int)w[0]);
  }
  static __device__ __forceinline__ void init(Acc& a, const uint32_t* w) {
    a.mn = a.mx = pack(w);
    a.nan = a.mn != a.mn;
  }
  static __device__ __forceinline__ void feed(Acc& a, const uint32_t* w) {
    const double x = pack(w);
    a.mn = fmin(a.mn, x);
    a.mx = fmax(a.mx, x);
    a.nan = a.nan || (x != x);
  }
  static __device__ __forceinline__ void finish(const Acc& a, Key& kmin, Key& kmax, bool& valid,
                                                bool& has_nan) {
    has_nan = a.nan;
    valid = !has_nan;  // any NaN decides the answer; the vector's values then do not matter
    kmin = key_f64((uint64_t)__double_as_longlong(a.mn));
    kmax = key_f64((uint64_t)__double_as_longlong(a.mx));
  }
  static __device__ __forceinline__ void classify(Elem raw, Key& k, bool& valid, bool& is_nan) {
    is_nan = nan_bits64(raw);
    valid = !is_nan;
    k = key_f64(raw);
  }
};

// index of the first element of a 16-byte vector whose key equals `want` (a key of a non-NaN
// element), or -1; policies with SIMD_EQ only
template <class P>
__device__ __forceinline__ int vec_first_eq(const uint32_t (&w)[4], typename P::Key want) {
  const uint32_t bc = P::raw_bcast(want);
  int found = -1;
#pragma unroll
  for (int i = 3; i >= 0; --i) {
    const int f = P::first_eq(w[i], bc);
    found = f >= 0 ? i * P::PER_WORD + f : found;
  }
  return found;
}

}  // namespace amm
