/*
 * oracle/argminmax_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference crate's *scalar* argmin/argmax
 * algorithms; this is the semantic oracle every parity test compares the CUDA
 * path against.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library
 * (argminmax_b200/csrc) never links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED -- tests/test_oracle_golden.py checks every function
 * here against the reference's own known-answer tests, transcribed with
 * file:line cites into tests/golden/golden_vectors.json.
 * The reference itself (Rust, needs cargo + crates.io) cannot be built in this
 * image, so there is no oracle/_ref.
 *
 * What is restated (all paths relative to the reference repository root):
 *   ints, f32, f64  : SCALAR<S>::argminmax / argmin / argmax
 *                       src/scalar/generic.rs:181-217, :220-250, :252-281
 *                     with the per-strategy init rules
 *                       Int            src/scalar/generic.rs:82-107
 *                       FloatReturnNaN src/scalar/generic.rs:109-136
 *                       FloatIgnoreNaN src/scalar/generic.rs:138-171
 *   f16             : scalar_argminmax_f16_return_nan  src/scalar/scalar_f16.rs:20-48
 *                     scalar_argminmax_f16_ignore_nan  src/scalar/scalar_f16.rs:100-136
 *                     (+ the single argmin/argmax variants :50-98, :138-180)
 *                     on the i16 "ord" key of src/scalar/scalar_f16.rs:10-13.
 *   f16 is carried as raw uint16_t bits; is_nan(f16) == ((bits & 0x7FFF) > 0x7C00)
 *   is what half::f16::is_nan computes (half 2.x, binary16 layout).
 *
 * mode: 0 = ArgMinMax (ints; floats ignore NaN), 1 = NaNArgMinMax (floats return
 * the index of the first NaN for both outputs).  mode is ignored for ints.
 * Return value: 0 ok, 1 empty input (the reference asserts/panics,
 * src/scalar/generic.rs:182).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <math.h>

#define ORACLE_OK 0
#define ORACLE_ERR_EMPTY 1

/* ------------------------------------------------------------------ ints --- */
/* SCALAR<Int>: low = high = arr[0]; no NaN rules (generic.rs:82-107, 181-217). */
#define DEF_INT(NAME, T)                                                              \
  int oracle_argminmax_##NAME(const T* arr, uint64_t n, int mode, uint64_t out[2]) {  \
    (void)mode;                                                                       \
    if (n == 0) return ORACLE_ERR_EMPTY;                                              \
    uint64_t low_index = 0, high_index = 0;                                           \
    T low = arr[0], high = arr[0];                                                    \
    for (uint64_t i = 0; i < n; ++i) {                                                \
      T v = arr[i];                                                                   \
      if (v < low) {                                                                  \
        low = v;                                                                      \
        low_index = i;                                                                \
      } else if (v > high) {                                                          \
        high = v;                                                                     \
        high_index = i;                                                               \
      }                                                                               \
    }                                                                                 \
    out[0] = low_index;                                                               \
    out[1] = high_index;                                                              \
    return ORACLE_OK;                                                                 \
  }                                                                                   \
  int oracle_argmin_##NAME(const T* arr, uint64_t n, int mode, uint64_t* out) {       \
    (void)mode;                                                                       \
    if (n == 0) return ORACLE_ERR_EMPTY;                                              \
    uint64_t low_index = 0;                                                           \
    T low = arr[0];                                                                   \
    for (uint64_t i = 0; i < n; ++i) {                                                \
      T v = arr[i];                                                                   \
      if (v < low) {                                                                  \
        low = v;                                                                      \
        low_index = i;                                                                \
      }                                                                               \
    }                                                                                 \
    *out = low_index;                                                                 \
    return ORACLE_OK;                                                                 \
  }                                                                                   \
  int oracle_argmax_##NAME(const T* arr, uint64_t n, int mode, uint64_t* out) {       \
    (void)mode;                                                                       \
    if (n == 0) return ORACLE_ERR_EMPTY;                                              \
    uint64_t high_index = 0;                                                          \
    T high = arr[0];                                                                  \
    for (uint64_t i = 0; i < n; ++i) {                                                \
      T v = arr[i];                                                                   \
      if (v > high) {                                                                 \
        high = v;                                                                     \
        high_index = i;                                                               \
      }                                                                               \
    }                                                                                 \
    *out = high_index;                                                                \
    return ORACLE_OK;                                                                 \
  }

DEF_INT(i8, int8_t)
DEF_INT(i16, int16_t)
DEF_INT(i32, int32_t)
DEF_INT(i64, int64_t)
DEF_INT(u8, uint8_t)
DEF_INT(u16, uint16_t)
DEF_INT(u32, uint32_t)
DEF_INT(u64, uint64_t)

/* ------------------------------------------------------------- f32 / f64 --- */
/*
 * SCALAR<FloatReturnNaN> (mode 1): _RETURN_AT_NAN = true, init low = high = arr[0],
 *   no first-non-NaN rule (generic.rs:109-136).
 * SCALAR<FloatIgnoreNaN> (mode 0): init low = +inf / high = -inf when arr[0] is
 *   NaN, and in that case the first non-NaN element sets both (generic.rs:138-171,
 *   197-207).  All compares are IEEE `<` / `>` (false for NaN; -0.0 == +0.0).
 */
#define DEF_FLOAT(NAME, T, ISNAN, PINF)                                               \
  int oracle_argminmax_##NAME(const T* arr, uint64_t n, int mode, uint64_t out[2]) {  \
    if (n == 0) return ORACLE_ERR_EMPTY;                                              \
    const int return_at_nan = (mode == 1);                                            \
    uint64_t low_index = 0, high_index = 0;                                           \
    T start = arr[0];                                                                 \
    T low = start, high = start;                                                      \
    int first_non_nan_update = 0;                                                     \
    if (!return_at_nan && ISNAN(start)) {                                             \
      low = PINF;                                                                     \
      high = -PINF;                                                                   \
      first_non_nan_update = 1;                                                       \
    }                                                                                 \
    for (uint64_t i = 0; i < n; ++i) {                                                \
      T v = arr[i];                                                                   \
      if (return_at_nan && ISNAN(v)) {                                                \
        out[0] = i;                                                                   \
        out[1] = i;                                                                   \
        return ORACLE_OK;                                                             \
      }                                                                               \
      if (first_non_nan_update) {                                                     \
        if (!ISNAN(v)) {                                                              \
          low = v;                                                                    \
          low_index = i;                                                              \
          high = v;                                                                   \
          high_index = i;                                                             \
          first_non_nan_update = 0;                                                   \
        }                                                                             \
      } else if (v < low) {                                                           \
        low = v;                                                                      \
        low_index = i;                                                                \
      } else if (v > high) {                                                          \
        high = v;                                                                     \
        high_index = i;                                                               \
      }                                                                               \
    }                                                                                 \
    out[0] = low_index;                                                               \
    out[1] = high_index;                                                              \
    return ORACLE_OK;                                                                 \
  }                                                                                   \
  int oracle_argmin_##NAME(const T* arr, uint64_t n, int mode, uint64_t* out) {       \
    if (n == 0) return ORACLE_ERR_EMPTY;                                              \
    const int return_at_nan = (mode == 1);                                            \
    uint64_t low_index = 0;                                                           \
    T start = arr[0];                                                                 \
    T low = start;                                                                    \
    int first_non_nan_update = 0;                                                     \
    if (!return_at_nan && ISNAN(start)) {                                             \
      low = PINF;                                                                     \
      first_non_nan_update = 1;                                                       \
    }                                                                                 \
    for (uint64_t i = 0; i < n; ++i) {                                                \
      T v = arr[i];                                                                   \
      if (return_at_nan && ISNAN(v)) {                                                \
        *out = i;                                                                     \
        return ORACLE_OK;                                                             \
      }                                                                               \
      if (first_non_nan_update) {                                                     \
        if (!ISNAN(v)) {                                                              \
          low = v;                                                                    \
          low_index = i;                                                              \
          first_non_nan_update = 0;                                                   \
        }                                                                             \
      } else if (v < low) {                                                           \
        low = v;                                                                      \
        low_index = i;                                                                \
      }                                                                               \
    }                                                                                 \
    *out = low_index;                                                                 \
    return ORACLE_OK;                                                                 \
  }                                                                                   \
  int oracle_argmax_##NAME(const T* arr, uint64_t n, int mode, uint64_t* out) {       \
    if (n == 0) return ORACLE_ERR_EMPTY;                                              \
    const int return_at_nan = (mode == 1);                                            \
    uint64_t high_index = 0;                                                          \
    T start = arr[0];                                                                 \
    T high = start;                                                                   \
    int first_non_nan_update = 0;                                                     \
    if (!return_at_nan && ISNAN(start)) {                                             \
      high = -PINF;                                                                   \
      first_non_nan_update = 1;                                                       \
    }                                                                                 \
    for (uint64_t i = 0; i < n; ++i) {                                                \
      T v = arr[i];                                                                   \
      if (return_at_nan && ISNAN(v)) {                                                \
        *out = i;                                                                     \
        return ORACLE_OK;                                                             \
      }                                                                               \
      if (first_non_nan_update) {                                                     \
        if (!ISNAN(v)) {                                                              \
          high = v;                                                                   \
          high_index = i;                                                             \
          first_non_nan_update = 0;                                                   \
        }                                                                             \
      } else if (v > high) {                                                          \
        high = v;                                                                     \
        high_index = i;                                                               \
      }                                                                               \
    }                                                                                 \
    *out = high_index;                                                                \
    return ORACLE_OK;                                                                 \
  }

/* bit-level NaN tests so that -ffast-math style folding can never remove them */
static inline int isnan_f32(float v) {
  uint32_t b;
  memcpy(&b, &v, 4);
  return (b & 0x7FFFFFFFu) > 0x7F800000u;
}
static inline int isnan_f64(double v) {
  uint64_t b;
  memcpy(&b, &v, 8);
  return (b & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull;
}

DEF_FLOAT(f32, float, isnan_f32, ((float)INFINITY))
DEF_FLOAT(f64, double, isnan_f64, ((double)INFINITY))

/* -------------------------------------------------------------------- f16 --- */
/* f16_to_i16ord: src/scalar/scalar_f16.rs:10-13 -- ((x >> 15) & 0x7FFF) ^ x on i16 */
static inline int16_t f16_to_i16ord(uint16_t bits) {
  int16_t x = (int16_t)bits;
  return (int16_t)((((int16_t)(x >> 15)) & 0x7FFF) ^ x);
}
static inline int isnan_f16(uint16_t bits) { return (bits & 0x7FFFu) > 0x7C00u; }
#define F16_PINF 0x7C00u
#define F16_NINF 0xFC00u

int oracle_argminmax_f16(const uint16_t* arr, uint64_t n, int mode, uint64_t out[2]) {
  if (n == 0) return ORACLE_ERR_EMPTY;
  uint64_t low_index = 0, high_index = 0;
  if (mode == 1) {
    /* scalar_argminmax_f16_return_nan, scalar_f16.rs:20-48 */
    int16_t low = f16_to_i16ord(arr[0]);
    int16_t high = f16_to_i16ord(arr[0]);
    for (uint64_t i = 0; i < n; ++i) {
      uint16_t raw = arr[i];
      if (isnan_f16(raw)) {
        out[0] = i;
        out[1] = i;
        return ORACLE_OK;
      }
      int16_t v = f16_to_i16ord(raw);
      if (v < low) {
        low = v;
        low_index = i;
      } else if (v > high) {
        high = v;
        high_index = i;
      }
    }
  } else {
    /* scalar_argminmax_f16_ignore_nan, scalar_f16.rs:100-136 */
    int16_t low = f16_to_i16ord(F16_PINF);
    int16_t high = f16_to_i16ord(F16_NINF);
    int first_non_nan_update = 1;
    for (uint64_t i = 0; i < n; ++i) {
      uint16_t raw = arr[i];
      if (isnan_f16(raw)) continue;
      int16_t v = f16_to_i16ord(raw);
      if (first_non_nan_update) {
        low = v;
        high = v;
        low_index = i;
        high_index = i;
        first_non_nan_update = 0;
      } else if (v < low) {
        low = v;
        low_index = i;
      } else if (v > high) {
        high = v;
        high_index = i;
      }
    }
  }
  out[0] = low_index;
  out[1] = high_index;
  return ORACLE_OK;
}

/*
 * Single-output f16 variants.  NOTE the reference's own quirk:
 * the ignore-NaN singles (scalar_f16.rs:138-180) have NO first-non-NaN rule, so
 * [NaN, +inf, ...] gives argmin 0 here but 1 through the fused function above.
 * Restated as written; the CUDA path's single-output entry points follow the
 * fused semantics, and the tests that compare singles stay outside that zone.
 */
int oracle_argmin_f16(const uint16_t* arr, uint64_t n, int mode, uint64_t* out) {
  if (n == 0) return ORACLE_ERR_EMPTY;
  uint64_t low_index = 0;
  if (mode == 1) { /* scalar_f16.rs:50-73 */
    int16_t low = f16_to_i16ord(arr[0]);
    for (uint64_t i = 0; i < n; ++i) {
      uint16_t raw = arr[i];
      if (isnan_f16(raw)) {
        *out = i;
        return ORACLE_OK;
      }
      int16_t v = f16_to_i16ord(raw);
      if (v < low) {
        low = v;
        low_index = i;
      }
    }
  } else { /* scalar_f16.rs:138-158 */
    int16_t low = f16_to_i16ord(F16_PINF);
    for (uint64_t i = 0; i < n; ++i) {
      uint16_t raw = arr[i];
      if (isnan_f16(raw)) continue;
      int16_t v = f16_to_i16ord(raw);
      if (v < low) {
        low = v;
        low_index = i;
      }
    }
  }
  *out = low_index;
  return ORACLE_OK;
}

int oracle_argmax_f16(const uint16_t* arr, uint64_t n, int mode, uint64_t* out) {
  if (n == 0) return ORACLE_ERR_EMPTY;
  uint64_t high_index = 0;
  if (mode == 1) { /* scalar_f16.rs:75-98 */
    int16_t high = f16_to_i16ord(arr[0]);
    for (uint64_t i = 0; i < n; ++i) {
      uint16_t raw = arr[i];
      if (isnan_f16(raw)) {
        *out = i;
        return ORACLE_OK;
      }
      int16_t v = f16_to_i16ord(raw);
      if (v > high) {
        high = v;
        high_index = i;
      }
    }
  } else { /* scalar_f16.rs:160-180 */
    int16_t high = f16_to_i16ord(F16_NINF);
    for (uint64_t i = 0; i < n; ++i) {
      uint16_t raw = arr[i];
      if (isnan_f16(raw)) continue;
      int16_t v = f16_to_i16ord(raw);
      if (v > high) {
        high = v;
        high_index = i;
      }
    }
  }
  *out = high_index;
  return ORACLE_OK;
}

/* ---------------------------------------------------------- generic entry --- */
/* dtype codes match include/argminmax_b200.h (amm_dtype). */
int oracle_argminmax(int dtype, const void* arr, uint64_t n, int mode, uint64_t out[2]) {
  switch (dtype) {
    case 0: return oracle_argminmax_i8((const int8_t*)arr, n, mode, out);
    case 1: return oracle_argminmax_i16((const int16_t*)arr, n, mode, out);
    case 2: return oracle_argminmax_i32((const int32_t*)arr, n, mode, out);
    case 3: return oracle_argminmax_i64((const int64_t*)arr, n, mode, out);
    case 4: return oracle_argminmax_u8((const uint8_t*)arr, n, mode, out);
    case 5: return oracle_argminmax_u16((const uint16_t*)arr, n, mode, out);
    case 6: return oracle_argminmax_u32((const uint32_t*)arr, n, mode, out);
    case 7: return oracle_argminmax_u64((const uint64_t*)arr, n, mode, out);
    case 8: return oracle_argminmax_f16((const uint16_t*)arr, n, mode, out);
    case 9: return oracle_argminmax_f32((const float*)arr, n, mode, out);
    case 10: return oracle_argminmax_f64((const double*)arr, n, mode, out);
    default: return 2;
  }
}
int oracle_argmin(int dtype, const void* arr, uint64_t n, int mode, uint64_t* out) {
  switch (dtype) {
    case 0: return oracle_argmin_i8((const int8_t*)arr, n, mode, out);
    case 1: return oracle_argmin_i16((const int16_t*)arr, n, mode, out);
    case 2: return oracle_argmin_i32((const int32_t*)arr, n, mode, out);
    case 3: return oracle_argmin_i64((const int64_t*)arr, n, mode, out);
    case 4: return oracle_argmin_u8((const uint8_t*)arr, n, mode, out);
    case 5: return oracle_argmin_u16((const uint16_t*)arr, n, mode, out);
    case 6: return oracle_argmin_u32((const uint32_t*)arr, n, mode, out);
    case 7: return oracle_argmin_u64((const uint64_t*)arr, n, mode, out);
    case 8: return oracle_argmin_f16((const uint16_t*)arr, n, mode, out);
    case 9: return oracle_argmin_f32((const float*)arr, n, mode, out);
    case 10: return oracle_argmin_f64((const double*)arr, n, mode, out);
    default: return 2;
  }
}
int oracle_argmax(int dtype, const void* arr, uint64_t n, int mode, uint64_t* out) {
  switch (dtype) {
    case 0: return oracle_argmax_i8((const int8_t*)arr, n, mode, out);
    case 1: return oracle_argmax_i16((const int16_t*)arr, n, mode, out);
    case 2: return oracle_argmax_i32((const int32_t*)arr, n, mode, out);
    case 3: return oracle_argmax_i64((const int64_t*)arr, n, mode, out);
    case 4: return oracle_argmax_u8((const uint8_t*)arr, n, mode, out);
    case 5: return oracle_argmax_u16((const uint16_t*)arr, n, mode, out);
    case 6: return oracle_argmax_u32((const uint32_t*)arr, n, mode, out);
    case 7: return oracle_argmax_u64((const uint64_t*)arr, n, mode, out);
    case 8: return oracle_argmax_f16((const uint16_t*)arr, n, mode, out);
    case 9: return oracle_argmax_f32((const float*)arr, n, mode, out);
    case 10: return oracle_argmax_f64((const double*)arr, n, mode, out);
    default: return 2;
  }
}

/* Batch form for the windowed tests (the reference has no windowed entry point: its
 * down-sampling users call `argminmax` once per bin, and so does this loop).  Window w is
 * [offsets[w], offsets[w+1]) when `offsets` is given, else [w*window, min(n,(w+1)*window)).
 * out[2w], out[2w+1] = absolute (argmin, argmax); an empty window yields UINT64_MAX twice. */
static const uint64_t elem_size_of[11] = {1, 2, 4, 8, 1, 2, 4, 8, 2, 4, 8};
int oracle_argminmax_windows(int dtype, const void* arr, uint64_t n, const uint64_t* offsets,
                             uint64_t window, uint64_t n_windows, int mode, uint64_t* out) {
  if (dtype < 0 || dtype > 10) return 2;
  const uint64_t es = elem_size_of[dtype];
  for (uint64_t w = 0; w < n_windows; ++w) {
    uint64_t b = offsets ? offsets[w] : w * window;
    uint64_t e = offsets ? offsets[w + 1] : b + window;
    if (e > n) e = n;
    if (b > e) b = e;
    if (b == e) {
      out[2 * w] = out[2 * w + 1] = UINT64_MAX;
      continue;
    }
    uint64_t r[2];
    const int rc = oracle_argminmax(dtype, (const char*)arr + b * es, e - b, mode, r);
    if (rc) return rc;
    out[2 * w] = r[0] + b;
    out[2 * w + 1] = r[1] + b;
  }
  return 0;
}
