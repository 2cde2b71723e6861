/*
 * oracle/simd_emul.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Lane-exact scalar EMULATION of the reference's dispatched SIMD path for all eleven dtypes
 * and both NaN strategies: what `slice.argminmax()` / `slice.nanargminmax()` compute on an
 * x86-64 host through src/lib.rs:272-664, lane by lane, chunk by chunk.  It exists to
 * cross-check the scalar oracle (argminmax_oracle.c) and the CUDA path against the reference's
 * SIMD algorithm, including the places where that algorithm is position/lane dependent
 * (SURVEY.md App. A.3).  It performs no intrinsics: every lane is a C variable.
 *
 * Followed (paths relative to the reference root):
 *   split prefix/tail + merge + NaN fix-up   src/simd/task.rs:4-55, 120-136, 149-260
 *   chunk restart loop                       src/simd/generic.rs:487-543, limit :91-93
 *   hot loop (per-lane value/index blends)   src/simd/generic.rs:362-412
 *   lane init: Int / ReturnNaN               src/simd/generic.rs:138-154
 *              IgnoreNaN                     src/simd/generic.rs:266-305
 *   horizontal reduce rule                   src/simd/task.rs:265-296 (and the in-register
 *                                            8/16-bit versions, e.g. simd_i8.rs:227, which
 *                                            implement the same rule)
 *   body domain per dtype:   ints native / u* xor-signed (simd_u32.rs:56);
 *     f32,f64 return-NaN  -> ord ints (simd_f32_return_nan.rs:324-329)
 *     f32,f64 ignore-NaN  -> float lanes, _CMP_LT_OQ/_GT_OQ (simd_f32_ignore_nan.rs:246-253)
 *     f16 return-NaN      -> i16 ord (simd_f16_return_nan.rs)
 *     f16 ignore-NaN      -> i16 ord, compares ANDed with a non-NaN test applied to the
 *                            ALREADY TRANSFORMED register (simd_f16_ignore_nan.rs:459-463,
 *                            501-508) -- reproduced as written, bug included.
 *   lanes / MAX_INDEX: App. C table (isa 0 = AVX-512 host as dispatched, isa 1 = AVX2 host).
 *
 * Values of the element type ("T domain": chunk merges generic.rs:513-520, prefix/tail merge
 * task.rs:149-228) are compared as the reference compares ScalarDType: ints natively, floats
 * with IEEE partial order (f16 through its exact widening to double).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int oracle_argminmax(int dtype, const void* arr, uint64_t n, int mode, uint64_t out[2]);

enum { DT_I8, DT_I16, DT_I32, DT_I64, DT_U8, DT_U16, DT_U32, DT_U64, DT_F16, DT_F32, DT_F64 };
enum { BODY_INT = 0, BODY_FLOAT = 1, BODY_F16_IGNORE = 2 };

static size_t dt_size(int dt) {
  switch (dt) {
    case DT_I8: case DT_U8: return 1;
    case DT_I16: case DT_U16: case DT_F16: return 2;
    case DT_I32: case DT_U32: case DT_F32: return 4;
    default: return 8;
  }
}

static double f16_to_double(uint16_t h) {
  const int s = h >> 15, e = (h >> 10) & 0x1F, m = h & 0x3FF;
  double v;
  if (e == 0) v = ldexp((double)m, -24);
  else if (e == 31) v = m ? NAN : INFINITY;
  else v = ldexp((double)(m | 0x400), e - 25);
  return s ? -v : v;
}

/* One element, three views: raw bits, T-domain double (floats only), body-domain key. */
typedef struct {
  int64_t ikey; /* ints: order-preserving signed key; floats: ord key of the raw bits */
  double fval;  /* floats: the value (exact); ints: unused */
  int is_nan;
} elem_t;

static elem_t load_elem(int dt, const void* arr, uint64_t i) {
  elem_t e;
  e.fval = 0.0;
  e.is_nan = 0;
  switch (dt) {
    case DT_I8: e.ikey = ((const int8_t*)arr)[i]; break;
    case DT_I16: e.ikey = ((const int16_t*)arr)[i]; break;
    case DT_I32: e.ikey = ((const int32_t*)arr)[i]; break;
    case DT_I64: e.ikey = ((const int64_t*)arr)[i]; break;
    case DT_U8: e.ikey = ((const uint8_t*)arr)[i]; break;
    case DT_U16: e.ikey = ((const uint16_t*)arr)[i]; break;
    case DT_U32: e.ikey = ((const uint32_t*)arr)[i]; break;
    case DT_U64: e.ikey = (int64_t)(((const uint64_t*)arr)[i] ^ 0x8000000000000000ull); break;
    case DT_F16: {
      const uint16_t b = ((const uint16_t*)arr)[i];
      const int16_t x = (int16_t)b;
      e.ikey = (int16_t)((((int16_t)(x >> 15)) & 0x7FFF) ^ x);
      e.fval = f16_to_double(b);
      e.is_nan = (b & 0x7FFF) > 0x7C00;
      break;
    }
    case DT_F32: {
      uint32_t b;
      memcpy(&b, (const char*)arr + 4 * i, 4);
      const int32_t x = (int32_t)b;
      e.ikey = (int32_t)(((x >> 31) & 0x7FFFFFFF) ^ x);
      float f;
      memcpy(&f, &b, 4);
      e.fval = f;
      e.is_nan = (b & 0x7FFFFFFFu) > 0x7F800000u;
      break;
    }
    default: {
      uint64_t b;
      memcpy(&b, (const char*)arr + 8 * i, 8);
      const int64_t x = (int64_t)b;
      e.ikey = ((x >> 63) & 0x7FFFFFFFFFFFFFFFll) ^ x;
      memcpy(&e.fval, &b, 8);
      e.is_nan = (b & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull;
    }
  }
  return e;
}

typedef struct {
  int dt, mode /* 0 ignore/int, 1 return */, is_float, body, lanes;
  uint64_t max_index;
  const void* arr;
} cfg_t;

/* ---- T-domain compares (ScalarDType PartialOrd) ------------------------------------- */
static int t_lt(const cfg_t* c, const elem_t* a, const elem_t* b) {
  return c->is_float ? (a->fval < b->fval) : (a->ikey < b->ikey);
}
static int t_gt(const cfg_t* c, const elem_t* a, const elem_t* b) {
  return c->is_float ? (a->fval > b->fval) : (a->ikey > b->ikey);
}
static int t_eq(const cfg_t* c, const elem_t* a, const elem_t* b) {
  return c->is_float ? (a->fval == b->fval) : (a->ikey == b->ikey);
}

/* ---- body-domain compares (the SIMD registers) --------------------------------------- */
static int f16_body_nonnan(int64_t ord) { return (ord & 0x7FFF) < 0x7C01; } /* on the ORD value */
static int b_lt(const cfg_t* c, const elem_t* a, const elem_t* b) {
  if (c->body == BODY_FLOAT) return a->fval < b->fval; /* _CMP_LT_OQ */
  if (c->body == BODY_F16_IGNORE) return (a->ikey < b->ikey) && f16_body_nonnan(a->ikey);
  return a->ikey < b->ikey;
}
static int b_gt(const cfg_t* c, const elem_t* a, const elem_t* b) {
  if (c->body == BODY_FLOAT) return a->fval > b->fval; /* _CMP_GT_OQ */
  if (c->body == BODY_F16_IGNORE) return (a->ikey > b->ikey) && f16_body_nonnan(a->ikey);
  return a->ikey > b->ikey;
}
/* horizontal reduce compares the lane VALUES as the register holds them */
static int h_lt(const cfg_t* c, const elem_t* a, const elem_t* b) {
  return c->body == BODY_FLOAT ? (a->fval < b->fval) : (a->ikey < b->ikey);
}
static int h_gt(const cfg_t* c, const elem_t* a, const elem_t* b) {
  return c->body == BODY_FLOAT ? (a->fval > b->fval) : (a->ikey > b->ikey);
}
static int h_eq(const cfg_t* c, const elem_t* a, const elem_t* b) {
  return c->body == BODY_FLOAT ? (a->fval == b->fval) : (a->ikey == b->ikey);
}

static elem_t const_elem(const cfg_t* c, double v) { /* +-inf as a lane value (ignore-NaN init) */
  elem_t e;
  e.fval = v;
  e.is_nan = 0;
  if (c->dt == DT_F16) {
    const int16_t x = (int16_t)(v > 0 ? 0x7C00 : 0xFC00);
    e.ikey = (int16_t)((((int16_t)(x >> 15)) & 0x7FFF) ^ x);
  } else if (c->dt == DT_F32) {
    const int32_t x = (int32_t)(v > 0 ? 0x7F800000u : 0xFF800000u);
    e.ikey = (int32_t)(((x >> 31) & 0x7FFFFFFF) ^ x);
  } else {
    const int64_t x = (int64_t)(v > 0 ? 0x7FF0000000000000ull : 0xFFF0000000000000ull);
    e.ikey = ((x >> 63) & 0x7FFFFFFFFFFFFFFFll) ^ x;
  }
  return e;
}

typedef struct {
  uint64_t min_index, max_index;
  elem_t min_value, max_value;
} core_t;

#define MAX_LANES 64

/* _core_argminmax, src/simd/generic.rs:362-412 on arr[start .. start+len), len % lanes == 0 */
static core_t core_argminmax(const cfg_t* c, uint64_t start, uint64_t len) {
  const int L = c->lanes;
  elem_t vlow[MAX_LANES], vhigh[MAX_LANES];
  uint64_t ilow[MAX_LANES], ihigh[MAX_LANES];
  const int ignore = c->is_float && c->mode == 0;
  const elem_t pinf = const_elem(c, INFINITY), ninf = const_elem(c, -INFINITY);
  for (int l = 0; l < L; ++l) {
    const elem_t nv = load_elem(c->dt, c->arr, start + l);
    if (ignore) { /* generic.rs:266-305 */
      const int ml = b_lt(c, &nv, &pinf);
      vlow[l] = ml ? nv : pinf;
      ilow[l] = ml ? (uint64_t)l : 0;
      const int mh = b_gt(c, &nv, &ninf);
      vhigh[l] = mh ? nv : ninf;
      ihigh[l] = mh ? (uint64_t)l : 0;
    } else { /* generic.rs:138-154 */
      vlow[l] = vhigh[l] = nv;
      ilow[l] = ihigh[l] = (uint64_t)l;
    }
  }
  for (uint64_t v = 1; v < len / (uint64_t)L; ++v) {
    for (int l = 0; l < L; ++l) {
      const uint64_t idx = v * (uint64_t)L + (uint64_t)l; /* new_index lane */
      const elem_t nv = load_elem(c->dt, c->arr, start + idx);
      if (b_lt(c, &nv, &vlow[l])) {
        vlow[l] = nv;
        ilow[l] = idx;
      }
      if (b_gt(c, &nv, &vhigh[l])) {
        vhigh[l] = nv;
        ihigh[l] = idx;
      }
    }
  }
  /* min_index_value / max_index_value, task.rs:265-296 */
  core_t r;
  r.min_index = ilow[0];
  r.min_value = vlow[0];
  r.max_index = ihigh[0];
  r.max_value = vhigh[0];
  for (int l = 0; l < L; ++l) {
    if (h_lt(c, &vlow[l], &r.min_value) || (h_eq(c, &vlow[l], &r.min_value) && ilow[l] < r.min_index)) {
      r.min_value = vlow[l];
      r.min_index = ilow[l];
    }
    if (h_gt(c, &vhigh[l], &r.max_value) || (h_eq(c, &vhigh[l], &r.max_value) && ihigh[l] < r.max_index)) {
      r.max_value = vhigh[l];
      r.max_index = ihigh[l];
    }
  }
  /* the horizontal result is handed back as a ScalarDType value (ord -> float where needed):
     is_nan must describe THAT value */
  return r;
}

static int return_check(const cfg_t* c, const elem_t* v) {
  return c->is_float && c->mode == 1 && v->is_nan; /* generic.rs:217-236 */
}

/* _overflow_safe_core_argminmax, src/simd/generic.rs:487-543 */
static core_t overflow_safe(const cfg_t* c, uint64_t len) {
  const uint64_t dtype_max = c->max_index - c->max_index % (uint64_t)c->lanes;
  const uint64_t n_loops = len / dtype_max;
  core_t acc;
  acc.min_index = acc.max_index = 0;
  if (c->is_float && c->mode == 0) {
    acc.min_value = const_elem(c, INFINITY);
    acc.max_value = const_elem(c, -INFINITY);
  } else {
    acc.min_value = acc.max_value = load_elem(c->dt, c->arr, 0);
  }
  uint64_t start = 0;
  for (uint64_t l = 0; l <= n_loops; ++l) {
    uint64_t this_len = dtype_max;
    if (l == n_loops) {
      if (start >= len) break;
      this_len = len - start;
    }
    if (return_check(c, &acc.min_value) || return_check(c, &acc.max_value)) return acc;
    const core_t r = core_argminmax(c, start, this_len);
    if (t_lt(c, &r.min_value, &acc.min_value) || return_check(c, &r.min_value)) {
      acc.min_index = start + r.min_index;
      acc.min_value = r.min_value;
    }
    if (t_gt(c, &r.max_value, &acc.max_value) || return_check(c, &r.max_value)) {
      acc.max_index = start + r.max_index;
      acc.max_value = r.max_value;
    }
    start += this_len;
  }
  return acc;
}

/* find_final_index_min / _max, src/simd/task.rs:149-228.  which: 0 = min, 1 = max */
static void find_final(const cfg_t* c, int which, uint64_t si, const elem_t* sv, uint64_t ri,
                       const elem_t* rv, uint64_t* oi, elem_t* ov) {
  const int ignore_nan = !(c->is_float && c->mode == 1);
  const int s_better = which == 0 ? t_lt(c, sv, rv) : t_gt(c, sv, rv);
  const int r_better = which == 0 ? t_gt(c, sv, rv) : t_lt(c, sv, rv);
  int take_simd;
  if (s_better || t_eq(c, sv, rv)) take_simd = 1;
  else if (r_better) take_simd = 0;
  else if (!ignore_nan) take_simd = sv->is_nan ? 1 : 0; /* prefer simd if it is the NaN */
  else take_simd = rv->is_nan ? 1 : 0; /* both NaN would panic in the reference */
  *oi = take_simd ? si : ri;
  *ov = take_simd ? *sv : *rv;
}

int simd_emul_argminmax(int dtype, const void* arr, uint64_t n, int mode, int isa, uint64_t out[2]) {
  if (n == 0) return 1;
  if (dtype < 0 || dtype > DT_F64 || (isa != 0 && isa != 1)) return 2;
  cfg_t c;
  c.dt = dtype;
  c.is_float = dtype >= DT_F16;
  c.mode = c.is_float ? (mode == 1) : 0;
  c.arr = arr;
  c.body = BODY_INT;
  const int half = (isa == 1) ? 2 : 1; /* AVX2: half the AVX-512 lanes (8-bit stays SSE) */
  switch (dtype) { /* SURVEY.md App. C */
    case DT_I8: case DT_U8: c.lanes = 16; c.max_index = 127; break;
    case DT_I16: case DT_U16: c.lanes = 32 / half; c.max_index = 32767; break;
    case DT_I32: case DT_U32: c.lanes = 16 / half; c.max_index = 2147483647ull; break;
    case DT_I64: case DT_U64: c.lanes = 8 / half; c.max_index = 9223372036854775807ull; break;
    case DT_F16:
      c.lanes = 32 / half;
      c.max_index = 32767;
      if (c.mode == 0) c.body = BODY_F16_IGNORE;
      break;
    case DT_F32:
      c.lanes = 16 / half;
      if (c.mode == 0) { c.body = BODY_FLOAT; c.max_index = 1ull << 24; }
      else c.max_index = 2147483647ull;
      break;
    default:
      c.lanes = 8 / half;
      if (c.mode == 0) { c.body = BODY_FLOAT; c.max_index = 1ull << 53; }
      else c.max_index = 9223372036854775807ull;
  }
  /* split_array, task.rs:120-136 */
  const uint64_t n_simd = n - n % (uint64_t)c.lanes, n_rem = n - n_simd;
  if (n_simd == 0) return oracle_argminmax(dtype, arr, n, c.mode, out);
  const core_t s = overflow_safe(&c, n_simd);
  uint64_t imin = s.min_index, imax = s.max_index;
  elem_t vmin = s.min_value, vmax = s.max_value;
  if (n_rem) {
    uint64_t rem[2];
    oracle_argminmax(dtype, (const char*)arr + n_simd * dt_size(dtype), n_rem, c.mode, rem);
    const elem_t rmin = load_elem(dtype, arr, n_simd + rem[0]);
    const elem_t rmax = load_elem(dtype, arr, n_simd + rem[1]);
    find_final(&c, 0, s.min_index, &s.min_value, n_simd + rem[0], &rmin, &imin, &vmin);
    find_final(&c, 1, s.max_index, &s.max_value, n_simd + rem[1], &rmax, &imax, &vmax);
  }
  /* get_correct_argminmax_result, task.rs:236-260 */
  if (c.is_float && c.mode == 1 && (vmin.is_nan || vmax.is_nan)) {
    uint64_t k;
    if (vmin.is_nan && vmax.is_nan) k = imin < imax ? imin : imax;
    else if (vmin.is_nan) k = imin;
    else k = imax;
    out[0] = out[1] = k;
    return 0;
  }
  out[0] = imin;
  out[1] = imax;
  return 0;
}
