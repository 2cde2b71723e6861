/*
 * oracle/cpu_simd_f32.c -- TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C-intrinsics restatement of the reference's runtime-dispatched SIMD path for
 * `<&[f32] as ArgMinMax>::argminmax` (ignore-NaN), i.e. what a user of the crate
 * gets on an x86-64 host.  It exists (1) as the CPU baseline that bench.py times
 * beside the GPU (`cpu_baseline`, `--impl reference`), because the Rust crate
 * cannot be built in this image (no cargo / crates.io), and (2) as a second,
 * lane-exact opinion for the scalar oracle on ordinary data.
 *
 * Followed, piece by piece (paths relative to the reference root):
 *   dispatch ladder                 src/lib.rs:422-459  (avx512f -> avx -> scalar)
 *   AVX-512 ops (LANE = 16)         src/simd/simd_f32_ignore_nan.rs:213-279
 *   AVX(2) ops  (LANE = 8)          src/simd/simd_f32_ignore_nan.rs:74-140
 *   MAX_INDEX = 1 << 24 (f32 index lanes)   src/simd/simd_f32_ignore_nan.rs:69
 *   lane init for ignore-NaN        src/simd/generic.rs:256-305
 *   hot loop _core_argminmax        src/simd/generic.rs:362-412
 *   horizontal reduce               src/simd/generic.rs:56-87 + src/simd/task.rs:265-296
 *   chunk restart loop              src/simd/generic.rs:487-543
 *   prefix/tail split + merge       src/simd/task.rs:4-55, 120-136, 149-228
 *   scalar tail                     src/scalar/generic.rs:181-217 (FloatIgnoreNaN)
 *
 * Like the reference, this path returns index 0 for inputs whose only non-NaN
 * values are +inf (min) / -inf (max) after a leading NaN (README.md:124-125);
 * that documented corner is outside the strict-parity set.
 *
 * The `_mt` entry point (contiguous shards over pthreads + ordered first-index
 * combine) is NOT something the reference does -- the crate is single-threaded --
 * it is reported only to show the host's memory-bandwidth ceiling.
 */
#define _GNU_SOURCE
#include <immintrin.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int isnan_f32(float v) {
  uint32_t b;
  memcpy(&b, &v, 4);
  return (b & 0x7FFFFFFFu) > 0x7F800000u;
}

typedef struct {
  uint64_t min_index;
  float min_value;
  uint64_t max_index;
  float max_value;
} core_result;

/* SCALAR<FloatIgnoreNaN>::argminmax, src/scalar/generic.rs:181-217 */
static void scalar_f32_ignore(const float* arr, uint64_t n, uint64_t out[2]) {
  uint64_t low_index = 0, high_index = 0;
  float start = arr[0];
  float low = start, high = start;
  int first_non_nan_update = 0;
  if (isnan_f32(start)) {
    low = INFINITY;
    high = -INFINITY;
    first_non_nan_update = 1;
  }
  for (uint64_t i = 0; i < n; ++i) {
    float v = arr[i];
    if (first_non_nan_update) {
      if (!isnan_f32(v)) {
        low = v;
        low_index = i;
        high = v;
        high_index = i;
        first_non_nan_update = 0;
      }
    } else if (v < low) {
      low = v;
      low_index = i;
    } else if (v > high) {
      high = v;
      high_index = i;
    }
  }
  out[0] = low_index;
  out[1] = high_index;
}

/* min_index_value / max_index_value on f32 index + value lanes, task.rs:265-296 */
static void horiz(const float* idx, const float* val, int lanes, int want_max, uint64_t* oi,
                  float* ov) {
  float bi = idx[0], bv = val[0];
  for (int i = 0; i < lanes; ++i) {
    float v = val[i], ix = idx[i];
    int better = want_max ? (v > bv) : (v < bv);
    if (better || (v == bv && ix < bi)) {
      bv = v;
      bi = ix;
    }
  }
  *oi = (uint64_t)bi;
  *ov = bv;
}

/* ---- AVX-512F, LANE = 16 (simd_f32_ignore_nan.rs:213-279; generic.rs:362-412) ---- */
__attribute__((target("avx512f"))) static core_result core_avx512(const float* arr, uint64_t n) {
  const __m512 pinf = _mm512_set1_ps(INFINITY), ninf = _mm512_set1_ps(-INFINITY);
  const __m512 zero = _mm512_setzero_ps();
  const __m512 initial_index = _mm512_setr_ps(0.f, 1.f, 2.f, 3.f, 4.f, 5.f, 6.f, 7.f, 8.f, 9.f, 10.f,
                                              11.f, 12.f, 13.f, 14.f, 15.f);
  const __m512 increment = _mm512_set1_ps(16.f);
  __m512 new_index = initial_index;
  /* _initialize_index_values_low/high, generic.rs:266-305 */
  __m512 nv = _mm512_loadu_ps(arr);
  __mmask16 ml = _mm512_cmp_ps_mask(nv, pinf, _CMP_LT_OQ);
  __m512 values_low = _mm512_mask_blend_ps(ml, pinf, nv);
  __m512 index_low = _mm512_mask_blend_ps(ml, zero, initial_index);
  __mmask16 mh = _mm512_cmp_ps_mask(nv, ninf, _CMP_GT_OQ);
  __m512 values_high = _mm512_mask_blend_ps(mh, ninf, nv);
  __m512 index_high = _mm512_mask_blend_ps(mh, zero, initial_index);
  const float* p = arr;
  for (uint64_t it = 0; it + 1 < n / 16; ++it) {
    new_index = _mm512_add_ps(new_index, increment);
    p += 16;
    nv = _mm512_loadu_ps(p);
    ml = _mm512_cmp_ps_mask(nv, values_low, _CMP_LT_OQ);
    values_low = _mm512_mask_blend_ps(ml, values_low, nv);
    index_low = _mm512_mask_blend_ps(ml, index_low, new_index);
    mh = _mm512_cmp_ps_mask(nv, values_high, _CMP_GT_OQ);
    values_high = _mm512_mask_blend_ps(mh, values_high, nv);
    index_high = _mm512_mask_blend_ps(mh, index_high, new_index);
  }
  float il[16], vl[16], ih[16], vh[16];
  _mm512_storeu_ps(il, index_low);
  _mm512_storeu_ps(vl, values_low);
  _mm512_storeu_ps(ih, index_high);
  _mm512_storeu_ps(vh, values_high);
  core_result r;
  horiz(il, vl, 16, 0, &r.min_index, &r.min_value);
  horiz(ih, vh, 16, 1, &r.max_index, &r.max_value);
  return r;
}

/* ---- AVX, LANE = 8 (simd_f32_ignore_nan.rs:74-140) ---- */
__attribute__((target("avx"))) static core_result core_avx(const float* arr, uint64_t n) {
  const __m256 pinf = _mm256_set1_ps(INFINITY), ninf = _mm256_set1_ps(-INFINITY);
  const __m256 zero = _mm256_setzero_ps();
  const __m256 initial_index = _mm256_setr_ps(0.f, 1.f, 2.f, 3.f, 4.f, 5.f, 6.f, 7.f);
  const __m256 increment = _mm256_set1_ps(8.f);
  __m256 new_index = initial_index;
  __m256 nv = _mm256_loadu_ps(arr);
  __m256 ml = _mm256_cmp_ps(nv, pinf, _CMP_LT_OQ);
  __m256 values_low = _mm256_blendv_ps(pinf, nv, ml);
  __m256 index_low = _mm256_blendv_ps(zero, initial_index, ml);
  __m256 mh = _mm256_cmp_ps(nv, ninf, _CMP_GT_OQ);
  __m256 values_high = _mm256_blendv_ps(ninf, nv, mh);
  __m256 index_high = _mm256_blendv_ps(zero, initial_index, mh);
  const float* p = arr;
  for (uint64_t it = 0; it + 1 < n / 8; ++it) {
    new_index = _mm256_add_ps(new_index, increment);
    p += 8;
    nv = _mm256_loadu_ps(p);
    ml = _mm256_cmp_ps(nv, values_low, _CMP_LT_OQ);
    values_low = _mm256_blendv_ps(values_low, nv, ml);
    index_low = _mm256_blendv_ps(index_low, new_index, ml);
    mh = _mm256_cmp_ps(nv, values_high, _CMP_GT_OQ);
    values_high = _mm256_blendv_ps(values_high, nv, mh);
    index_high = _mm256_blendv_ps(index_high, new_index, mh);
  }
  float il[8], vl[8], ih[8], vh[8];
  _mm256_storeu_ps(il, index_low);
  _mm256_storeu_ps(vl, values_low);
  _mm256_storeu_ps(ih, index_high);
  _mm256_storeu_ps(vh, values_high);
  core_result r;
  horiz(il, vl, 8, 0, &r.min_index, &r.min_value);
  horiz(ih, vh, 8, 1, &r.max_index, &r.max_value);
  return r;
}

typedef core_result (*core_fn)(const float*, uint64_t);

/* _overflow_safe_core_argminmax, generic.rs:487-543 (IgnoreNaN: _return_check == false) */
static core_result overflow_safe(const float* arr, uint64_t n, int lanes, core_fn core) {
  const uint64_t max_index = 1ull << 24; /* simd_f32_ignore_nan.rs:69 */
  const uint64_t dtype_max = max_index - max_index % (uint64_t)lanes;
  const uint64_t n_loops = n / dtype_max;
  core_result acc = {0, INFINITY, 0, -INFINITY};
  uint64_t start = 0;
  for (uint64_t l = 0; l < n_loops; ++l) {
    core_result c = core(arr + start, dtype_max);
    if (c.min_value < acc.min_value) {
      acc.min_index = start + c.min_index;
      acc.min_value = c.min_value;
    }
    if (c.max_value > acc.max_value) {
      acc.max_index = start + c.max_index;
      acc.max_value = c.max_value;
    }
    start += dtype_max;
  }
  if (start < n) {
    core_result c = core(arr + start, n - start);
    if (c.min_value < acc.min_value) {
      acc.min_index = start + c.min_index;
      acc.min_value = c.min_value;
    }
    if (c.max_value > acc.max_value) {
      acc.max_index = start + c.max_index;
      acc.max_value = c.max_value;
    }
  }
  return acc;
}

/* argminmax_generic, task.rs:4-55 with ignore_nan = true */
static void generic_f32_ignore(const float* arr, uint64_t n, int lanes, core_fn core,
                               uint64_t out[2]) {
  const uint64_t n_simd = n - n % (uint64_t)lanes;
  const uint64_t n_rem = n - n_simd;
  if (n_simd == 0) {
    scalar_f32_ignore(arr, n, out);
    return;
  }
  core_result s = overflow_safe(arr, n_simd, lanes, core);
  if (n_rem == 0) {
    out[0] = s.min_index;
    out[1] = s.max_index;
    return;
  }
  uint64_t rem[2];
  scalar_f32_ignore(arr + n_simd, n_rem, rem);
  const float rmin = arr[n_simd + rem[0]], rmax = arr[n_simd + rem[1]];
  /* find_final_index_min, task.rs:149-182: Less|Equal -> simd, Greater -> rem,
     unordered -> rem unless rem is the NaN (simd lanes are never NaN here) */
  if (s.min_value <= rmin)
    out[0] = s.min_index;
  else if (s.min_value > rmin)
    out[0] = n_simd + rem[0];
  else
    out[0] = isnan_f32(rmin) ? s.min_index : n_simd + rem[0];
  if (s.max_value >= rmax)
    out[1] = s.max_index;
  else if (s.max_value < rmax)
    out[1] = n_simd + rem[1];
  else
    out[1] = isnan_f32(rmax) ? s.max_index : n_simd + rem[1];
}

/* which ISA the reference's ladder (src/lib.rs:427-443) would pick on this host:
   2 = avx512f, 1 = avx, 0 = scalar fallback */
int cpu_simd_isa(void) {
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512f")) return 2;
  if (__builtin_cpu_supports("avx")) return 1;
  return 0;
}

/* force_isa < 0: runtime dispatch like the reference; otherwise 0/1/2 as above */
int cpu_simd_argminmax_f32_ignore(const float* arr, uint64_t n, int force_isa, uint64_t out[2]) {
  if (n == 0) return 1;
  int isa = force_isa < 0 ? cpu_simd_isa() : force_isa;
  if (isa > cpu_simd_isa()) return 2;
  if (isa == 2)
    generic_f32_ignore(arr, n, 16, core_avx512, out);
  else if (isa == 1)
    generic_f32_ignore(arr, n, 8, core_avx, out);
  else
    scalar_f32_ignore(arr, n, out);
  return 0;
}

/* --------------------------------------------------------------------------
 * NOT THE REFERENCE: all-cores variant (contiguous shards, ordered combine).
 * -------------------------------------------------------------------------- */
typedef struct {
  const float* arr;
  uint64_t begin, len;
  int isa;
  uint64_t out[2];
  int cpu;
} mt_job;

static void* mt_worker(void* p) {
  mt_job* j = (mt_job*)p;
  if (j->cpu >= 0) {
    cpu_set_t set;
    CPU_ZERO(&set);
    CPU_SET(j->cpu, &set);
    pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
  }
  if (j->len) cpu_simd_argminmax_f32_ignore(j->arr + j->begin, j->len, j->isa, j->out);
  return NULL;
}

int cpu_simd_argminmax_f32_ignore_mt(const float* arr, uint64_t n, int nthreads, uint64_t out[2]) {
  if (n == 0) return 1;
  if (nthreads < 1) nthreads = 1;
  if ((uint64_t)nthreads > n) nthreads = (int)n;
  mt_job* jobs = (mt_job*)calloc((size_t)nthreads, sizeof(mt_job));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  const int isa = cpu_simd_isa();
  uint64_t per = (n + (uint64_t)nthreads - 1) / (uint64_t)nthreads;
  per = (per + 15) & ~15ull;
  for (int t = 0; t < nthreads; ++t) {
    uint64_t b = per * (uint64_t)t;
    if (b > n) b = n;
    uint64_t e = b + per;
    if (e > n) e = n;
    jobs[t].arr = arr;
    jobs[t].begin = b;
    jobs[t].len = e - b;
    jobs[t].isa = isa;
    jobs[t].cpu = -1;
    pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
  }
  int have = 0;
  float bmin = 0.f, bmax = 0.f;
  uint64_t imin = 0, imax = 0;
  for (int t = 0; t < nthreads; ++t) {
    pthread_join(th[t], NULL);
    if (!jobs[t].len) continue;
    const uint64_t a = jobs[t].begin + jobs[t].out[0], b = jobs[t].begin + jobs[t].out[1];
    const float va = arr[a], vb = arr[b];
    /* ordered combine: strict compares, so the earlier shard wins ties; a shard
       whose value is NaN (an all-NaN shard) never displaces a real value */
    if (!have) {
      bmin = va;
      imin = a;
      bmax = vb;
      imax = b;
    } else {
      if (isnan_f32(bmin) ? !isnan_f32(va) : (va < bmin)) {
        bmin = va;
        imin = a;
      }
      if (isnan_f32(bmax) ? !isnan_f32(vb) : (vb > bmax)) {
        bmax = vb;
        imax = b;
      }
    }
    have = 1;
  }
  /* all-NaN input: scalar semantics say (0,0) */
  if (isnan_f32(bmin)) imin = 0;
  if (isnan_f32(bmax)) imax = 0;
  out[0] = imin;
  out[1] = imax;
  free(jobs);
  free(th);
  return 0;
}
