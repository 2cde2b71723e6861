// libamm.cu -- C ABI (include/argminmax_b200.h) over the sm_100a scan kernel:
// workspace lifecycle, dtype/mode/variant dispatch, launch geometry, the pinned
// result slot, the host-slice pipeline and the record combine.
//
// Stands where the reference's trait impls + run-time ISA dispatch stand
// (src/lib.rs:246-664): there the ladder picks SSE/AVX2/AVX-512/NEON, here there is
// exactly one backend (sm_100a) and NO CPU fallback -- every failure is a status.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "../../include/argminmax_b200.h"
#include "amm_internal.h"
#include "scan_kernel.cuh"

namespace amm {

thread_local int g_last_cuda_error = 0;

// Load shapes compiled for every policy.  {VEC bytes, U vectors per thread per tile,
// software prefetch}.  Tile = threads * U * VEC bytes.
struct VariantDesc {
  int vec, u, prefetch;
  const char* name;
};
static const VariantDesc kVariants[] = {
    {16, 4, 1, "v16x4_pf"},  // 0: 128-bit loads, 64 B/thread/tile, 2 tiles in flight
    {32, 2, 1, "v32x2_pf"},  // 1: 256-bit loads, 64 B/thread/tile, 2 tiles in flight
    {32, 4, 1, "v32x4_pf"},  // 2: 256-bit loads, 128 B/thread/tile, 2 tiles in flight
    {16, 8, 0, "v16x8"},     // 3: 128-bit loads, 128 B/thread/tile, no prefetch
    {32, 4, 0, "v32x4"},     // 4: 256-bit loads, 128 B/thread/tile, no prefetch
    {16, 2, 1, "v16x2_pf"},  // 5: small tiles (latency-oriented)
};
constexpr int kNumVariants = (int)(sizeof(kVariants) / sizeof(kVariants[0]));

template <class P, int VEC, int U, bool PF>
static int launch_one(amm_workspace* ws, ScanParams& p, int dtype_slot, int variant,
                      cudaStream_t stream) {
  auto kernel = argminmax_scan_kernel<P, VEC, U, PF>;
  int& occ = ws->occupancy[dtype_slot][variant];
  if (occ == 0) {
    int o = 0;
    AMM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kernel, ws->threads, 0));
    occ = o < 1 ? 1 : o;
  }
  const uint64_t vec_elems = VEC / sizeof(typename P::Elem);
  const uint64_t tile_vecs = (uint64_t)ws->threads * U;
  // geometry: unaligned head | whole vectors | sub-vector tail
  const uint64_t addr = (uint64_t)(uintptr_t)p.base;
  uint64_t head = ((VEC - (addr % VEC)) % VEC) / sizeof(typename P::Elem);
  if (head > p.n) head = p.n;
  p.head_elems = head;
  p.n_vec = (p.n - head) / vec_elems;
  p.tail_start = head + p.n_vec * vec_elems;
  const uint64_t full = p.n_vec / tile_vecs;
  if (full >= 0xFFFFFFFEull) return AMM_ERR_TOO_LARGE;
  p.n_full_tiles = (uint32_t)full;
  p.rem_vecs = (uint32_t)(p.n_vec % tile_vecs);
  const uint64_t tiles = full + (p.rem_vecs ? 1 : 0);
  int per_sm = ws->ctas_per_sm < occ ? ws->ctas_per_sm : occ;
  uint64_t grid = (uint64_t)ws->sm_count * (uint64_t)per_sm;
  if (grid > tiles) grid = tiles;
  if (grid < 1) grid = 1;
  if (grid > (uint64_t)ws->max_grid) grid = (uint64_t)ws->max_grid;
  kernel<<<(unsigned)grid, ws->threads, 0, stream>>>(p);
  AMM_CUDA(cudaGetLastError());
  ws->last_grid = (int)grid;
  ws->last_threads = ws->threads;
  ws->launches += 1;
  return AMM_OK;
}

template <class P>
static int launch_variant(amm_workspace* ws, ScanParams& p, int dtype_slot, cudaStream_t stream) {
  switch (ws->variant) {
    case 0: return launch_one<P, 16, 4, true>(ws, p, dtype_slot, 0, stream);
    case 1: return launch_one<P, 32, 2, true>(ws, p, dtype_slot, 1, stream);
    case 2: return launch_one<P, 32, 4, true>(ws, p, dtype_slot, 2, stream);
    case 3: return launch_one<P, 16, 8, false>(ws, p, dtype_slot, 3, stream);
    case 4: return launch_one<P, 32, 4, false>(ws, p, dtype_slot, 4, stream);
    case 5: return launch_one<P, 16, 2, true>(ws, p, dtype_slot, 5, stream);
    default: return AMM_ERR_ARG;
  }
}

// The B200 counterpart of the reference's strategy dispatch
// (Int / FloatIgnoreNaN / FloatReturnNaN, src/dtype_strategy.rs + src/lib.rs:272-664).
static int dispatch(amm_workspace* ws, int dtype, int mode, ScanParams& p, cudaStream_t stream) {
  const bool ret = (mode == AMM_RETURN_NAN);
  switch (dtype) {
    case AMM_I8: return launch_variant<PolInt8<true>>(ws, p, 0, stream);
    case AMM_I16: return launch_variant<PolInt16<true>>(ws, p, 1, stream);
    case AMM_I32: return launch_variant<PolInt32<true>>(ws, p, 2, stream);
    case AMM_I64: return launch_variant<PolInt64<true>>(ws, p, 3, stream);
    case AMM_U8: return launch_variant<PolInt8<false>>(ws, p, 4, stream);
    case AMM_U16: return launch_variant<PolInt16<false>>(ws, p, 5, stream);
    case AMM_U32: return launch_variant<PolInt32<false>>(ws, p, 6, stream);
    case AMM_U64: return launch_variant<PolInt64<false>>(ws, p, 7, stream);
    case AMM_F16:
      return ret ? launch_variant<PolF16<MODE_RETURN>>(ws, p, 9, stream)
                 : launch_variant<PolF16<MODE_IGNORE>>(ws, p, 8, stream);
    case AMM_F32:
      return ret ? launch_variant<PolF32<MODE_RETURN>>(ws, p, 11, stream)
                 : launch_variant<PolF32<MODE_IGNORE>>(ws, p, 10, stream);
    case AMM_F64:
      return ret ? launch_variant<PolF64<MODE_RETURN>>(ws, p, 13, stream)
                 : launch_variant<PolF64<MODE_IGNORE>>(ws, p, 12, stream);
    default: return AMM_ERR_ARG;
  }
}

int set_device(int device) {
  int cur = -1;
  AMM_CUDA(cudaGetDevice(&cur));
  if (cur != device) AMM_CUDA(cudaSetDevice(device));
  return AMM_OK;
}

int launch(int dtype, const void* dev_ptr, uint64_t n, int mode, uint64_t global_offset,
           amm_workspace* ws, cudaStream_t stream, void* dev_record_out) {
  if (ws == nullptr || dev_ptr == nullptr) return AMM_ERR_ARG;
  if (dtype < 0 || dtype >= AMM_DTYPE_COUNT) return AMM_ERR_ARG;
  if (mode != AMM_IGNORE_NAN && mode != AMM_RETURN_NAN) return AMM_ERR_ARG;
  if (n == 0) return AMM_ERR_EMPTY;
  const size_t es = amm_dtype_size(dtype);
  if (((uintptr_t)dev_ptr) % es != 0) return AMM_ERR_ALIGN;
  int rc = set_device(ws->device);
  if (rc) return rc;
  ScanParams p;
  memset(&p, 0, sizeof(p));
  p.base = (const uint8_t*)dev_ptr;
  p.n = n;
  p.global_offset = global_offset;
  p.partials = ws->partials;
  p.ticket = ws->ticket;
  p.rec_dev = dev_record_out ? (amm_record*)dev_record_out : ws->rec_dev;
  p.rec_host = ws->rec_host_dev;
  p.seq = ++ws->seq;
  p.mode = mode;
  return dispatch(ws, dtype, mode, p, stream);
}

// Fold records (ascending shard order) into one record.  Strict compares on the key,
// index as tie-break: the lowest shard wins ties, which is the global first
// occurrence (cf. the reference's chunk merge, src/simd/generic.rs:513-520).
amm_record fold_records(const amm_record* recs, int count) {
  bool have = false;
  amm_record o;
  memset(&o, 0, sizeof(o));
  o.min_idx = o.max_idx = o.nan_idx = AMM_NO_INDEX;
  for (int g = 0; g < count; ++g) {
    const amm_record& r = recs[g];
    o.n += r.n;
    if (r.seq > o.seq) o.seq = r.seq;
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
  o.flags = (have ? AMM_REC_HAS_VALUE : 0ull) | (o.nan_idx != AMM_NO_INDEX ? AMM_REC_HAS_NAN : 0ull);
  return o;
}

void combine(const amm_record* recs, int count, int mode, uint64_t out[2]) {
  const amm_record o = fold_records(recs, count);
  if (mode == AMM_RETURN_NAN && (o.flags & AMM_REC_HAS_NAN)) {
    out[0] = out[1] = o.nan_idx;  // first NaN for both (src/scalar/generic.rs:193-196)
  } else if (!(o.flags & AMM_REC_HAS_VALUE)) {
    out[0] = out[1] = 0;  // all NaN in ignore mode (src/simd/test_utils.rs:508-527)
  } else {
    out[0] = o.min_idx;
    out[1] = o.max_idx;
  }
}

// Device-side twin of combine(): one thread folds `count` records (ascending shard
// order, strict compares) into one record whose indices are the final answer, and
// publishes it to the workspace's device + pinned host slots.  Lets a multi-GPU step
// (scan -> all-gather -> combine) stay fully asynchronous on one stream.
__global__ void combine_records_kernel(const amm_record* recs, int count, int mode,
                                       amm_record* rec_dev, amm_record* rec_host, uint64_t seq) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  bool have = false;
  int64_t kmin = 0, kmax = 0;
  uint64_t imin = 0, imax = 0, inan = AMM_NO_INDEX, n = 0;
  for (int g = 0; g < count; ++g) {
    const amm_record r = recs[g];
    n += r.n;
    if ((r.flags & AMM_REC_HAS_NAN) && r.nan_idx < inan) inan = r.nan_idx;
    if (!(r.flags & AMM_REC_HAS_VALUE)) continue;
    if (!have || r.min_key < kmin || (r.min_key == kmin && r.min_idx < imin)) {
      kmin = r.min_key;
      imin = r.min_idx;
    }
    if (!have || r.max_key > kmax || (r.max_key == kmax && r.max_idx < imax)) {
      kmax = r.max_key;
      imax = r.max_idx;
    }
    have = true;
  }
  amm_record o;
  o.min_key = kmin;
  o.max_key = kmax;
  o.min_idx = have ? imin : AMM_NO_INDEX;
  o.max_idx = have ? imax : AMM_NO_INDEX;
  o.nan_idx = inan;
  o.flags = (have ? AMM_REC_HAS_VALUE : 0ull) | (inan != AMM_NO_INDEX ? AMM_REC_HAS_NAN : 0ull);
  o.n = n;
  o.seq = seq;
  (void)mode;
  *rec_dev = o;
  if (rec_host != nullptr) {
    volatile amm_record* h = rec_host;
    h->min_key = o.min_key;
    h->min_idx = o.min_idx;
    h->max_key = o.max_key;
    h->max_idx = o.max_idx;
    h->nan_idx = o.nan_idx;
    h->flags = o.flags;
    h->n = o.n;
    __threadfence_system();
    h->seq = o.seq;
  }
}

}  // namespace amm

using namespace amm;

extern "C" {

int amm_combine_records_launch(const void* dev_records, int count, int mode, amm_workspace* ws,
                               void* stream) {
  if (!dev_records || !ws || count < 1) return AMM_ERR_ARG;
  if (mode != AMM_IGNORE_NAN && mode != AMM_RETURN_NAN) return AMM_ERR_ARG;
  int rc = set_device(ws->device);
  if (rc) return rc;
  combine_records_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(
      (const amm_record*)dev_records, count, mode, ws->rec_dev, ws->rec_host_dev, ++ws->seq);
  AMM_CUDA(cudaGetLastError());
  ws->launches += 1;
  return AMM_OK;
}

int amm_workspace_create(int device, amm_workspace** out) {
  if (out == nullptr) return AMM_ERR_ARG;
  *out = nullptr;
  int count = 0;
  AMM_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return AMM_ERR_ARG;
  int rc = set_device(device);
  if (rc) return rc;
  amm_workspace* ws = new (std::nothrow) amm_workspace();
  if (!ws) return AMM_ERR_ARG;
  memset(ws, 0, sizeof(*ws));
  ws->device = device;
  cudaDeviceProp prop;
  AMM_CUDA(cudaGetDeviceProperties(&prop, device));
  ws->sm_count = prop.multiProcessorCount;
  ws->threads = 512;
  ws->ctas_per_sm = 2;
  ws->variant = 0;
  ws->max_grid = ws->sm_count * 32;
  // partials sized for the widest Best<int64_t> (40 B) -> 64 B slots
  AMM_CUDA(cudaMalloc(&ws->partials, (size_t)ws->max_grid * 64));
  AMM_CUDA(cudaMalloc((void**)&ws->ticket, 256));
  AMM_CUDA(cudaMemset(ws->ticket, 0, 256));
  AMM_CUDA(cudaMalloc((void**)&ws->rec_dev, sizeof(amm_record)));
  AMM_CUDA(cudaMemset(ws->rec_dev, 0, sizeof(amm_record)));
  AMM_CUDA(cudaHostAlloc((void**)&ws->rec_host, sizeof(amm_record), cudaHostAllocMapped));
  memset(ws->rec_host, 0, sizeof(amm_record));
  AMM_CUDA(cudaHostGetDevicePointer((void**)&ws->rec_host_dev, ws->rec_host, 0));
  AMM_CUDA(cudaDeviceSynchronize());
  *out = ws;
  return AMM_OK;
}

int amm_workspace_destroy(amm_workspace* ws) {
  if (!ws) return AMM_OK;
  set_device(ws->device);
  host_pipeline_destroy(ws);
  if (ws->partials) cudaFree(ws->partials);
  if (ws->ticket) cudaFree(ws->ticket);
  if (ws->rec_dev) cudaFree(ws->rec_dev);
  if (ws->rec_host) cudaFreeHost(ws->rec_host);
  delete ws;
  return AMM_OK;
}

int amm_argminmax_launch(int dtype, const void* dev_ptr, uint64_t n, int mode,
                         uint64_t global_offset, amm_workspace* ws, void* stream,
                         void* dev_record_out) {
  return launch(dtype, dev_ptr, n, mode, global_offset, ws, (cudaStream_t)stream, dev_record_out);
}

int amm_workspace_wait(amm_workspace* ws, void* stream, int mode, uint64_t out[2]) {
  if (!ws || !out) return AMM_ERR_ARG;
  // The kernel publishes the record into pinned host memory and writes `seq` last (after a
  // system-scope fence), so the host can pick the answer up by polling that word -- a few
  // microseconds sooner than cudaStreamSynchronize returns.  The stream is queried now and
  // then so that a failed launch surfaces as an error instead of a hang; after the poll
  // budget the call falls back to a blocking synchronize.
  volatile amm_record* h = ws->rec_host;
  const uint64_t want = ws->seq;
  bool seen = false;
  for (int spin = 0; spin < (1 << 22); ++spin) {
    if (h->seq == want) {
      seen = true;
      break;
    }
    if ((spin & 0x3FFF) == 0x3FFF) {
      cudaError_t q = cudaStreamQuery((cudaStream_t)stream);
      if (q == cudaSuccess) break;  // finished: the record must be there now
      if (q != cudaErrorNotReady) {
        g_last_cuda_error = (int)q;
        (void)cudaGetLastError();
        return AMM_ERR_CUDA;
      }
    }
  }
  if (!seen) AMM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  __sync_synchronize();
  amm_record r;
  memcpy(&r, (const void*)ws->rec_host, sizeof(r));
  if (r.seq != want) return AMM_ERR_CUDA;  // the kernel did not publish this call's record
  combine(&r, 1, mode, out);
  return AMM_OK;
}

int amm_workspace_device_record(amm_workspace* ws, void** dev_record) {
  if (!ws || !dev_record) return AMM_ERR_ARG;
  *dev_record = ws->rec_dev;
  return AMM_OK;
}

int amm_workspace_host_record(amm_workspace* ws, const amm_record** host_record) {
  if (!ws || !host_record) return AMM_ERR_ARG;
  *host_record = ws->rec_host;
  return AMM_OK;
}

int amm_argminmax(int dtype, const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws,
                  void* stream, uint64_t out[2]) {
  if (!out) return AMM_ERR_ARG;
  int rc = launch(dtype, dev_ptr, n, mode, 0, ws, (cudaStream_t)stream, nullptr);
  if (rc) return rc;
  return amm_workspace_wait(ws, stream, mode, out);
}

#define AMM_TYPED(NAME, CODE)                                                               \
  int amm_argminmax_##NAME(const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws,    \
                           void* stream, uint64_t out[2]) {                                 \
    return amm_argminmax(CODE, dev_ptr, n, mode, ws, stream, out);                          \
  }
AMM_TYPED(i8, AMM_I8)
AMM_TYPED(i16, AMM_I16)
AMM_TYPED(i32, AMM_I32)
AMM_TYPED(i64, AMM_I64)
AMM_TYPED(u8, AMM_U8)
AMM_TYPED(u16, AMM_U16)
AMM_TYPED(u32, AMM_U32)
AMM_TYPED(u64, AMM_U64)
AMM_TYPED(f16, AMM_F16)
AMM_TYPED(f32, AMM_F32)
AMM_TYPED(f64, AMM_F64)
#undef AMM_TYPED

int amm_argmin(int dtype, const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws,
               void* stream, uint64_t* out) {
  if (!out) return AMM_ERR_ARG;
  uint64_t both[2];
  int rc = amm_argminmax(dtype, dev_ptr, n, mode, ws, stream, both);
  if (rc) return rc;
  *out = both[0];
  return AMM_OK;
}

int amm_argmax(int dtype, const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws,
               void* stream, uint64_t* out) {
  if (!out) return AMM_ERR_ARG;
  uint64_t both[2];
  int rc = amm_argminmax(dtype, dev_ptr, n, mode, ws, stream, both);
  if (rc) return rc;
  *out = both[1];
  return AMM_OK;
}

int amm_combine_records(const amm_record* records, int count, int mode, uint64_t out[2]) {
  if (!records || !out || count < 1) return AMM_ERR_ARG;
  if (mode != AMM_IGNORE_NAN && mode != AMM_RETURN_NAN) return AMM_ERR_ARG;
  combine(records, count, mode, out);
  return AMM_OK;
}

const char* amm_status_string(int status) {
  switch (status) {
    case AMM_OK: return "ok";
    case AMM_ERR_EMPTY: return "empty input (the reference panics: assert!(!arr.is_empty()))";
    case AMM_ERR_ARG: return "invalid argument";
    case AMM_ERR_ALIGN: return "device pointer not aligned to the element size";
    case AMM_ERR_CUDA: return "CUDA error (see amm_last_cuda_error)";
    case AMM_ERR_NCCL: return "NCCL error or NCCL unavailable";
    case AMM_ERR_TOO_LARGE: return "input too large";
    default: return "unknown status";
  }
}

int amm_last_cuda_error(void) { return g_last_cuda_error; }

size_t amm_dtype_size(int dtype) {
  switch (dtype) {
    case AMM_I8: case AMM_U8: return 1;
    case AMM_I16: case AMM_U16: case AMM_F16: return 2;
    case AMM_I32: case AMM_U32: case AMM_F32: return 4;
    case AMM_I64: case AMM_U64: case AMM_F64: return 8;
    default: return 0;
  }
}

int amm_version(void) { return AMM_VERSION_MAJOR * 100 + AMM_VERSION_MINOR; }

int amm_workspace_set_tuning(amm_workspace* ws, int threads, int ctas_per_sm, int variant) {
  if (!ws) return AMM_ERR_ARG;
  if (threads > 0) {
    if (threads % 32 != 0 || threads < 64 || threads > MAX_THREADS) return AMM_ERR_ARG;
    if (threads != ws->threads) memset(ws->occupancy, 0, sizeof(ws->occupancy));
    ws->threads = threads;
  }
  if (ctas_per_sm > 0) {
    if (ctas_per_sm > 32) return AMM_ERR_ARG;
    ws->ctas_per_sm = ctas_per_sm;
  }
  if (variant >= 0) {
    if (variant >= kNumVariants) return AMM_ERR_ARG;
    ws->variant = variant;
  }
  return AMM_OK;
}

int amm_workspace_get_tuning(const amm_workspace* ws, int* threads, int* ctas_per_sm, int* variant,
                             int* sm_count) {
  if (!ws) return AMM_ERR_ARG;
  if (threads) *threads = ws->threads;
  if (ctas_per_sm) *ctas_per_sm = ws->ctas_per_sm;
  if (variant) *variant = ws->variant;
  if (sm_count) *sm_count = ws->sm_count;
  return AMM_OK;
}

int amm_variant_count(void) { return kNumVariants; }

const char* amm_variant_name(int variant) {
  if (variant < 0 || variant >= kNumVariants) return "";
  return kVariants[variant].name;
}

int amm_workspace_last_launch(const amm_workspace* ws, int* grid, int* threads,
                              uint64_t* launches_total) {
  if (!ws) return AMM_ERR_ARG;
  if (grid) *grid = ws->last_grid;
  if (threads) *threads = ws->last_threads;
  if (launches_total) *launches_total = ws->launches;
  return AMM_OK;
}

}  // extern "C"
