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
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <new>
#include <vector>

#include "../../include/argminmax_b200.h"
#include "amm_internal.h"
#include "scan_launch.cuh"

namespace amm {

thread_local int g_last_cuda_error = 0;

int launch_group0(int, amm_workspace*, ScanParams&, cudaStream_t);
int launch_group1(int, amm_workspace*, ScanParams&, cudaStream_t);
int launch_group2(int, amm_workspace*, ScanParams&, cudaStream_t);
int launch_group3(int, amm_workspace*, ScanParams&, cudaStream_t);
int launch_group4(int, amm_workspace*, ScanParams&, cudaStream_t);
int launch_group5(int, amm_workspace*, ScanParams&, cudaStream_t);
int launch_group6(int, amm_workspace*, ScanParams&, cudaStream_t);

int windows_group0(int, amm_workspace*, const WindowParams&, uint64_t, cudaStream_t);
int windows_group1(int, amm_workspace*, const WindowParams&, uint64_t, cudaStream_t);
int windows_group2(int, amm_workspace*, const WindowParams&, uint64_t, cudaStream_t);
int windows_group3(int, amm_workspace*, const WindowParams&, uint64_t, cudaStream_t);
int windows_group4(int, amm_workspace*, const WindowParams&, uint64_t, cudaStream_t);
int windows_group5(int, amm_workspace*, const WindowParams&, uint64_t, cudaStream_t);
int windows_group6(int, amm_workspace*, const WindowParams&, uint64_t, cudaStream_t);

static int windows_slot(int slot, amm_workspace* ws, const WindowParams& p, uint64_t mean,
                        cudaStream_t s) {
  switch (slot) {
    case 0: case 4: return windows_group0(slot, ws, p, mean, s);
    case 1: case 5: return windows_group1(slot, ws, p, mean, s);
    case 2: case 6: return windows_group2(slot, ws, p, mean, s);
    case 3: case 7: return windows_group3(slot, ws, p, mean, s);
    case 8: case 9: return windows_group4(slot, ws, p, mean, s);
    case 10: case 11: return windows_group5(slot, ws, p, mean, s);
    default: return windows_group6(slot, ws, p, mean, s);
  }
}

static int launch_slot(int slot, amm_workspace* ws, ScanParams& p, cudaStream_t s) {
  switch (slot) {
    case 0: case 4: return launch_group0(slot, ws, p, s);
    case 1: case 5: return launch_group1(slot, ws, p, s);
    case 2: case 6: return launch_group2(slot, ws, p, s);
    case 3: case 7: return launch_group3(slot, ws, p, s);
    case 8: case 9: return launch_group4(slot, ws, p, s);
    case 10: case 11: return launch_group5(slot, ws, p, s);
    default: return launch_group6(slot, ws, p, s);
  }
}

// The kernels are instantiated in scan_inst.cu, compiled once per policy group (parallel
// build); this is the strategy dispatch (Int / FloatIgnoreNaN / FloatReturnNaN,
// src/dtype_strategy.rs + src/lib.rs:272-664) down to a policy slot.
static int policy_slot(int dtype, int mode) {
  const bool ret = (mode == AMM_RETURN_NAN);
  switch (dtype) {
    case AMM_I8: return 0;
    case AMM_I16: return 1;
    case AMM_I32: return 2;
    case AMM_I64: return 3;
    case AMM_U8: return 4;
    case AMM_U16: return 5;
    case AMM_U32: return 6;
    case AMM_U64: return 7;
    case AMM_F16: return ret ? 9 : 8;
    case AMM_F32: return ret ? 11 : 10;
    case AMM_F64: return ret ? 13 : 12;
    default: return -1;
  }
}

static int dispatch(amm_workspace* ws, int dtype, int mode, ScanParams& p, cudaStream_t stream) {
  const int slot = policy_slot(dtype, mode);
  if (slot < 0) return AMM_ERR_ARG;
  return launch_slot(slot, ws, p, stream);
}

// Pinned host slot = 8 record words + 1 checksum word (see publish_record).  True when the
// slot holds the complete record of call `want`.
bool read_host_slot(const amm_workspace* ws, uint64_t want, amm_record* out) {
  const volatile uint64_t* h = reinterpret_cast<const volatile uint64_t*>(ws->rec_host);
  if (h[7] != want) return false;
  uint64_t w[9];
  for (int i = 0; i < 9; ++i) w[i] = h[i];
  if (mix_checksum(w, HOST_SLOT_SEED) != w[8] || w[7] != want) return false;
  memcpy(out, w, sizeof(amm_record));
  return true;
}

int set_device(int device) {
  int cur = -1;
  AMM_CUDA(cudaGetDevice(&cur));
  if (cur != device) AMM_CUDA(cudaSetDevice(device));
  return AMM_OK;
}

int validate_scan(int dtype, const void* dev_ptr, uint64_t n, int mode, bool empty_allowed) {
  if (dtype < 0 || dtype >= AMM_DTYPE_COUNT) return AMM_ERR_ARG;
  if (mode != AMM_IGNORE_NAN && mode != AMM_RETURN_NAN) return AMM_ERR_ARG;
  if (empty_allowed) {
    if (n != 0 && dev_ptr == nullptr) return AMM_ERR_ARG;
  } else {
    if (dev_ptr == nullptr) return AMM_ERR_ARG;
    if (n == 0) return AMM_ERR_EMPTY;
  }
  if (((uintptr_t)dev_ptr) % amm_dtype_size(dtype) != 0) return AMM_ERR_ALIGN;
  return AMM_OK;
}

int launch(int dtype, const void* dev_ptr, uint64_t n, int mode, uint64_t global_offset,
           amm_workspace* ws, cudaStream_t stream, void* dev_record_out, bool exchange) {
  if (ws == nullptr) return AMM_ERR_ARG;
  if (exchange) {
    // an empty shard still takes part in the exchange (it contributes "no value")
    if (ws->xworld < 2 || !ws->xpeers_dev) return AMM_ERR_ARG;
    if (ws->xbroken) return AMM_ERR_PEER;  // an earlier call timed out: re-create the exchange
  }
  {
    const int vrc = validate_scan(dtype, dev_ptr, n, mode, exchange);
    if (vrc) return vrc;
  }
  const size_t es = amm_dtype_size(dtype);
  DeviceGuard guard(ws->device);
  if (guard.rc) return guard.rc;
  ScanParams p;
  memset(&p, 0, sizeof(p));
  p.base = (const uint8_t*)dev_ptr;
  p.n = n;
  p.n_bytes = n * es;
#ifdef AMM_BOUNDS_CHECK
  // self-test of the checker: pretend the array is shorter than what the kernels will scan
  if (getenv("AMM_DEBUG_SHRINK_BYTES")) {
    const uint64_t k = strtoull(getenv("AMM_DEBUG_SHRINK_BYTES"), nullptr, 10);
    p.n_bytes = p.n_bytes > k ? p.n_bytes - k : 0;
  }
#endif
  p.global_offset = global_offset;
  p.partials = ws->partials;
  p.ticket = ws->ticket;
  p.rec_dev = dev_record_out ? (amm_record*)dev_record_out : ws->rec_dev;
  p.rec_host = ws->rec_host_dev;
  p.seq = ws->seq + 1;  // committed below, once the kernel is really on the stream
  p.mode = mode;
  p.publish_fence = ws->publish_fence;
  p.pipelined = (ws->pdl && ws->pipelined) ? 1 : 0;
  if (exchange) {
    p.world = ws->xworld;
    p.rank = ws->xrank;
    p.peer_slots = ws->xpeers_dev;
    p.xseq = ws->xseq + 1;
    p.xtimeout_ns = ws->xtimeout_ms * 1000000ull;
  }
  p.stamps = ws->stamps;
  const int rc = dispatch(ws, dtype, mode, p, stream);
  // the exchange sequence number only advances when the kernel really went out: a rank whose
  // launch failed must not get out of step with its peers' tags
  if (rc == AMM_OK) {
    ws->seq += 1;
    if (exchange) ws->xseq += 1;
  }
  return rc;
}

void exchange_destroy(amm_workspace* ws) {
  for (int i = 0; i < 16; ++i) {
    if (ws->xopened[i]) cudaIpcCloseMemHandle(ws->xopened[i]);
    ws->xopened[i] = nullptr;
  }
  if (ws->xpeers_dev) cudaFree(ws->xpeers_dev);
  if (ws->xbuf) cudaFree(ws->xbuf);
  ws->xpeers_dev = nullptr;
  ws->xbuf = nullptr;
  ws->xworld = 0;
  ws->xbroken = 0;
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
  ScanParams p;
  p.rec_dev = rec_dev;
  p.rec_host = rec_host;
  p.publish_fence = 0;
  publish_record(p, o);
}

}  // namespace amm

using namespace amm;

extern "C" {

int amm_workspace_destroy(amm_workspace* ws);

int amm_combine_records_launch(const void* dev_records, int count, int mode, amm_workspace* ws,
                               void* stream) {
  if (!dev_records || !ws || count < 1) return AMM_ERR_ARG;
  if (mode != AMM_IGNORE_NAN && mode != AMM_RETURN_NAN) return AMM_ERR_ARG;
  DeviceGuard guard(ws->device);
  if (guard.rc) return guard.rc;
  combine_records_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(
      (const amm_record*)dev_records, count, mode, ws->rec_dev, ws->rec_host_dev, ++ws->seq);
  AMM_CUDA(cudaGetLastError());
  ws->launches += 1;
  return AMM_OK;
}

static int workspace_init(amm_workspace* ws, int device) {
  memset(ws, 0, sizeof(*ws));
  ws->device = device;
  cudaDeviceProp prop;
  AMM_CUDA(cudaGetDeviceProperties(&prop, device));
  ws->sm_count = prop.multiProcessorCount;
  ws->smem_per_sm = prop.sharedMemPerMultiprocessor;
  ws->smem_reserved_per_cta = prop.reservedSharedMemPerBlock;
  ws->smem_optin_per_cta = prop.sharedMemPerBlockOptin;
  ws->threads = 0;      // automatic
  ws->ctas_per_sm = 0;  // automatic
  ws->variant = -1;     // automatic
  ws->direct_bytes = (uint64_t)8 << 10;
  ws->direct_threads = 256;
  ws->direct_ctas_per_sm = 2;
  // experiment knobs (tools/, profiling); the defaults are the tuned values
  ws->direct_auto = 1;
  if (getenv("AMM_DIRECT_BYTES")) {
    ws->direct_bytes = strtoull(getenv("AMM_DIRECT_BYTES"), nullptr, 10);
    ws->direct_auto = 0;
  }
  if (getenv("AMM_DIRECT_THREADS")) ws->direct_threads = atoi(getenv("AMM_DIRECT_THREADS"));
  if (getenv("AMM_DIRECT_CTAS")) ws->direct_ctas_per_sm = atoi(getenv("AMM_DIRECT_CTAS"));
  ws->direct_packed = getenv("AMM_DIRECT_PACKED") ? atoi(getenv("AMM_DIRECT_PACKED")) : 1;
  ws->direct_one_cta_vecs = getenv("AMM_DIRECT_ONE_CTA_VECS") ? atoi(getenv("AMM_DIRECT_ONE_CTA_VECS")) : 2048;
  ws->scan_packed = getenv("AMM_SCAN_PACKED") ? atoi(getenv("AMM_SCAN_PACKED")) : 1;
  ws->scan_contig = getenv("AMM_SCAN_CONTIG") ? atoi(getenv("AMM_SCAN_CONTIG")) : -1;
  ws->scan_dyn_pct = getenv("AMM_SCAN_DYN_PCT") ? atoi(getenv("AMM_SCAN_DYN_PCT")) : -1;
  ws->scan_ring = getenv("AMM_SCAN_RING") ? atoi(getenv("AMM_SCAN_RING")) : -1;
  ws->ring_ctas_per_sm = getenv("AMM_RING_CTAS") ? atoi(getenv("AMM_RING_CTAS")) : 2;
  if (ws->ring_ctas_per_sm < 1 || ws->ring_ctas_per_sm > 2) ws->ring_ctas_per_sm = 2;
  ws->scan_dyn_pipelined = getenv("AMM_SCAN_DYN_PIPELINED") ? atoi(getenv("AMM_SCAN_DYN_PIPELINED")) : 1;
  ws->direct_elems32 = getenv("AMM_DIRECT_ELEMS32") ? strtoull(getenv("AMM_DIRECT_ELEMS32"), nullptr, 10) : (1ull << 20);
  // 64-bit keys: 2^19 elements (4 MiB, the most the one-trip kernel holds) since its CTA reduce and grid
  // step were rebuilt (in-process A/B: i64 x 2^18 10.8 -> 10.2 us, x 2^19 11.7 -> 10.8; 2^16 before)
  ws->direct_elems64 = getenv("AMM_DIRECT_ELEMS64") ? strtoull(getenv("AMM_DIRECT_ELEMS64"), nullptr, 10) : (1ull << 19);
  ws->xtimeout_ms = getenv("AMM_PEER_TIMEOUT_MS") ? strtoull(getenv("AMM_PEER_TIMEOUT_MS"), nullptr, 10) : 60000;
  ws->direct_vecs_per_thread = getenv("AMM_DIRECT_VPT") ? atoi(getenv("AMM_DIRECT_VPT")) : 1;
  if (ws->direct_vecs_per_thread < 1 || ws->direct_vecs_per_thread > 16) ws->direct_vecs_per_thread = 1;
  if (ws->direct_threads < 64 || ws->direct_threads > MAX_THREADS || ws->direct_threads % 32)
    ws->direct_threads = 256;
  if (ws->direct_ctas_per_sm < 1 || ws->direct_ctas_per_sm > 32) ws->direct_ctas_per_sm = 2;
  ws->publish_fence = getenv("AMM_PUBLISH_FENCE") ? atoi(getenv("AMM_PUBLISH_FENCE")) : 0;
  ws->pdl = getenv("AMM_PDL") ? atoi(getenv("AMM_PDL")) : 1;
  ws->win_staged = getenv("AMM_WIN_STAGED") ? atoi(getenv("AMM_WIN_STAGED")) : 1;
  ws->win_fast = getenv("AMM_WIN_FAST") ? atoi(getenv("AMM_WIN_FAST")) : 1;
  ws->win_fixed = getenv("AMM_WIN_FIXED") ? atoi(getenv("AMM_WIN_FIXED")) : 1;
  ws->win_chunks_per_sm = getenv("AMM_WIN_CHUNKS_PER_SM") ? atoi(getenv("AMM_WIN_CHUNKS_PER_SM")) : 16;
  if (ws->win_chunks_per_sm < 4 || ws->win_chunks_per_sm > 256) ws->win_chunks_per_sm = 16;
  ws->win_chunked = getenv("AMM_WIN_CHUNKED") ? atoi(getenv("AMM_WIN_CHUNKED")) : 1;
  ws->win_staged_max = getenv("AMM_WIN_STAGED_MAX") ? atoi(getenv("AMM_WIN_STAGED_MAX")) : 4096;
  ws->win_lane_bytes = getenv("AMM_WIN_LANE_BYTES") ? atoi(getenv("AMM_WIN_LANE_BYTES")) : 128;
  if (ws->win_lane_bytes < 16 || ws->win_lane_bytes > 1024) ws->win_lane_bytes = 128;
  ws->max_grid = ws->sm_count * 32;
  // one 64-byte tagged slot per CTA (grid_reduce: five words + integrity word).  Zeroed: recycled
  // memory may hold a destroyed workspace's slots, whose sequence numbers also started at 1
  AMM_CUDA(cudaMalloc(&ws->partials, (size_t)ws->max_grid * 64));
  AMM_CUDA(cudaMemset(ws->partials, 0, (size_t)ws->max_grid * 64));
  AMM_CUDA(cudaMalloc((void**)&ws->ticket, WS_SCRATCH_BYTES));  // layout: scan_kernel.cuh, WS_*
  AMM_CUDA(cudaMemset(ws->ticket, 0, WS_SCRATCH_BYTES));
  {  // packed candidate words start at their identities (packed_grid_reduce)
    const unsigned long long ident[3] = {PACK_MIN_NONE, PACK_MAX_NONE, POS_NONE};
    const uint32_t at[3] = {WS_PACK_MIN, WS_PACK_MAX, WS_PACK_NAN};
    for (int i = 0; i < 3; ++i)
      AMM_CUDA(cudaMemcpy(ws->ticket + at[i], &ident[i], sizeof(ident[i]), cudaMemcpyHostToDevice));
  }
  AMM_CUDA(cudaMalloc((void**)&ws->rec_dev, sizeof(amm_record)));
  AMM_CUDA(cudaMemset(ws->rec_dev, 0, sizeof(amm_record)));
  AMM_CUDA(cudaHostAlloc((void**)&ws->rec_host, 128, cudaHostAllocMapped));
  memset(ws->rec_host, 0, 128);
  AMM_CUDA(cudaHostGetDevicePointer((void**)&ws->rec_host_dev, ws->rec_host, 0));
#ifdef AMM_PHASE_STAMPS
  AMM_CUDA(cudaMalloc((void**)&ws->stamps, (size_t)ws->max_grid * 8 * sizeof(uint64_t)));
  AMM_CUDA(cudaMemset(ws->stamps, 0, (size_t)ws->max_grid * 8 * sizeof(uint64_t)));
#endif
  AMM_CUDA(cudaDeviceSynchronize());
  return AMM_OK;
}

int amm_workspace_create(int device, amm_workspace** out) {
  if (out == nullptr) return AMM_ERR_ARG;
  *out = nullptr;
  int count = 0;
  AMM_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return AMM_ERR_ARG;
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  int rc = AMM_OK;
  amm_workspace* ws = new (std::nothrow) amm_workspace();
  if (!ws) return AMM_ERR_ARG;
  rc = workspace_init(ws, device);
  if (rc) {  // release whatever was allocated before the failure
    const int saved = g_last_cuda_error;
    amm_workspace_destroy(ws);
    g_last_cuda_error = saved;
    return rc;
  }
  *out = ws;
  return AMM_OK;
}

int amm_workspace_destroy(amm_workspace* ws) {
  if (!ws) return AMM_OK;
  DeviceGuard guard(ws->device);
  host_pipeline_destroy(ws);
  exchange_destroy(ws);
  if (ws->win_scratch) cudaFree(ws->win_scratch);
  if (ws->stamps) cudaFree(ws->stamps);
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

/* ---- fused cross-GPU exchange: setup ------------------------------------------------- */
int amm_exchange_create(amm_workspace* ws, int world, int rank, void* ipc_handle_out) {
  if (!ws || world < 2 || world > XCHG_MAX_WORLD || rank < 0 || rank >= world) return AMM_ERR_ARG;
  // One exchange per workspace: silently replacing a live one would free a buffer that peers
  // still have mapped (cudaIpcOpenMemHandle) and reset the sequence numbers under its user.
  if (ws->xbuf != nullptr) return AMM_ERR_ARG;
  DeviceGuard guard(ws->device);
  if (guard.rc) return guard.rc;
  const size_t bytes = (size_t)2 * world * XCHG_SLOT_WORDS * sizeof(uint64_t);
  AMM_CUDA(cudaMalloc((void**)&ws->xbuf, bytes));
  AMM_CUDA(cudaMemset(ws->xbuf, 0, bytes));
  AMM_CUDA(cudaMalloc((void**)&ws->xpeers_dev, sizeof(uint64_t*) * (size_t)world));
  AMM_CUDA(cudaDeviceSynchronize());
  ws->xworld = world;
  ws->xrank = rank;
  ws->xseq = 0;
  if (ipc_handle_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is exchanged as 64 bytes");
    cudaIpcMemHandle_t h;
    AMM_CUDA(cudaIpcGetMemHandle(&h, ws->xbuf));
    memcpy(ipc_handle_out, &h, sizeof(h));
  }
  return AMM_OK;
}

int amm_exchange_set_timeout(amm_workspace* ws, uint64_t milliseconds) {
  if (!ws) return AMM_ERR_ARG;
  ws->xtimeout_ms = milliseconds;
  return AMM_OK;
}

int amm_exchange_disconnect(amm_workspace* ws) {
  if (!ws) return AMM_ERR_ARG;
  DeviceGuard guard(ws->device);
  if (guard.rc) return guard.rc;
  for (int i = 0; i < 16; ++i) {
    if (ws->xopened[i]) cudaIpcCloseMemHandle(ws->xopened[i]);
    ws->xopened[i] = nullptr;
  }
  (void)cudaGetLastError();
  ws->xbroken = 1;  // no peer table any more: launches are refused until re-created
  return AMM_OK;
}

int amm_exchange_destroy(amm_workspace* ws) {
  if (!ws) return AMM_ERR_ARG;
  DeviceGuard guard(ws->device);
  if (guard.rc) return guard.rc;
  exchange_destroy(ws);
  (void)cudaGetLastError();
  return AMM_OK;
}

int amm_exchange_local_buffer(amm_workspace* ws, void** dev_ptr) {
  if (!ws || !dev_ptr || !ws->xbuf) return AMM_ERR_ARG;
  *dev_ptr = ws->xbuf;
  return AMM_OK;
}

static int exchange_upload(amm_workspace* ws, uint64_t** table) {
  AMM_CUDA(cudaMemcpy(ws->xpeers_dev, table, sizeof(uint64_t*) * (size_t)ws->xworld,
                      cudaMemcpyHostToDevice));
  return AMM_OK;
}

int amm_device_uuid(int device, void* uuid_out) {
  if (!uuid_out) return AMM_ERR_ARG;
  cudaDeviceProp prop;
  AMM_CUDA(cudaGetDeviceProperties(&prop, device));
  static_assert(sizeof(prop.uuid) == 16, "GPU UUIDs travel as 16 bytes");
  memcpy(uuid_out, &prop.uuid, 16);
  return AMM_OK;
}

int amm_exchange_connect_ipc(amm_workspace* ws, const void* ipc_handles, const void* device_ids) {
  if (!ws || !ipc_handles || ws->xworld < 2 || !ws->xbuf) return AMM_ERR_ARG;
  // Two ranks on ONE GPU (MPS, time slicing, a mis-set CUDA_VISIBLE_DEVICES) would have their
  // kernels spin-wait for each other on the same device: refuse, the caller falls back to NCCL.
  if (device_ids) {
    const unsigned char* ids = (const unsigned char*)device_ids;
    for (int a = 0; a < ws->xworld; ++a)
      for (int b = a + 1; b < ws->xworld; ++b)
        if (memcmp(ids + 16 * a, ids + 16 * b, 16) == 0) return AMM_ERR_ARG;
  }
  DeviceGuard guard(ws->device);
  if (guard.rc) return guard.rc;
  uint64_t* table[XCHG_MAX_WORLD];
  for (int r = 0; r < ws->xworld; ++r) {
    if (r == ws->xrank) {
      table[r] = ws->xbuf;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)ipc_handles + (size_t)r * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    AMM_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    ws->xopened[r] = ptr;
    table[r] = (uint64_t*)ptr;
  }
  return exchange_upload(ws, table);
}

int amm_exchange_connect_ptrs(amm_workspace* ws, void* const* peer_buffers) {
  if (!ws || !peer_buffers || ws->xworld < 2) return AMM_ERR_ARG;
  DeviceGuard guard(ws->device);
  if (guard.rc) return guard.rc;
  uint64_t* table[XCHG_MAX_WORLD];
  for (int r = 0; r < ws->xworld; ++r) table[r] = (uint64_t*)peer_buffers[r];
  table[ws->xrank] = ws->xbuf;
  return exchange_upload(ws, table);
}

int amm_argminmax_launch_exchange(int dtype, const void* dev_ptr, uint64_t n, int mode,
                                  uint64_t global_offset, amm_workspace* ws, void* stream) {
  return launch(dtype, dev_ptr, n, mode, global_offset, ws, (cudaStream_t)stream, nullptr, true);
}

int amm_workspace_wait(amm_workspace* ws, void* stream, int mode, uint64_t out[2]) {
  if (!ws || !out) return AMM_ERR_ARG;
  if (mode != AMM_IGNORE_NAN && mode != AMM_RETURN_NAN) return AMM_ERR_ARG;
  if (ws->seq == 0) return AMM_ERR_ARG;  // nothing was ever launched on this workspace
  // `stream` may be NULL = the default stream OF THE CURRENT DEVICE: make that the workspace's
  DeviceGuard guard(ws->device);
  if (guard.rc) return guard.rc;
  // The kernel publishes the record into pinned host memory as eight words plus a ninth,
  // position-sensitive checksum word (publish_record; no system-scope fence: it would cost
  // microseconds on the PCIe path).  The host accepts the slot once `seq` is this call's and the
  // checksum matches -- a torn slot is simply polled again -- which is a few microseconds sooner
  // than cudaStreamSynchronize returns.  The stream is queried now and then so that a failed
  // launch surfaces as an error instead of a hang; after the poll budget the call falls back to
  // a blocking synchronize.
  const uint64_t want = ws->seq;
  amm_record r;
  bool seen = false;
  for (int spin = 0; spin < (1 << 22) && !seen; ++spin) {
    seen = read_host_slot(ws, want, &r);
    if (!seen && (spin & 0x3FFF) == 0x3FFF) {
      cudaError_t q = cudaStreamQuery((cudaStream_t)stream);
      if (q == cudaSuccess) break;  // finished: the record must be there now
      if (q != cudaErrorNotReady) {
        g_last_cuda_error = (int)q;
        (void)cudaGetLastError();
        return AMM_ERR_CUDA;
      }
    }
  }
  if (!seen) {
    AMM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (!read_host_slot(ws, want, &r)) return AMM_ERR_CUDA;  // record never published
  }
  if (r.flags & AMM_REC_PEER_TIMEOUT) {
    ws->xbroken = 1;  // the parity banks are out of step now; amm_exchange_destroy + create again
    return AMM_ERR_PEER;
  }
  if (r.flags & AMM_REC_BOUNDS) return AMM_ERR_CUDA;  // debug builds: a load left the input range
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

/* ---- windowed / segmented form --------------------------------------------------------- */
// A few HUGE windows (>= 32 MiB each): one streaming scan per window (back to back on the
// stream, pipelined after the first) writing its 64-byte record to scratch, then one tiny
// kernel turns the records into (argmin, argmax) pairs.  Keeps every SM busy where the
// one-CTA-per-window kernel would use only n_windows CTAs.
__global__ void window_records_to_pairs_kernel(const amm_record* recs, const uint64_t* begins,
                                               const uint64_t* ends, uint64_t n_windows, int mode,
                                               uint64_t* out) {
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_windows) return;
  uint64_t imin, imax;
  if (begins[w] >= ends[w]) {
    imin = imax = AMM_NO_INDEX;
  } else {
    const amm_record r = recs[w];
    if (mode == AMM_RETURN_NAN && (r.flags & AMM_REC_HAS_NAN)) {
      imin = imax = r.nan_idx;
    } else if (!(r.flags & AMM_REC_HAS_VALUE)) {
      imin = imax = begins[w];
    } else {
      imin = r.min_idx;
      imax = r.max_idx;
    }
  }
  out[2 * w] = imin;
  out[2 * w + 1] = imax;
}

static int windows_huge(int dtype, const void* dev_ptr, uint64_t n, const uint64_t* dev_offsets,
                        uint64_t window, uint64_t n_windows, int mode, amm_workspace* ws,
                        cudaStream_t stream, uint64_t* dev_out) {
  const size_t es = amm_dtype_size(dtype);
  std::vector<uint64_t> bounds(2 * (size_t)n_windows);  // begins | ends
  if (dev_offsets) {
    std::vector<uint64_t> offs((size_t)n_windows + 1);
    AMM_CUDA(cudaMemcpyAsync(offs.data(), dev_offsets, sizeof(uint64_t) * offs.size(),
                             cudaMemcpyDeviceToHost, stream));
    AMM_CUDA(cudaStreamSynchronize(stream));
    for (uint64_t w = 0; w < n_windows; ++w) {
      uint64_t e = offs[w + 1] < n ? offs[w + 1] : n;
      uint64_t b = offs[w] < e ? offs[w] : e;
      bounds[w] = b;
      bounds[n_windows + w] = e;
    }
  } else {
    for (uint64_t w = 0; w < n_windows; ++w) {
      uint64_t b = w * window, e = b + window;
      e = e < n ? e : n;
      bounds[w] = b < e ? b : e;
      bounds[n_windows + w] = e;
    }
  }
  // scratch: records + bounds on the device, grown on demand and kept in the workspace
  const size_t need = (size_t)n_windows * (sizeof(amm_record) + 2 * sizeof(uint64_t));
  if (ws->win_scratch_bytes < need) {
    if (ws->win_scratch) AMM_CUDA(cudaFree(ws->win_scratch));
    ws->win_scratch = nullptr;
    ws->win_scratch_bytes = 0;
    AMM_CUDA(cudaMalloc(&ws->win_scratch, need));
    ws->win_scratch_bytes = need;
  }
  amm_record* recs = (amm_record*)ws->win_scratch;
  uint64_t* dbounds = (uint64_t*)(recs + n_windows);
  AMM_CUDA(cudaMemcpyAsync(dbounds, bounds.data(), sizeof(uint64_t) * bounds.size(),
                           cudaMemcpyHostToDevice, stream));
  const int saved = ws->pipelined;
  // The first scan is non-pipelined: it waits for whatever produced the data BEFORE it signals
  // launch_dependents (pdl_entry), so scans 2..N -- pipelined, no wait of their own before their
  // first load -- cannot start ahead of that producer either, even one that triggers early.
  ws->pipelined = 0;
  int rc = AMM_OK;
  for (uint64_t w = 0; w < n_windows && rc == AMM_OK; ++w) {
    const uint64_t b = bounds[w], e = bounds[n_windows + w];
    if (b == e) continue;
    rc = launch(dtype, (const uint8_t*)dev_ptr + b * es, e - b, mode, b, ws, stream, recs + w);
    ws->pipelined = 1;  // later scans follow one of ours and read data it did not write
  }
  ws->pipelined = saved;
  if (rc) return rc;
  window_records_to_pairs_kernel<<<(unsigned)((n_windows + 127) / 128), 128, 0, stream>>>(
      recs, dbounds, dbounds + n_windows, n_windows, mode, dev_out);
  AMM_CUDA(cudaGetLastError());
  ws->launches += 1;
  // `bounds` (pageable) was handed to cudaMemcpyAsync: it is staged before the call returns
  return AMM_OK;
}

constexpr int kNotApplicable = -12345;  // internal: "this path cannot take the input"

// Window bounds on the host: fixed windows are computed, ragged ones are read back from the
// device (one small copy + stream sync; only taken for few, large windows).
static int host_window_bounds(const uint64_t* dev_offsets, uint64_t n, uint64_t window,
                              uint64_t n_windows, cudaStream_t stream,
                              std::vector<uint64_t>& begins, std::vector<uint64_t>& ends) {
  begins.resize((size_t)n_windows);
  ends.resize((size_t)n_windows);
  if (dev_offsets) {
    std::vector<uint64_t> offs((size_t)n_windows + 1);
    AMM_CUDA(cudaMemcpyAsync(offs.data(), dev_offsets, sizeof(uint64_t) * offs.size(),
                             cudaMemcpyDeviceToHost, stream));
    AMM_CUDA(cudaStreamSynchronize(stream));
    for (uint64_t w = 0; w < n_windows; ++w) {
      const uint64_t e = offs[w + 1] < n ? offs[w + 1] : n;
      begins[w] = offs[w] < e ? offs[w] : e;
      ends[w] = e;
    }
  } else {
    for (uint64_t w = 0; w < n_windows; ++w) {
      uint64_t b = w * window, e = b + window;
      e = e < n ? e : n;
      begins[w] = b < e ? b : e;
      ends[w] = e;
    }
  }
  return AMM_OK;
}

// Chunk plan of windows_chunked (pure host logic, unit-tested without a GPU through
// amm_debug_plan_window_chunks): cuts the adjacent windows [begins[w], ends[w]) into equal pieces.
// chunk k = [cuts[k], cuts[k+1]); window w owns chunks [first[w], first[w+1]).
static int plan_window_chunks(const std::vector<uint64_t>& begins, const std::vector<uint64_t>& ends,
                              size_t es, int sm_count, int chunks_per_sm, std::vector<uint64_t>& cuts,
                              std::vector<uint64_t>& first) {
  const uint64_t n_windows = begins.size();
  if (n_windows == 0 || es == 0) return AMM_ERR_ARG;
  for (uint64_t w = 1; w < n_windows; ++w)
    if (begins[w] != ends[w - 1]) return kNotApplicable;  // not adjacent: caller falls back
  // Aim at ~16 equal chunks per SM (>= 128 KiB each, ~2400 in all): they run one per 256-thread
  // CTA and the hardware block scheduler balances the tail; the per-chunk reduce + locate stays
  // below 2 %.  Measured on 1 GiB with 4 / 8 / 16 / 32 / 64 chunks per SM: f32 269 x 4 MB windows
  // 4.9 / 5.6 / 5.9 / 5.9 / 5.4 TB/s, f64 135 x 8 MB -- / -- / 5.8 / 5.1 / 5.4 (at 32 the chunks
  // outnumber 4096 and go one per WARP, which for 8-byte dtypes is 1.4 resident rounds).
  const uint64_t slots = (uint64_t)sm_count * (uint64_t)chunks_per_sm;
  const uint64_t span = ends[n_windows - 1] - begins[0];
  const uint64_t room = slots > n_windows ? slots - n_windows : 1;
  uint64_t chunk = (span + room - 1) / room;  // elements per chunk
  const uint64_t min_chunk = ((uint64_t)128 << 10) / es;
  if (chunk < min_chunk) chunk = min_chunk;
  // chunk k = [cuts[k], cuts[k+1]); window w owns chunks [first[w], first[w+1]) -- windows are
  // adjacent, so the last cut of one window is the first cut of the next
  cuts.clear();
  first.assign((size_t)n_windows + 1, 0);
  cuts.reserve((size_t)(span / chunk) + 2 * (size_t)n_windows + 2);
  cuts.push_back(begins[0]);
  for (uint64_t w = 0; w < n_windows; ++w) {
    first[w] = cuts.size() - 1;
    const uint64_t len = ends[w] - begins[w];
    if (len == 0) continue;
    // equal pieces (a short last piece would leave its CTA idle while the others finish);
    // multiples of 256 elements keep the piece starts vector-aligned
    const uint64_t pieces = (len + chunk - 1) / chunk;
    const uint64_t piece = ((len + pieces - 1) / pieces + 255) / 256 * 256;
    for (uint64_t b = begins[w] + piece; b < ends[w]; b += piece) cuts.push_back(b);
    cuts.push_back(ends[w]);
  }
  first[n_windows] = cuts.size() - 1;
  return AMM_OK;
}

// Few LARGE windows (1 MiB ... 64 MiB each, fewer than 4 per SM): one CTA per window would
// leave most of the GPU idle and one streaming scan per window pays ~6 us each.  Instead every
// window is cut into equal chunks (>= 128 KiB, about 16 per SM), the chunks run through the
// ordinary windowed kernels as if they were windows -- thousands of them, so the whole GPU
// streams -- and a combine kernel folds each window's chunk answers (first occurrence: the
// lowest chunk wins ties).  One extra launch, < 1 % extra traffic.
static int windows_chunked(int dtype, int slot, const void* dev_ptr, uint64_t n,
                           const uint64_t* dev_offsets, uint64_t window, uint64_t n_windows,
                           int mode, amm_workspace* ws, cudaStream_t stream, uint64_t* dev_out) {
  const size_t es = amm_dtype_size(dtype);
  std::vector<uint64_t> begins, ends;
  int rc = host_window_bounds(dev_offsets, n, window, n_windows, stream, begins, ends);
  if (rc) return rc;
  std::vector<uint64_t> cuts, first;
  rc = plan_window_chunks(begins, ends, es, ws->sm_count, ws->win_chunks_per_sm, cuts, first);
  if (rc) return rc;
  const uint64_t n_chunks = cuts.size() - 1;
  if (n_chunks == 0) return kNotApplicable;  // every window is empty: the plain kernels write the NO_INDEX pairs
  // scratch on the device: cuts | first | pairs
  const size_t need = sizeof(uint64_t) * (cuts.size() + first.size() + 2 * (size_t)n_chunks);
  if (ws->win_scratch_bytes < need) {
    if (ws->win_scratch) AMM_CUDA(cudaFree(ws->win_scratch));
    ws->win_scratch = nullptr;
    ws->win_scratch_bytes = 0;
    AMM_CUDA(cudaMalloc(&ws->win_scratch, need));
    ws->win_scratch_bytes = need;
  }
  uint64_t* d_cuts = (uint64_t*)ws->win_scratch;
  uint64_t* d_first = d_cuts + cuts.size();
  uint64_t* d_pairs = d_first + first.size();
  AMM_CUDA(cudaMemcpyAsync(d_cuts, cuts.data(), sizeof(uint64_t) * cuts.size(),
                           cudaMemcpyHostToDevice, stream));
  AMM_CUDA(cudaMemcpyAsync(d_first, first.data(), sizeof(uint64_t) * first.size(),
                           cudaMemcpyHostToDevice, stream));
  WindowParams pc;
  pc.window = 0;
  pc.base = (const uint8_t*)dev_ptr;
  pc.n = n;
  pc.offsets = d_cuts;
  pc.n_windows = n_chunks;
  pc.out = d_pairs;
  pc.mode = mode;
  rc = windows_slot(slot, ws, pc, n / n_chunks, stream);
  if (rc) return rc;
  WindowParams pw;
  pw.base = (const uint8_t*)dev_ptr;
  pw.n = n;
  pw.offsets = dev_offsets;
  pw.window = window;
  pw.n_windows = n_windows;
  pw.out = dev_out;
  pw.mode = mode;
  pw.chunk_first = d_first;
  pw.chunk_pairs = d_pairs;
  return windows_slot(slot, ws, pw, n / n_windows, stream);
}

static int windows_common(int dtype, const void* dev_ptr, uint64_t n, const uint64_t* dev_offsets,
                          uint64_t window, uint64_t n_windows, int mode, amm_workspace* ws,
                          void* stream, uint64_t* dev_out) {
  if (!ws || !dev_ptr || !dev_out) return AMM_ERR_ARG;
  if (mode != AMM_IGNORE_NAN && mode != AMM_RETURN_NAN) return AMM_ERR_ARG;
  const int slot = policy_slot(dtype, mode);
  if (slot < 0) return AMM_ERR_ARG;
  if (n == 0 || n_windows == 0) return AMM_ERR_EMPTY;
  if (((uintptr_t)dev_ptr) % amm_dtype_size(dtype) != 0) return AMM_ERR_ALIGN;
  DeviceGuard guard(ws->device);
  if (guard.rc) return guard.rc;
  int rc = AMM_OK;
  WindowParams p;
  p.base = (const uint8_t*)dev_ptr;
  p.n = n;
  p.offsets = dev_offsets;
  p.window = window;
  p.n_windows = n_windows;
  p.out = dev_out;
  p.mode = mode;
  const uint64_t mean = n / n_windows;
  // Few, large windows: the one-CTA-per-window kernel keeps only n_windows CTAs busy (~40 GB/s
  // each, ~180 CTAs saturate HBM); a streaming scan per window costs max(~6 us, bytes / HBM rate)
  // but uses the whole GPU.  Pick whichever the model says is faster.
  const double total_bytes = (double)n * (double)amm_dtype_size(dtype);
  const double win_bytes = total_bytes / (double)n_windows;
  const double busy_ctas = n_windows < 180 ? (double)n_windows : 180.0;
  const double t_cta_us = total_bytes / (busy_ctas * 40.0e3);
  const double per_scan_us = win_bytes / 7.0e6 > 6.0 ? win_bytes / 7.0e6 : 6.0;
  const double t_scan_us = (double)n_windows * per_scan_us;
  // ... unless the windows are of middling size: then chunks of them fill the GPU in one launch
  const double t_chunk_us = total_bytes / (amm_dtype_size(dtype) == 8 ? 4.9e6 : 5.8e6) + 8.0;
  const double t_huge_us = (double)n_windows * (win_bytes / 6.9e6 + 0.9 > 6.5 ? win_bytes / 6.9e6 + 0.9 : 6.5);
  if (ws->win_chunked && n_windows < (uint64_t)ws->sm_count * 4 && win_bytes >= (double)(1 << 20) &&
      t_chunk_us < t_huge_us) {
    rc = windows_chunked(dtype, slot, dev_ptr, n, dev_offsets, window, n_windows, mode, ws,
                         (cudaStream_t)stream, dev_out);
    if (rc != kNotApplicable) return rc;  // windows not adjacent (unordered offsets): paths below
  }
  if (n_windows <= 65536 && win_bytes >= (double)(1 << 20) && t_scan_us < t_cta_us)
    return windows_huge(dtype, dev_ptr, n, dev_offsets, window, n_windows, mode, ws,
                        (cudaStream_t)stream, dev_out);
  return windows_slot(slot, ws, p, mean, (cudaStream_t)stream);
}

int amm_argminmax_windows(int dtype, const void* dev_ptr, uint64_t n, const uint64_t* dev_offsets,
                          uint64_t n_windows, int mode, amm_workspace* ws, void* stream,
                          uint64_t* dev_out) {
  if (!dev_offsets) return AMM_ERR_ARG;
  return windows_common(dtype, dev_ptr, n, dev_offsets, 0, n_windows, mode, ws, stream, dev_out);
}

int amm_argminmax_fixed_windows(int dtype, const void* dev_ptr, uint64_t n, uint64_t window,
                                int mode, amm_workspace* ws, void* stream, uint64_t* dev_out) {
  if (window == 0) return AMM_ERR_ARG;
  return windows_common(dtype, dev_ptr, n, nullptr, window, (n + window - 1) / window, mode, ws,
                        stream, dev_out);
}

/* ---- device memory helpers (what DeviceSlice::from_slice needs on the Rust side) ---- */
int amm_device_count(int* count) {
  if (!count) return AMM_ERR_ARG;
  AMM_CUDA(cudaGetDeviceCount(count));
  return AMM_OK;
}

int amm_device_alloc(int device, size_t bytes, void** dev_ptr) {
  if (!dev_ptr) return AMM_ERR_ARG;
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  AMM_CUDA(cudaMalloc(dev_ptr, bytes ? bytes : 1));
  return AMM_OK;
}

int amm_device_free(int device, void* dev_ptr) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  AMM_CUDA(cudaFree(dev_ptr));
  return AMM_OK;
}

int amm_copy_to_device(int device, void* dst_dev, const void* src_host, size_t bytes) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  AMM_CUDA(cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice));
  return AMM_OK;
}

int amm_copy_to_host(int device, void* dst_host, const void* src_dev, size_t bytes) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  AMM_CUDA(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
  return AMM_OK;
}

/* Pure host logic of the chunked window path, exposed for the CPU test-suite: plans the chunks
   of `n_windows` adjacent windows.  Returns AMM_OK, AMM_ERR_ARG, AMM_ERR_TOO_LARGE (outputs too
   small) or 1000 when the windows are not adjacent (the path then does not apply). */
int amm_debug_plan_window_chunks(const uint64_t* begins, const uint64_t* ends, uint64_t n_windows,
                                 size_t elem_size, int sm_count, int chunks_per_sm,
                                 uint64_t* cuts_out, uint64_t cuts_capacity, uint64_t* n_cuts,
                                 uint64_t* first_out) {
  if (!begins || !ends || !cuts_out || !n_cuts || !first_out || sm_count < 1 || chunks_per_sm < 1)
    return AMM_ERR_ARG;
  std::vector<uint64_t> b(begins, begins + n_windows), e(ends, ends + n_windows), cuts, first;
  const int rc = plan_window_chunks(b, e, elem_size, sm_count, chunks_per_sm, cuts, first);
  if (rc == kNotApplicable) return 1000;
  if (rc) return rc;
  if (cuts.size() > cuts_capacity) return AMM_ERR_TOO_LARGE;
  memcpy(cuts_out, cuts.data(), cuts.size() * sizeof(uint64_t));
  memcpy(first_out, first.data(), first.size() * sizeof(uint64_t));
  *n_cuts = cuts.size();
  return AMM_OK;
}

/* Host wall-clock latency of the synchronous call, measured inside the library so that no
   binding overhead (ctypes, JNI, ...) is included: `calls` timed amm_argminmax calls after
   calls/10 + 10 warm-up calls; reports median / min / p99 in microseconds. */
int amm_measure_call_latency(int dtype, const void* dev_ptr, uint64_t n, int mode,
                             amm_workspace* ws, void* stream, int calls, double out_us[3]) {
  if (!out_us || calls < 1) return AMM_ERR_ARG;
  uint64_t r[2];
  for (int i = 0; i < calls / 10 + 10; ++i) {
    int rc = amm_argminmax(dtype, dev_ptr, n, mode, ws, stream, r);
    if (rc) return rc;
  }
  std::vector<double> us((size_t)calls);
  for (int i = 0; i < calls; ++i) {
    const auto t0 = std::chrono::steady_clock::now();
    int rc = amm_argminmax(dtype, dev_ptr, n, mode, ws, stream, r);
    const auto t1 = std::chrono::steady_clock::now();
    if (rc) return rc;
    us[(size_t)i] = std::chrono::duration<double, std::micro>(t1 - t0).count();
  }
  std::sort(us.begin(), us.end());
  out_us[0] = us[us.size() / 2];
  out_us[1] = us[0];
  out_us[2] = us[us.size() * 99 / 100];
  return AMM_OK;
}

/* Pure host logic of the streaming kernel's tile plan, exposed for the CPU test-suite
   (tests/test_scan_plan.py replays the kernel's index formulas on it).  out[0..11] = head_elems,
   n_vec, tail_start, per_vecs, grid, n_full_tiles, rem_vecs, rem_owner, dyn_tiles, dyn_koff,
   kind (0 grid-stride, 1 contiguous, 2 dynamic tail), status. */
int amm_debug_plan_scan(uint64_t base_addr, uint64_t n, int elem_size, int vec_bytes, int u, int threads,
                        uint64_t grid_cap, int sm_count, int pipelined, int scan_contig, int scan_dyn_pct,
                        uint64_t out[12]) {
  if (!out || n == 0 || threads < 32 || u < 1 || (vec_bytes != 16 && vec_bytes != 32) ||
      (elem_size != 1 && elem_size != 2 && elem_size != 4 && elem_size != 8) || base_addr % (uint64_t)elem_size)
    return AMM_ERR_ARG;
  const ScanPlanKnobs knobs = {sm_count, scan_contig, scan_dyn_pct, 1};
  const ScanPlan pl = plan_scan(base_addr, n, (uint32_t)elem_size, (uint32_t)vec_bytes, (uint32_t)u, (uint32_t)threads,
                                grid_cap, pipelined != 0, knobs);
  const uint64_t v[12] = {pl.head_elems, pl.n_vec, pl.tail_start, pl.per_vecs, pl.grid, pl.n_full_tiles,
                          pl.rem_vecs, pl.rem_owner, pl.dyn_tiles, pl.dyn_koff, (uint64_t)pl.kind, (uint64_t)pl.status};
  memcpy(out, v, sizeof(v));
  return AMM_OK;
}

/* Pure host logic of the latency kernel's launch shape (scan_launch.cuh: plan_direct), exposed for
   the CPU test-suite.  out[0..6] = head_elems, n_vec, tail_start, grid, threads, vectors per thread,
   fits (0: the input does not fit one trip -> streaming kernel). */
int amm_debug_plan_direct(uint64_t base_addr, uint64_t n, int elem_size, int sm_count, int ctas_per_sm,
                          int threads, int min_vpt, int one_cta_vecs, uint64_t out[8]) {
  if (!out || n == 0 || sm_count < 1 || ctas_per_sm < 1 || threads < 32 || threads % 32 || threads > MAX_THREADS ||
      min_vpt < 1 || (elem_size != 1 && elem_size != 2 && elem_size != 4 && elem_size != 8) ||
      base_addr % (uint64_t)elem_size)
    return AMM_ERR_ARG;
  const DirectPlanKnobs knobs = {sm_count, ctas_per_sm, threads, min_vpt, one_cta_vecs};
  const DirectPlan pl = plan_direct(base_addr, n, (uint32_t)elem_size, knobs);
  const uint64_t v[8] = {pl.head_elems, pl.n_vec, pl.tail_start, pl.grid, pl.threads, pl.vpt, pl.fits ? 1u : 0u, 0};
  memcpy(out, v, sizeof(v));
  return AMM_OK;
}

/* AMM_PHASE_STAMPS builds only (tools/phase_stamps.py): copies the %globaltimer stamps of the
   most recent scan, 8 words per CTA, for `ctas` CTAs.  Other builds: AMM_ERR_ARG. */
int amm_debug_read_stamps(amm_workspace* ws, uint64_t* out, int ctas) {
  if (!ws || !out || ctas < 1 || !ws->stamps || ctas > ws->max_grid) return AMM_ERR_ARG;
  DeviceGuard guard(ws->device);
  if (guard.rc) return guard.rc;
  AMM_CUDA(cudaMemcpy(out, ws->stamps, (size_t)ctas * 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return AMM_OK;
}

/* The same for the collective call (amm_argminmax_launch_exchange + amm_workspace_wait): every
   rank of the exchange must call it with the same `calls`. */
int amm_measure_exchange_latency(int dtype, const void* dev_ptr, uint64_t n, int mode,
                                 uint64_t global_offset, amm_workspace* ws, void* stream, int calls,
                                 double out_us[3]) {
  if (!out_us || calls < 1) return AMM_ERR_ARG;
  uint64_t r[2];
  auto once = [&]() -> int {
    int rc = launch(dtype, dev_ptr, n, mode, global_offset, ws, (cudaStream_t)stream, nullptr, true);
    if (rc) return rc;
    return amm_workspace_wait(ws, stream, mode, r);
  };
  for (int i = 0; i < calls / 10 + 10; ++i) {
    int rc = once();
    if (rc) return rc;
  }
  std::vector<double> us((size_t)calls);
  for (int i = 0; i < calls; ++i) {
    const auto t0 = std::chrono::steady_clock::now();
    int rc = once();
    const auto t1 = std::chrono::steady_clock::now();
    if (rc) return rc;
    us[(size_t)i] = std::chrono::duration<double, std::micro>(t1 - t0).count();
  }
  std::sort(us.begin(), us.end());
  out_us[0] = us[us.size() / 2];
  out_us[1] = us[0];
  out_us[2] = us[us.size() * 99 / 100];
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
    case AMM_ERR_PEER: return "peer exchange failed (a rank did not take part in the call in time; re-create the exchange)";
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
  if (threads > 0 && (threads % 32 != 0 || threads < 64 || threads > MAX_THREADS)) return AMM_ERR_ARG;
  if (ctas_per_sm > 32) return AMM_ERR_ARG;
  if (variant >= kNumVariants) return AMM_ERR_ARG;
  ws->threads = threads > 0 ? threads : 0;
  ws->ctas_per_sm = ctas_per_sm > 0 ? ctas_per_sm : 0;
  ws->variant = variant >= 0 ? variant : -1;
  return AMM_OK;
}

int amm_workspace_set_pipelined(amm_workspace* ws, int on) {
  if (!ws) return AMM_ERR_ARG;
  ws->pipelined = on ? 1 : 0;
  return AMM_OK;
}

int amm_workspace_set_direct_threshold(amm_workspace* ws, uint64_t bytes) {
  if (!ws) return AMM_ERR_ARG;
  ws->direct_bytes = bytes;
  ws->direct_auto = 0;
  return AMM_OK;
}

int amm_workspace_get_tuning(const amm_workspace* ws, int* threads, int* ctas_per_sm, int* variant,
                             int* sm_count) {
  if (!ws) return AMM_ERR_ARG;
  if (threads) *threads = ws->eff_threads;
  if (ctas_per_sm) *ctas_per_sm = ws->eff_ctas_per_sm;
  if (variant) *variant = ws->eff_variant;
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
