// host_slice.cu -- `&[T]` / Vec<T> in HOST memory (the step before the path,
// src/lib.rs:679-712): the slice is cut into chunks, chunk c is copied H2D on
// stream c%2 and scanned on the same stream by the ordinary device kernel with
// global_offset = chunk start, so the copy of chunk c+1 overlaps the scan of
// chunk c.  The per-chunk 64-byte records are combined on the host in chunk order
// (strict compares => first occurrence).  End to end this is PCIe-bound.
//
// PAGEABLE input (what a Vec<T> or a numpy array is): cudaMemcpyAsync from pageable memory
// goes through the driver's single-threaded staging copy (11 GB/s measured here against 55 GB/s
// from pinned memory).  Such input is therefore bounced through a ring of pinned buffers filled by a
// small team of host threads (parallel memcpy), so that the DMA engine always reads pinned memory:
// memcpy(chunk c) || H2D(chunk c-1) || scan(chunk c-2).
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <new>
#include <thread>
#include <vector>

#include "amm_internal.h"

constexpr int kBounce = 3;  // pinned bounce buffers for pageable input

struct amm_host_pipeline {
  size_t chunk_bytes;
  void* stage[2];
  cudaStream_t stream[2];
  amm_workspace* child[2];
  void* bounce[kBounce];         // pinned, chunk_bytes each, allocated on first pageable call
  cudaEvent_t bounce_done[kBounce];  // H2D out of bounce[i] has completed
  int copy_threads;
  size_t bounce_bytes;  // chunk size of the bounced (pageable) path, <= chunk_bytes
};

namespace {
// A team of host threads that copies one chunk at a time, each thread its own slice.  The
// threads live for one call; between chunks they spin on `go` (a call is PCIe-bound and short).
struct CopyTeam {
  std::vector<std::thread> threads;
  std::atomic<uint64_t> go{0};    // chunks released so far
  std::atomic<uint64_t> done{0};  // slices finished so far
  std::atomic<bool> quit{false};
  const uint8_t* src = nullptr;
  uint8_t* dst = nullptr;
  size_t bytes = 0;
  int nthreads = 0;
  explicit CopyTeam(int n) : nthreads(n) {
    for (int t = 1; t < n; ++t) threads.emplace_back([this, t] { worker(t); });
  }
  ~CopyTeam() {
    quit.store(true, std::memory_order_release);
    for (auto& th : threads) th.join();
  }
  void slice(int t) const {
    const size_t per = ((bytes + nthreads - 1) / nthreads + 4095) & ~(size_t)4095;
    const size_t b = per * (size_t)t;
    if (b >= bytes) return;
    const size_t len = bytes - b < per ? bytes - b : per;
    memcpy(dst + b, src + b, len);
  }
  void worker(int t) {
    uint64_t seen = 0;
    for (;;) {
      while (go.load(std::memory_order_acquire) == seen) {
        if (quit.load(std::memory_order_acquire)) return;
        std::this_thread::yield();
      }
      ++seen;
      slice(t);
      done.fetch_add(1, std::memory_order_release);
    }
  }
  // copy `n` bytes with the whole team (the caller is member 0); returns when all are done
  void copy(uint8_t* d, const uint8_t* s, size_t n) {
    dst = d;
    src = s;
    bytes = n;
    const uint64_t target = done.load(std::memory_order_relaxed) + (uint64_t)(nthreads - 1);
    go.fetch_add(1, std::memory_order_release);
    slice(0);
    while (done.load(std::memory_order_acquire) < target) {
    }
  }
};
}  // namespace

namespace amm {

static int pipeline_create(amm_workspace* ws) {
  amm_host_pipeline* p = new (std::nothrow) amm_host_pipeline();
  if (!p) return AMM_ERR_ARG;
  p->chunk_bytes = (size_t)32 << 20;
  const char* env = getenv("AMM_HOST_CHUNK_MB");
  if (env && atoi(env) > 0) p->chunk_bytes = (size_t)atoi(env) << 20;
  p->stage[0] = p->stage[1] = nullptr;
  p->child[0] = p->child[1] = nullptr;
  p->stream[0] = p->stream[1] = nullptr;
  for (int i = 0; i < kBounce; ++i) {
    p->bounce[i] = nullptr;
    p->bounce_done[i] = nullptr;
  }
  int hw = (int)std::thread::hardware_concurrency();
  p->copy_threads = hw >= 16 ? 8 : (hw >= 4 ? hw / 2 : 1);
  p->bounce_bytes = (size_t)16 << 20;
  const char* bb = getenv("AMM_HOST_BOUNCE_MB");
  if (bb && atoi(bb) > 0) p->bounce_bytes = (size_t)atoi(bb) << 20;
  if (p->bounce_bytes > p->chunk_bytes) p->bounce_bytes = p->chunk_bytes;
  const char* ct = getenv("AMM_HOST_COPY_THREADS");
  if (ct && atoi(ct) > 0 && atoi(ct) <= 64) p->copy_threads = atoi(ct);
  ws->pipe = p;
  for (int i = 0; i < 2; ++i) {
    AMM_CUDA(cudaMalloc(&p->stage[i], p->chunk_bytes));
    AMM_CUDA(cudaStreamCreateWithFlags(&p->stream[i], cudaStreamNonBlocking));
    int rc = amm_workspace_create(ws->device, &p->child[i]);
    if (rc) return rc;
  }
  return AMM_OK;
}

void host_pipeline_destroy(amm_workspace* ws) {
  amm_host_pipeline* p = ws->pipe;
  if (!p) return;
  for (int i = 0; i < 2; ++i) {
    if (p->stream[i]) {
      cudaStreamSynchronize(p->stream[i]);
      cudaStreamDestroy(p->stream[i]);
    }
    if (p->stage[i]) cudaFree(p->stage[i]);
    if (p->child[i]) amm_workspace_destroy(p->child[i]);
  }
  for (int i = 0; i < kBounce; ++i) {
    if (p->bounce_done[i]) cudaEventDestroy(p->bounce_done[i]);
    if (p->bounce[i]) cudaFreeHost(p->bounce[i]);
  }
  delete p;
  ws->pipe = nullptr;
}

}  // namespace amm

using namespace amm;

static int host_scan(int dtype, const void* host_ptr, uint64_t n, int mode, uint64_t global_offset,
                     amm_workspace* ws, std::vector<amm_record>& recs) {
  if (!ws || !host_ptr) return AMM_ERR_ARG;
  if (dtype < 0 || dtype >= AMM_DTYPE_COUNT) return AMM_ERR_ARG;
  if (mode != AMM_IGNORE_NAN && mode != AMM_RETURN_NAN) return AMM_ERR_ARG;
  if (n == 0) return AMM_ERR_EMPTY;
  DeviceGuard guard(ws->device);
  if (guard.rc) return guard.rc;
  int rc = AMM_OK;
  if (!ws->pipe) {
    rc = pipeline_create(ws);
    if (rc) {
      host_pipeline_destroy(ws);
      return rc;
    }
  }
  amm_host_pipeline* p = ws->pipe;
  const size_t es = amm_dtype_size(dtype);
  // pageable input of more than one chunk: bounce through pinned memory (see the file header),
  // in smaller chunks -- with 16 MiB the three bounce buffers stay in the host's last-level
  // cache, so the DMA engine reads what the copy threads just wrote without a DRAM round trip
  // (measured, 1 GiB f32: 53 GB/s at 16 MiB chunks, 32-44 at 32 MiB, 29-40 at 128 MiB)
  bool bounce = false;
  if (n * es > p->chunk_bytes && p->copy_threads > 0 && !getenv("AMM_HOST_NO_BOUNCE")) {
    cudaPointerAttributes attr;
    const cudaError_t e = cudaPointerGetAttributes(&attr, host_ptr);
    if (e != cudaSuccess) (void)cudaGetLastError();
    bounce = (e != cudaSuccess) || attr.type == cudaMemoryTypeUnregistered;
  }
  const uint64_t chunk_elems = (bounce ? p->bounce_bytes : p->chunk_bytes) / es;
  const uint64_t nchunks = (n + chunk_elems - 1) / chunk_elems;
  for (int i = 0; i < 2; ++i) {
    amm_workspace_set_tuning(p->child[i], ws->threads, ws->ctas_per_sm, ws->variant);
    p->child[i]->direct_bytes = ws->direct_bytes;
    p->child[i]->direct_auto = ws->direct_auto;
    p->child[i]->direct_packed = ws->direct_packed;
    p->child[i]->direct_one_cta_vecs = ws->direct_one_cta_vecs;
    p->child[i]->direct_vecs_per_thread = ws->direct_vecs_per_thread;
    p->child[i]->direct_elems32 = ws->direct_elems32;
    p->child[i]->direct_elems64 = ws->direct_elems64;
    p->child[i]->scan_packed = ws->scan_packed;
    p->child[i]->scan_contig = ws->scan_contig;
    p->child[i]->scan_dyn_pct = ws->scan_dyn_pct;
    p->child[i]->scan_ring = ws->scan_ring;
    p->child[i]->scan_dyn_pipelined = ws->scan_dyn_pipelined;
    p->child[i]->pdl = ws->pdl;
    p->child[i]->direct_threads = ws->direct_threads;
    p->child[i]->direct_ctas_per_sm = ws->direct_ctas_per_sm;
  }
  recs.clear();
  recs.reserve((size_t)nchunks);
  const uint8_t* src = (const uint8_t*)host_ptr;
  if (bounce && !p->bounce[0]) {
    for (int i = 0; i < kBounce; ++i) {
      AMM_CUDA(cudaHostAlloc(&p->bounce[i], p->bounce_bytes, cudaHostAllocDefault));
      AMM_CUDA(cudaEventCreateWithFlags(&p->bounce_done[i], cudaEventDisableTiming));
    }
  }
  CopyTeam* team = bounce ? new (std::nothrow) CopyTeam(p->copy_threads) : nullptr;
  struct TeamGuard {
    CopyTeam* t;
    ~TeamGuard() { delete t; }
  } team_guard{team};
  if (bounce && !team) return AMM_ERR_ARG;
  for (uint64_t c = 0; c < nchunks; ++c) {
    const int s = (int)(c & 1);
    const uint64_t begin = c * chunk_elems;
    const uint64_t len = (n - begin) < chunk_elems ? (n - begin) : chunk_elems;
    const int b = (int)(c % kBounce);
    if (bounce) {  // fill the bounce buffer while the DMAs of the two chunks before are in flight
      if (c >= (uint64_t)kBounce) AMM_CUDA(cudaEventSynchronize(p->bounce_done[b]));  // its DMA is over
      team->copy((uint8_t*)p->bounce[b], src + begin * es, len * es);
    }
    if (c >= 2) {  // stage s is free once chunk c-2 has been scanned; harvest its record
      AMM_CUDA(cudaStreamSynchronize(p->stream[s]));
      amm_record r;
      if (!read_host_slot(p->child[s], p->child[s]->seq, &r)) return AMM_ERR_CUDA;
      recs.push_back(r);
    }
    if (bounce) {
      AMM_CUDA(cudaMemcpyAsync(p->stage[s], p->bounce[b], len * es, cudaMemcpyHostToDevice,
                               p->stream[s]));
      AMM_CUDA(cudaEventRecord(p->bounce_done[b], p->stream[s]));
    } else {
      AMM_CUDA(cudaMemcpyAsync(p->stage[s], src + begin * es, len * es, cudaMemcpyHostToDevice,
                               p->stream[s]));
    }
    rc = launch(dtype, p->stage[s], len, mode, global_offset + begin, p->child[s], p->stream[s],
                nullptr);
    if (rc) return rc;
    ws->launches += 1;
  }
  for (uint64_t c = (nchunks >= 2 ? nchunks - 2 : 0); c < nchunks; ++c) {
    const int s = (int)(c & 1);
    AMM_CUDA(cudaStreamSynchronize(p->stream[s]));
    amm_record r;
    if (!read_host_slot(p->child[s], p->child[s]->seq, &r)) return AMM_ERR_CUDA;
    recs.push_back(r);
  }
  return AMM_OK;
}

extern "C" int amm_host_argminmax(int dtype, const void* host_ptr, uint64_t n, int mode,
                                  amm_workspace* ws, uint64_t out[2]) {
  if (!out) return AMM_ERR_ARG;
  std::vector<amm_record> recs;
  int rc = host_scan(dtype, host_ptr, n, mode, 0, ws, recs);
  if (rc) return rc;
  combine(recs.data(), (int)recs.size(), mode, out);
  return AMM_OK;
}

// Same pipeline, but returns the shard-level 64-byte record (keys + global indices) so
// that several host shards (ranks) can be combined with amm_combine_records.
extern "C" int amm_host_argminmax_record(int dtype, const void* host_ptr, uint64_t n, int mode,
                                         uint64_t global_offset, amm_workspace* ws,
                                         amm_record* out) {
  if (!out) return AMM_ERR_ARG;
  std::vector<amm_record> recs;
  int rc = host_scan(dtype, host_ptr, n, mode, global_offset, ws, recs);
  if (rc) return rc;
  *out = fold_records(recs.data(), (int)recs.size());
  return AMM_OK;
}
