// host_slice.cu -- `&[T]` / Vec<T> in HOST memory (the step before the path,
// src/lib.rs:679-712): the slice is cut into chunks, chunk c is copied H2D on
// stream c%2 and scanned on the same stream by the ordinary device kernel with
// global_offset = chunk start, so the copy of chunk c+1 overlaps the scan of
// chunk c.  The per-chunk 64-byte records are combined on the host in chunk order
// (strict compares => first occurrence).  End to end this is PCIe-bound.
#include <cuda_runtime.h>
#include <stdlib.h>

#include <new>
#include <vector>

#include "amm_internal.h"

struct amm_host_pipeline {
  size_t chunk_bytes;
  void* stage[2];
  cudaStream_t stream[2];
  amm_workspace* child[2];
};

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
  (void)rc;
  if (!ws->pipe) {
    rc = pipeline_create(ws);
    if (rc) {
      host_pipeline_destroy(ws);
      return rc;
    }
  }
  amm_host_pipeline* p = ws->pipe;
  const size_t es = amm_dtype_size(dtype);
  const uint64_t chunk_elems = p->chunk_bytes / es;
  const uint64_t nchunks = (n + chunk_elems - 1) / chunk_elems;
  for (int i = 0; i < 2; ++i) {
    amm_workspace_set_tuning(p->child[i], ws->threads, ws->ctas_per_sm, ws->variant);
    p->child[i]->direct_bytes = ws->direct_bytes;
    p->child[i]->direct_auto = ws->direct_auto;
    p->child[i]->direct_packed = ws->direct_packed;
    p->child[i]->direct_vecs_per_thread = ws->direct_vecs_per_thread;
    p->child[i]->pdl = ws->pdl;
    p->child[i]->direct_threads = ws->direct_threads;
    p->child[i]->direct_ctas_per_sm = ws->direct_ctas_per_sm;
  }
  recs.clear();
  recs.reserve((size_t)nchunks);
  const uint8_t* src = (const uint8_t*)host_ptr;
  for (uint64_t c = 0; c < nchunks; ++c) {
    const int s = (int)(c & 1);
    if (c >= 2) {  // stage s is free once chunk c-2 has been scanned; harvest its record
      AMM_CUDA(cudaStreamSynchronize(p->stream[s]));
      amm_record r;
      if (!read_host_slot(p->child[s], p->child[s]->seq, &r)) return AMM_ERR_CUDA;
      recs.push_back(r);
    }
    const uint64_t begin = c * chunk_elems;
    const uint64_t len = (n - begin) < chunk_elems ? (n - begin) : chunk_elems;
    AMM_CUDA(cudaMemcpyAsync(p->stage[s], src + begin * es, len * es, cudaMemcpyHostToDevice,
                             p->stream[s]));
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
