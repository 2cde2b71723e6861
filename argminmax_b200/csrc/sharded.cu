// sharded.cu -- ShardedDeviceSlice<T>: one logical array split into consecutive
// index ranges over the GPUs of one box, driven from ONE host thread.
//
// No reference sibling exists (the crate is single-threaded, single-device); the
// algebra is the reference's own merge rule applied across devices: argmin/argmax
// with first-occurrence tie-break is an associative reduction over (key, index)
// (chunk merge src/simd/generic.rs:513-520, prefix/tail merge src/simd/task.rs:149-228).
//
// Per call, default ("peer"): every device runs ONE kernel on its own stream -- the ordinary scan
// (global_offset = shard start) fused with the exchange of the G 64-byte records over NVLink
// (peer-mapped buffers, scan_kernel.cuh: exchange_and_publish_warp) and the fold in ascending
// shard order with strict compares; the host reads the global answer from the pinned slots.
// "nccl" (north_star's baseline, AMM_SHARDED_EXCHANGE=nccl, and the fallback without P2P): scan
// kernels, then ONE ncclAllGather (G x 64 B, latency-bound) enqueued on the same streams, the
// gathered table is read back from device 0 and combined on the host.  NCCL is resolved with
// dlopen at run time so the library has no link-time NCCL dependency and shares the copy already
// loaded by the host process (e.g. torch's).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "amm_internal.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;  // 0 == ncclSuccess
constexpr int kNcclUint8 = 1;

struct NcclApi {
  void* handle;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
};

bool load_nccl(NcclApi& api) {
  memset(&api, 0, sizeof(api));
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return false;
  api.CommInitAll = (decltype(api.CommInitAll))dlsym(api.handle, "ncclCommInitAll");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
  api.GroupStart = (decltype(api.GroupStart))dlsym(api.handle, "ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.handle, "ncclGroupEnd");
  api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
  return api.CommInitAll && api.CommDestroy && api.GroupStart && api.GroupEnd && api.AllGather;
}

constexpr int kMaxGpus = 16;

}  // namespace

struct amm_sharded {
  int ngpu;
  int devices[kMaxGpus];
  amm_workspace* ws[kMaxGpus];
  cudaStream_t streams[kMaxGpus];
  amm_record* gathered[kMaxGpus];  // device, ngpu records each
  amm_record* gathered_host;       // pinned, ngpu records
  bool use_nccl;
  bool use_peer;  // fused in-kernel exchange over peer-mapped memory (default when possible)
  NcclApi api;
  ncclComm_t comms[kMaxGpus];
};

using namespace amm;

extern "C" {

int amm_sharded_destroy(amm_sharded* sh) {
  if (!sh) return AMM_OK;
  DeviceGuard restore(sh->devices[0]);  // the per-device set_device calls below are undone on exit
  for (int g = 0; g < sh->ngpu; ++g) {
    set_device(sh->devices[g]);
    if (sh->streams[g]) {
      cudaStreamSynchronize(sh->streams[g]);
    }
    if (sh->use_nccl && sh->comms[g]) sh->api.CommDestroy(sh->comms[g]);
    if (sh->streams[g]) cudaStreamDestroy(sh->streams[g]);
    if (sh->gathered[g]) cudaFree(sh->gathered[g]);
    if (sh->ws[g]) amm_workspace_destroy(sh->ws[g]);
  }
  if (sh->gathered_host) cudaFreeHost(sh->gathered_host);
  delete sh;
  return AMM_OK;
}

int amm_sharded_create(int ngpu, const int* devices, amm_sharded** out) {
  if (!out || ngpu < 1 || ngpu > kMaxGpus) return AMM_ERR_ARG;
  *out = nullptr;
  amm_sharded* sh = new (std::nothrow) amm_sharded();
  if (!sh) return AMM_ERR_ARG;
  memset(sh, 0, sizeof(*sh));
  sh->ngpu = ngpu;
  for (int g = 0; g < ngpu; ++g) sh->devices[g] = devices ? devices[g] : g;
  DeviceGuard restore(sh->devices[0]);
  for (int g = 0; g < ngpu; ++g) {
    int rc = amm_workspace_create(sh->devices[g], &sh->ws[g]);
    if (rc) {
      amm_sharded_destroy(sh);
      return rc;
    }
    if (set_device(sh->devices[g]) != AMM_OK) {  // streams / buffers belong to device g
      amm_sharded_destroy(sh);
      return AMM_ERR_CUDA;
    }
    cudaError_t e = cudaStreamCreateWithFlags(&sh->streams[g], cudaStreamNonBlocking);
    if (e == cudaSuccess)
      e = cudaMalloc((void**)&sh->gathered[g], sizeof(amm_record) * (size_t)ngpu);
    if (e != cudaSuccess) {
      g_last_cuda_error = (int)e;
      amm_sharded_destroy(sh);
      return AMM_ERR_CUDA;
    }
  }
  cudaError_t e = cudaHostAlloc((void**)&sh->gathered_host, sizeof(amm_record) * (size_t)ngpu,
                                cudaHostAllocDefault);
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    amm_sharded_destroy(sh);
    return AMM_ERR_CUDA;
  }
  // Exchange mode: "peer" (default when every pair of devices is distinct and P2P-capable):
  // the scan kernel's last CTA writes its record into the peers' buffers -- no NCCL call, no
  // extra launch per call; "nccl": one ncclAllGather of the records; "host": pinned slots.
  const char* want = getenv("AMM_SHARDED_EXCHANGE");
  const bool allow_peer = !want || strcmp(want, "peer") == 0;
  const bool allow_nccl = !want || strcmp(want, "nccl") == 0 || strcmp(want, "peer") == 0;
  sh->use_peer = false;
  if (ngpu > 1 && allow_peer) {
    bool ok = true;
    for (int a = 0; a < ngpu && ok; ++a)
      for (int b = 0; b < ngpu && ok; ++b) {
        if (a == b) continue;
        if (sh->devices[a] == sh->devices[b]) ok = false;  // spin-waiting kernels need own GPUs
        int can = 0;
        if (ok && cudaDeviceCanAccessPeer(&can, sh->devices[a], sh->devices[b]) != cudaSuccess) ok = false;
        if (!can) ok = false;
      }
    if (ok) {
      for (int a = 0; a < ngpu && ok; ++a) {
        if (set_device(sh->devices[a])) ok = false;
        for (int b = 0; b < ngpu && ok; ++b) {
          if (a == b) continue;
          cudaError_t pe = cudaDeviceEnablePeerAccess(sh->devices[b], 0);
          if (pe == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
          else if (pe != cudaSuccess) ok = false;
        }
        if (ok && amm_exchange_create(sh->ws[a], ngpu, a, nullptr) != AMM_OK) ok = false;
      }
      void* bufs[kMaxGpus];
      for (int a = 0; a < ngpu && ok; ++a)
        if (amm_exchange_local_buffer(sh->ws[a], &bufs[a]) != AMM_OK) ok = false;
      for (int a = 0; a < ngpu && ok; ++a)
        if (amm_exchange_connect_ptrs(sh->ws[a], bufs) != AMM_OK) ok = false;
      (void)cudaGetLastError();
      sh->use_peer = ok;
    }
  }
  sh->use_nccl = false;
  if (ngpu > 1 && !sh->use_peer && allow_nccl && load_nccl(sh->api)) {
    if (sh->api.CommInitAll(sh->comms, ngpu, sh->devices) == 0) {
      sh->use_nccl = true;
    } else {
      memset(sh->comms, 0, sizeof(sh->comms));
    }
  }
  *out = sh;
  return AMM_OK;
}

int amm_sharded_uses_nccl(const amm_sharded* sh) { return (sh && sh->use_nccl) ? 1 : 0; }
int amm_sharded_uses_peer_exchange(const amm_sharded* sh) { return (sh && sh->use_peer) ? 1 : 0; }

int amm_sharded_workspace(amm_sharded* sh, int g, amm_workspace** ws) {
  if (!sh || !ws || g < 0 || g >= sh->ngpu) return AMM_ERR_ARG;
  *ws = sh->ws[g];
  return AMM_OK;
}

// A peer-mode call whose launch loop stopped at device `failed`: the kernels already running on
// devices [0, failed) wait for records that will never be sent.  Release them by writing a
// poisoned record (AMM_REC_PEER_TIMEOUT, this call's tag) for every missing rank into their
// exchange buffers, drain their streams, and take the handle off the peer path for good.
static void abort_peer_call(amm_sharded* sh, int failed) {
  for (int a = 0; a < failed; ++a) {
    amm_workspace* ws = sh->ws[a];
    if (set_device(sh->devices[a]) != AMM_OK) continue;
    const uint64_t xseq = ws->xseq;  // the tag the running kernel waits for
    const uint64_t bank = (xseq & 1ull) * (uint64_t)sh->ngpu;
    for (int r = failed; r < sh->ngpu; ++r) {
      uint64_t slot[XCHG_SLOT_WORDS];
      memset(slot, 0, sizeof(slot));
      amm_record rec;
      memset(&rec, 0, sizeof(rec));
      rec.min_idx = rec.max_idx = rec.nan_idx = AMM_NO_INDEX;
      rec.flags = AMM_REC_PEER_TIMEOUT;
      memcpy(slot, &rec, sizeof(rec));
      slot[8] = mix_checksum(slot, PEER_SLOT_SEED ^ xseq);
      slot[9] = xseq;
      // default-stream copy: the per-device streams are non-blocking, so this does not queue
      // behind the spinning kernel
      cudaMemcpy(ws->xbuf + (bank + (uint64_t)r) * XCHG_SLOT_WORDS, slot, sizeof(slot),
                 cudaMemcpyHostToDevice);
    }
  }
  for (int a = 0; a < failed; ++a) {
    if (set_device(sh->devices[a]) != AMM_OK) continue;
    cudaStreamSynchronize(sh->streams[a]);
    sh->ws[a]->xbroken = 1;
  }
  (void)cudaGetLastError();
  sh->use_peer = false;
  if (!sh->use_nccl && sh->api.CommInitAll == nullptr && load_nccl(sh->api) &&
      sh->api.CommInitAll(sh->comms, sh->ngpu, sh->devices) == 0)
    sh->use_nccl = true;
}

int amm_sharded_argminmax(amm_sharded* sh, int dtype, const void* const* dev_ptrs,
                          const uint64_t* shard_lens, int mode, uint64_t out[2]) {
  if (!sh || !dev_ptrs || !shard_lens || !out) return AMM_ERR_ARG;
  DeviceGuard restore(sh->devices[0]);
  // validate EVERY shard before anything is launched: a bad shard g must not leave devices
  // 0..g-1 running a collective that devices g.. never join
  uint64_t total = 0;
  for (int g = 0; g < sh->ngpu; ++g) {
    const int rc = validate_scan(dtype, dev_ptrs[g], shard_lens[g], mode, /*empty_allowed=*/true);
    if (rc) return rc;
    total += shard_lens[g];
  }
  if (total == 0) return AMM_ERR_EMPTY;
  if (sh->use_peer) {
    // fused path: one kernel per device does scan + exchange + combine
    uint64_t off = 0;
    for (int g = 0; g < sh->ngpu; ++g) {
      int rc = launch(dtype, dev_ptrs[g], shard_lens[g], mode, off, sh->ws[g], sh->streams[g],
                      nullptr, true);
      if (rc) {  // (a CUDA launch error; arguments were checked above)
        abort_peer_call(sh, g);
        return rc;
      }
      off += shard_lens[g];
    }
    uint64_t first[2] = {0, 0};
    int status = AMM_OK;
    for (int g = 0; g < sh->ngpu; ++g) {  // wait for ALL devices even if one reports an error
      uint64_t r[2] = {0, 0};
      int rc = set_device(sh->devices[g]);
      if (rc == AMM_OK) rc = amm_workspace_wait(sh->ws[g], sh->streams[g], mode, r);
      if (rc) {
        if (status == AMM_OK) status = rc;
        continue;
      }
      if (g == 0) {
        first[0] = r[0];
        first[1] = r[1];
      } else if (status == AMM_OK && (r[0] != first[0] || r[1] != first[1])) {
        status = AMM_ERR_PEER;  // every device must have folded the same table
      }
    }
    if (status != AMM_OK) {
      if (status == AMM_ERR_PEER) abort_peer_call(sh, 0);  // nothing left running: just leave the peer path
      return status;
    }
    out[0] = first[0];
    out[1] = first[1];
    return AMM_OK;
  }
  // 1. every device scans its shard (async, one stream per device)
  uint64_t offset = 0;
  for (int g = 0; g < sh->ngpu; ++g) {
    if (shard_lens[g] == 0) {  // an empty shard contributes a record without value / NaN
      int rc = set_device(sh->devices[g]);
      if (rc) return rc;
      AMM_CUDA(cudaMemsetAsync(sh->ws[g]->rec_dev, 0, sizeof(amm_record), sh->streams[g]));
    } else {
      int rc = launch(dtype, dev_ptrs[g], shard_lens[g], mode, offset, sh->ws[g], sh->streams[g],
                      nullptr);
      if (rc) return rc;
    }
    offset += shard_lens[g];
  }
  const amm_record* table = nullptr;
  amm_record local[kMaxGpus];
  if (sh->use_nccl) {
    // 2. one all-gather of the 64-byte records over NVLink, on the compute streams
    if (sh->api.GroupStart() != 0) return AMM_ERR_NCCL;
    for (int g = 0; g < sh->ngpu; ++g) {
      if (sh->api.AllGather(sh->ws[g]->rec_dev, sh->gathered[g], sizeof(amm_record), kNcclUint8,
                            sh->comms[g], sh->streams[g]) != 0) {
        sh->api.GroupEnd();
        return AMM_ERR_NCCL;
      }
    }
    if (sh->api.GroupEnd() != 0) return AMM_ERR_NCCL;
    int rc = set_device(sh->devices[0]);
    if (rc) return rc;
    AMM_CUDA(cudaMemcpyAsync(sh->gathered_host, sh->gathered[0],
                             sizeof(amm_record) * (size_t)sh->ngpu, cudaMemcpyDeviceToHost,
                             sh->streams[0]));
    for (int g = 0; g < sh->ngpu; ++g) {
      rc = set_device(sh->devices[g]);
      if (rc) return rc;
      AMM_CUDA(cudaStreamSynchronize(sh->streams[g]));
    }
    table = sh->gathered_host;
  } else {
    // no NCCL in this process: the records are read from each device's pinned slot
    for (int g = 0; g < sh->ngpu; ++g) {
      int rc = set_device(sh->devices[g]);
      if (rc) return rc;
      AMM_CUDA(cudaStreamSynchronize(sh->streams[g]));
      if (shard_lens[g] == 0)
        memset(&local[g], 0, sizeof(amm_record));
      else if (!read_host_slot(sh->ws[g], sh->ws[g]->seq, &local[g]))
        return AMM_ERR_CUDA;
    }
    table = local;
  }
  // 3. deterministic combine, ascending shard order
  combine(table, sh->ngpu, mode, out);
  return AMM_OK;
}

}  // extern "C"
