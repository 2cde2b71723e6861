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

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>

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

struct amm_sharded;

// One launcher thread per device (devices 1..G-1; the caller's thread drives device 0).  A call
// on the fused peer path is G x (set device, launch, wait for the pinned slot): from ONE host
// thread the G launches alone cost G x ~3 us (35 us per call at 8 GPUs, profiles/r01c_scaling.md)
// although the kernels themselves overlap; with a thread per device they go out together.
// A worker spins for ~100 us after its last job (a tight loop of calls never sleeps), then parks
// on a condition variable.
struct ShardWorker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::atomic<uint64_t> go{0};     // job sequence number posted by the caller
  std::atomic<uint64_t> done{0};   // ... and completed by the worker
  std::atomic<bool> parked{false};
  std::atomic<bool> quit{false};
  amm_sharded* sh = nullptr;
  int g = 0;
  // job
  int dtype = 0, mode = 0;
  const void* ptr = nullptr;
  uint64_t n = 0, offset = 0;
  // result
  int rc = 0;
  uint64_t out[2] = {0, 0};
};

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
  ShardWorker* workers[kMaxGpus];  // [0] unused; nullptr = the caller's thread drives that device
  uint64_t job_seq;
};

using namespace amm;

extern "C" {

static void start_workers(amm_sharded* sh);

int amm_sharded_destroy(amm_sharded* sh) {
  if (!sh) return AMM_OK;
  DeviceGuard restore(sh->devices[0]);  // the per-device set_device calls below are undone on exit
  for (int g = 1; g < sh->ngpu; ++g) {
    ShardWorker* w = sh->workers[g];
    if (!w) continue;
    w->quit.store(true);
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->go.fetch_add(1);
    }
    w->cv.notify_one();
    if (w->th.joinable()) w->th.join();
    delete w;
    sh->workers[g] = nullptr;
  }
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
      if (ok) start_workers(sh);
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

// A peer-mode call in which rank `missing` could not launch: the kernels of the other devices wait
// for a record that will never be sent.  Release them by writing a poisoned record
// (AMM_REC_PEER_TIMEOUT, this call's tag) into rank `missing`'s slot of every other device's
// exchange buffer; their waits then return AMM_ERR_PEER and mark the exchange broken.
static void poison_as_missing_rank(amm_sharded* sh, int missing) {
  const uint64_t xseq = sh->ws[missing]->xseq + 1;  // the tag of the call that did not go out here
  const uint64_t bank = (xseq & 1ull) * (uint64_t)sh->ngpu;
  uint64_t slot[XCHG_SLOT_WORDS];
  memset(slot, 0, sizeof(slot));
  amm_record rec;
  memset(&rec, 0, sizeof(rec));
  rec.min_idx = rec.max_idx = rec.nan_idx = AMM_NO_INDEX;
  rec.flags = AMM_REC_PEER_TIMEOUT;
  memcpy(slot, &rec, sizeof(rec));
  slot[8] = mix_checksum(slot, PEER_SLOT_SEED ^ xseq);
  slot[9] = xseq;
  for (int a = 0; a < sh->ngpu; ++a) {
    if (a == missing || !sh->ws[a]->xbuf) continue;
    // legacy-default-stream copy (UVA: any current device): the per-device streams are
    // non-blocking, so this does not queue behind the spinning kernel
    cudaMemcpy(sh->ws[a]->xbuf + (bank + (uint64_t)missing) * XCHG_SLOT_WORDS, slot, sizeof(slot),
               cudaMemcpyHostToDevice);
  }
  (void)cudaGetLastError();
}

// one device's part of a peer-mode call: launch the fused kernel, wait for the global answer
static int peer_job(amm_sharded* sh, int g, int dtype, const void* ptr, uint64_t n, int mode,
                    uint64_t offset, uint64_t out[2]) {
  int rc = launch(dtype, ptr, n, mode, offset, sh->ws[g], sh->streams[g], nullptr, true);
  if (rc) {  // (a CUDA launch error; arguments were checked before anything was launched)
    poison_as_missing_rank(sh, g);
    return rc;
  }
  return amm_workspace_wait(sh->ws[g], sh->streams[g], mode, out);
}

static void worker_main(ShardWorker* w) {
  cudaSetDevice(w->sh->devices[w->g]);
  uint64_t seen = 0;
  for (;;) {
    int spins = 0;
    while (w->go.load() == seen) {
      if (++spins > 4000) {  // ~100 us of polling, then sleep until the next call
        std::unique_lock<std::mutex> lk(w->mu);
        w->parked.store(true);
        w->cv.wait(lk, [&] { return w->go.load() != seen; });
        w->parked.store(false);
        break;
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    if (w->quit.load()) return;
    seen = w->go.load();
    w->rc = peer_job(w->sh, w->g, w->dtype, w->ptr, w->n, w->mode, w->offset, w->out);
    w->done.store(seen);
  }
}

static void start_workers(amm_sharded* sh) {
  const char* env = getenv("AMM_SHARDED_THREADS");
  if (env && atoi(env) == 0) return;
  for (int g = 1; g < sh->ngpu; ++g) {
    ShardWorker* w = new (std::nothrow) ShardWorker();
    if (!w) return;
    w->sh = sh;
    w->g = g;
    sh->workers[g] = w;
    w->th = std::thread(worker_main, w);
  }
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
    // fused path: one kernel per device does scan + exchange + combine; devices 1..G-1 are
    // driven by their launcher threads, device 0 by this thread
    const uint64_t job = ++sh->job_seq;
    uint64_t off = shard_lens[0];
    int rcs[kMaxGpus];
    uint64_t res[kMaxGpus][2];
    memset(res, 0, sizeof(res));
    for (int g = 1; g < sh->ngpu; ++g) {
      ShardWorker* w = sh->workers[g];
      if (w) {
        w->dtype = dtype;
        w->mode = mode;
        w->ptr = dev_ptrs[g];
        w->n = shard_lens[g];
        w->offset = off;
        w->go.store(job);
        if (w->parked.load()) {
          std::lock_guard<std::mutex> lk(w->mu);
          w->cv.notify_one();
        }
      }
      off += shard_lens[g];
    }
    if (sh->workers[1] == nullptr && sh->ngpu > 1) {
      // no launcher threads (AMM_SHARDED_THREADS=0): launch everything, then wait for everything
      uint64_t o = 0;
      int failed = -1;
      for (int g = 0; g < sh->ngpu && failed < 0; ++g) {
        rcs[g] = launch(dtype, dev_ptrs[g], shard_lens[g], mode, o, sh->ws[g], sh->streams[g], nullptr, true);
        if (rcs[g]) {
          poison_as_missing_rank(sh, g);
          failed = g;
        }
        o += shard_lens[g];
      }
      for (int g = 0; g < sh->ngpu; ++g) {
        if (failed >= 0 && g >= failed) {
          if (g > failed) rcs[g] = AMM_ERR_PEER;
          continue;
        }
        rcs[g] = set_device(sh->devices[g]);
        if (rcs[g] == AMM_OK) rcs[g] = amm_workspace_wait(sh->ws[g], sh->streams[g], mode, res[g]);
      }
    } else {
      set_device(sh->devices[0]);
      rcs[0] = peer_job(sh, 0, dtype, dev_ptrs[0], shard_lens[0], mode, 0, res[0]);
      for (int g = 1; g < sh->ngpu; ++g) {
        ShardWorker* w = sh->workers[g];
        while (w->done.load() != job) {
#if defined(__x86_64__)
          __builtin_ia32_pause();
#endif
        }
        rcs[g] = w->rc;
        res[g][0] = w->out[0];
        res[g][1] = w->out[1];
      }
    }
    int status = AMM_OK;
    for (int g = 0; g < sh->ngpu; ++g) {
      if (rcs[g] != AMM_OK && (status == AMM_OK || status == AMM_ERR_PEER)) status = rcs[g];
      // every device must have folded the same table
      if (status == AMM_OK && (res[g][0] != res[0][0] || res[g][1] != res[0][1])) status = AMM_ERR_PEER;
    }
    if (status != AMM_OK) {
      // the exchange is out of step (or a device is in trouble): leave the peer path for good
      for (int g = 0; g < sh->ngpu; ++g) {
        if (set_device(sh->devices[g]) == AMM_OK) cudaStreamSynchronize(sh->streams[g]);
        sh->ws[g]->xbroken = 1;
      }
      (void)cudaGetLastError();
      sh->use_peer = false;
      if (!sh->use_nccl && sh->api.CommInitAll == nullptr && load_nccl(sh->api) &&
          sh->api.CommInitAll(sh->comms, sh->ngpu, sh->devices) == 0)
        sh->use_nccl = true;
      return status;
    }
    out[0] = res[0][0];
    out[1] = res[0][1];
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
