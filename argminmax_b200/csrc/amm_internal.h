// amm_internal.h -- definitions shared by the translation units of libargminmax_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/argminmax_b200.h"
#include "record_slot.h"

namespace amm {
extern thread_local int g_last_cuda_error;
}

#define AMM_CUDA(expr)                                  \
  do {                                                  \
    cudaError_t _e = (expr);                            \
    if (_e != cudaSuccess) {                            \
      amm::g_last_cuda_error = (int)_e;                 \
      (void)cudaGetLastError();                         \
      return AMM_ERR_CUDA;                              \
    }                                                   \
  } while (0)

struct amm_host_pipeline;  // host_slice.cu

// 14 policy slots (8 ints + 3 floats x 2 modes) x load-shape variants
constexpr int AMM_POLICY_SLOTS = 14;
constexpr int AMM_MAX_VARIANTS = 32;

struct amm_workspace {
  int device;
  int sm_count;
  // streaming-kernel geometry; 0 / 0 / -1 = automatic (chosen per call from the input size)
  int threads;      // CTA size
  int ctas_per_sm;  // persistent CTAs per SM (capped by occupancy)
  int variant;      // load shape, see kVariants in scan_launch.cuh
  int eff_threads, eff_ctas_per_sm, eff_variant;  // what the last streaming launch used
  int max_grid;
  uint64_t direct_bytes;  // inputs up to this size take the one-pass latency kernel ...
  int direct_auto;        // ... unless 1: size rule per key width (see launch_variant)
  int direct_threads, direct_ctas_per_sm;
  int direct_vecs_per_thread;  // target vectors per thread when sizing the latency kernel's grid
  int direct_packed;      // 32-bit keys: packed-word atomics instead of tagged slots (AMM_DIRECT_PACKED=0: off)
  int direct_one_cta_vecs;  // inputs of up to this many 16-byte vectors (and 16 Ki elements) take ONE CTA
  int scan_packed;        // the same cross-CTA step in the streaming kernel (AMM_SCAN_PACKED=0: off)
  int scan_dyn_pct;       // share of the tiles claimed at run time: -1 automatic, 0 none, 1..90 percent (AMM_SCAN_DYN_PCT)
  int scan_ring;          // TMA-ring kernel: -1 automatic (inputs >= 2 GiB), 0 never, 1 always (AMM_SCAN_RING)
  int ring_ctas_per_sm;   // ... CTAs per SM (2: 2 x 96 KiB of stages; AMM_RING_CTAS)
  int ring_ready[AMM_POLICY_SLOTS][4];  // ... dynamic shared memory opted in (per instantiation)
  int scan_dyn_pipelined; // claimed tail also for pipelined scans (AMM_SCAN_DYN_PIPELINED=0: off)
  int scan_contig;        // tile partition of the streaming kernel: -1 automatic, 0 grid-stride, 1 contiguous ranges
  uint64_t direct_elems32, direct_elems64;  // direct_auto: latency kernel up to this many elements (32- / 64-bit keys)
  int occupancy[AMM_POLICY_SLOTS][AMM_MAX_VARIANTS];      // cached CTAs/SM ...
  int occupancy_threads[AMM_POLICY_SLOTS][AMM_MAX_VARIANTS];  // ... for this CTA size
  size_t static_smem[AMM_POLICY_SLOTS][AMM_MAX_VARIANTS];   // static smem of the kernel + 1 (0 = not queried)
  size_t dyn_smem_set[AMM_POLICY_SLOTS][AMM_MAX_VARIANTS];  // MaxDynamicSharedMemorySize set so far
  size_t smem_per_sm, smem_reserved_per_cta, smem_optin_per_cta;  // device limits (co-residency cap)
  void* partials;            // max_grid x 64 B tagged slots, zeroed at creation (grid_reduce)
  uint32_t* ticket;          // scratch block of WS_SCRATCH_BYTES: ticket, bounds-check errors, claim-counter
                             // ring, packed candidate words (layout: scan_kernel.cuh, WS_*)
  amm_record* rec_dev;       // device record slot
  amm_record* rec_host;      // pinned + mapped host slot: record + checksum word (128 B)
  amm_record* rec_host_dev;  // ... its device address
  uint64_t seq;              // calls issued on this workspace
  int publish_fence;
  int pdl;        // launch scans with programmatic stream serialization (AMM_PDL=0 disables)
  int pipelined;  // let a scan's streaming phase overlap the previous scan's tail
  int last_grid, last_threads;
  uint64_t launches;
  amm_host_pipeline* pipe;  // lazily created by amm_host_argminmax
  int win_staged;           // short windows take the TMA-staged kernel (AMM_WIN_STAGED=0: old path)
  int win_staged_max;       // ... for windows up to this many bytes (mean)
  int win_lane_bytes;       // ... with one lane per this many bytes of window (1..8 lanes)
  int win_fixed;            // 32-bit-geometry instantiation for fixed windows that are not a multiple of 16 B (AMM_WIN_FIXED=0: off)
  int win_fast;             // geometry-free instantiation for aligned fixed windows (AMM_WIN_FAST=0: off)
  int win_chunks_per_sm;    // ... about this many chunks per SM in total
  int win_chunked;          // few 1-64 MiB windows are cut into chunks (AMM_WIN_CHUNKED=0: off)
  void* win_scratch;        // records + bounds of the huge-window path (amm_argminmax_windows)
  size_t win_scratch_bytes;
  // fused cross-GPU exchange (amm_exchange_*): this rank's buffer, the peer pointer table
  int xworld, xrank;
  uint64_t* xbuf;            // cudaMalloc'ed, 2 parities x world slots x 128 B
  uint64_t** xpeers_dev;     // device array [world] of peer-mapped buffer pointers
  void* xopened[16];         // pointers obtained from cudaIpcOpenMemHandle (to close)
  uint64_t xseq;
  uint64_t xtimeout_ms;      // in-kernel wait for a peer's record (0 = for ever); amm_exchange_set_timeout
  int xbroken;               // a call timed out: the banks are out of step, the exchange must be re-created
  uint64_t* stamps;          // AMM_PHASE_STAMPS builds: max_grid x 8 device words (else nullptr)
};

namespace amm {
// Every public entry point leaves the caller's current CUDA device as it found it.
struct DeviceGuard {
  int orig = -1;  // the caller's current device
  int rc = AMM_OK;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&orig) != cudaSuccess) {
      orig = -1;
      rc = AMM_ERR_CUDA;
      return;
    }
    if (orig != device && cudaSetDevice(device) != cudaSuccess) rc = AMM_ERR_CUDA;
  }
  // restores the caller's device whatever the guarded code made current in between (functions
  // that walk over several devices call set_device() inside the guard)
  ~DeviceGuard() {
    int now = -1;
    if (orig >= 0 && cudaGetDevice(&now) == cudaSuccess && now != orig) cudaSetDevice(orig);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
int set_device(int device);
// argument checks of one scan (what `launch` would reject), without touching the device
int validate_scan(int dtype, const void* dev_ptr, uint64_t n, int mode, bool empty_allowed);
int launch(int dtype, const void* dev_ptr, uint64_t n, int mode, uint64_t global_offset,
           amm_workspace* ws, cudaStream_t stream, void* dev_record_out, bool exchange = false);
void exchange_destroy(amm_workspace* ws);
bool read_host_slot(const amm_workspace* ws, uint64_t want, amm_record* out);
void combine(const amm_record* recs, int count, int mode, uint64_t out[2]);
amm_record fold_records(const amm_record* recs, int count);
void host_pipeline_destroy(amm_workspace* ws);
}  // namespace amm
