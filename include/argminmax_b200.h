/*
 * argminmax_b200.h -- C ABI of the B200-native fused argmin+argmax library.
 *
 * This is the drop-in boundary for ONE hot path of the reference crate
 * (jvdd/argminmax): `ArgMinMax::argminmax` / `NaNArgMinMax::nanargminmax` over a
 * contiguous 1-D array of one of 11 primitive dtypes, for data that already
 * lives in B200 HBM.  Each entry point below names the reference interface it
 * stands in for (paths relative to the reference repository root).  The Rust
 * side (`mod cuda`: DeviceSlice<T> / ShardedDeviceSlice<T>, see INTEGRATION.md
 * and rust/) binds exactly these symbols through `extern "C"`.
 *
 * Conventions
 *   - plain pointers and sizes only; no CUDA / torch types in any signature
 *     (`stream` is a cudaStream_t passed as void*, NULL = legacy default stream).
 *   - every function returns an amm_status (0 = ok).  The reference's signatures
 *     are infallible and PANIC on empty input (src/scalar/generic.rs:182,
 *     src/simd/task.rs:12); a binding turns any non-zero status into a panic.
 *   - indices are uint64_t element indices (Rust usize on 64-bit hosts).
 *   - the library only READS the caller's device memory; all scratch lives in the
 *     amm_workspace handle; there is no hidden global state and NO CPU fallback:
 *     without a usable CUDA device every compute entry point fails with
 *     AMM_ERR_CUDA.
 *   - one workspace must not be used by two calls at the same time; distinct
 *     workspaces / streams are independent.
 *
 * Semantics (bit-exact with the reference's scalar implementation,
 * src/scalar/generic.rs:181-217 and src/scalar/scalar_f16.rs:20-48,100-136):
 *   mode AMM_IGNORE_NAN (ArgMinMax):  ints: plain order.  floats: NaNs are
 *     skipped; all-NaN input -> (0,0); f32/f64 compare as IEEE values
 *     (-0.0 == +0.0), f16 through the reference's i16 "ord" key (-0.0 < +0.0).
 *   mode AMM_RETURN_NAN (NaNArgMinMax, floats only; ints ignore the mode): if
 *     any NaN is present both outputs are the index of the FIRST NaN.
 *   ties -> first occurrence, for min and for max, at every level.
 */
#ifndef ARGMINMAX_B200_H
#define ARGMINMAX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMM_VERSION_MAJOR 0
#define AMM_VERSION_MINOR 1

/* dtype codes: the eleven dtypes of README.md:22 / src/lib.rs:657-664 */
typedef enum amm_dtype {
  AMM_I8 = 0, AMM_I16 = 1, AMM_I32 = 2, AMM_I64 = 3,
  AMM_U8 = 4, AMM_U16 = 5, AMM_U32 = 6, AMM_U64 = 7,
  AMM_F16 = 8, AMM_F32 = 9, AMM_F64 = 10,
  AMM_DTYPE_COUNT = 11
} amm_dtype;

/* the dtype strategies of src/dtype_strategy.rs:6,15,24 collapsed to a run-time tag:
   Int + FloatIgnoreNaN -> AMM_IGNORE_NAN, FloatReturnNaN -> AMM_RETURN_NAN */
typedef enum amm_mode { AMM_IGNORE_NAN = 0, AMM_RETURN_NAN = 1 } amm_mode;

typedef enum amm_status {
  AMM_OK = 0,
  AMM_ERR_EMPTY = 1,    /* n == 0: the reference asserts !arr.is_empty()            */
  AMM_ERR_ARG = 2,      /* NULL pointer, unknown dtype / mode, bad handle           */
  AMM_ERR_ALIGN = 3,    /* device pointer not aligned to sizeof(element)            */
  AMM_ERR_CUDA = 4,     /* a CUDA call failed (amm_last_cuda_error() has the code)  */
  AMM_ERR_NCCL = 5,     /* NCCL unavailable or a collective failed                  */
  AMM_ERR_TOO_LARGE = 6, /* more than 2^32-2 scan tiles (>= 32 TiB of input)        */
  AMM_ERR_PEER = 7       /* fused peer exchange: a rank never delivered its record  */
} amm_status;

/*
 * Per-device result record (64 bytes).  One record is what each GPU contributes
 * to the cross-GPU combine (ShardedDeviceSlice): keys are the order-preserving
 * signed-integer images of the element values (the reference's key transforms,
 * src/simd/simd_u32.rs:56, src/simd/simd_f32_return_nan.rs:5-9,
 * src/scalar/scalar_f16.rs:10-13) widened to 64 bit so that one combine routine
 * serves every dtype.  Indices are GLOBAL (shard offset already added).
 */
typedef struct amm_record {
  int64_t min_key;
  uint64_t min_idx;
  int64_t max_key;
  uint64_t max_idx;
  uint64_t nan_idx; /* index of the first NaN, UINT64_MAX if none seen           */
  uint64_t flags;   /* AMM_REC_HAS_VALUE | AMM_REC_HAS_NAN                        */
  uint64_t n;       /* elements scanned for this record                          */
  uint64_t seq;     /* workspace call counter                                    */
} amm_record;
#define AMM_REC_HAS_VALUE 1ull /* at least one non-NaN element (always set for ints) */
#define AMM_REC_HAS_NAN 2ull
#define AMM_NO_INDEX UINT64_MAX

typedef struct amm_workspace amm_workspace; /* per-device scratch + pinned result slot */
typedef struct amm_sharded amm_sharded;     /* one workspace + stream per GPU + NCCL comms */

/* ------------------------------------------------------------------ lifecycle */
/* Scratch for calls on CUDA device `device` (per-CTA result slots, ticket and packed candidate
   words, 64-B device record, 64-B pinned+mapped host record).  No reference analogue:
   the CPU path needs no scratch. */
int amm_workspace_create(int device, amm_workspace** ws);
int amm_workspace_destroy(amm_workspace* ws);

/* -------------------------------------------------- the path, single device */
/*
 * Fused argmin+argmax over `n` elements of `dtype` at DEVICE pointer `dev_ptr`.
 * Replaces  <&[T] as ArgMinMax>::argminmax     src/lib.rs:272-412 (ints), :422-459 (floats)
 *      and  <&[T] as NaNArgMinMax>::nanargminmax  src/lib.rs:541-576
 * i.e. the whole SIMD backend below them (src/simd/generic.rs:670-786 ->
 * src/simd/task.rs:4-55 -> src/simd/generic.rs:362-412, 487-543).
 * Synchronous from the caller's view: launches on `stream`, waits, writes
 * out[0] = argmin, out[1] = argmax.
 */
int amm_argminmax(int dtype, const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws,
                  void* stream, uint64_t out[2]);

/* dtype-suffixed symbols, one per `impl ArgMinMax for &[T]` instantiation
   (src/lib.rs:657 ints, :660 f32/f64, :664 f16) */
int amm_argminmax_i8(const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws, void* stream, uint64_t out[2]);
int amm_argminmax_i16(const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws, void* stream, uint64_t out[2]);
int amm_argminmax_i32(const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws, void* stream, uint64_t out[2]);
int amm_argminmax_i64(const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws, void* stream, uint64_t out[2]);
int amm_argminmax_u8(const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws, void* stream, uint64_t out[2]);
int amm_argminmax_u16(const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws, void* stream, uint64_t out[2]);
int amm_argminmax_u32(const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws, void* stream, uint64_t out[2]);
int amm_argminmax_u64(const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws, void* stream, uint64_t out[2]);
int amm_argminmax_f16(const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws, void* stream, uint64_t out[2]);
int amm_argminmax_f32(const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws, void* stream, uint64_t out[2]);
int amm_argminmax_f64(const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws, void* stream, uint64_t out[2]);

/*
 * Single-output forms: ArgMinMax::argmin / argmax (src/lib.rs:170,185) and
 * NaNArgMinMax::nanargmin / nanargmax (src/lib.rs:227,243).  The scan is
 * bandwidth-bound, so these run the fused kernel and return one half -- which is
 * literally how the reference defines them for the return-NaN strategy
 * (src/simd/simd_f32_return_nan.rs:407-413).
 */
int amm_argmin(int dtype, const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws,
               void* stream, uint64_t* out);
int amm_argmax(int dtype, const void* dev_ptr, uint64_t n, int mode, amm_workspace* ws,
               void* stream, uint64_t* out);

/* ------------------------------------------ split-phase (async) single device */
/*
 * Enqueue the scan on `stream` and return immediately.  The 64-byte record lands
 * in the workspace's device record (amm_workspace_device_record) and in its pinned
 * host record; `global_offset` is added to every index (shard offset for
 * ShardedDeviceSlice, 0 otherwise).  amm_workspace_wait() synchronises `stream`
 * and turns the record into (argmin, argmax).  Used by the sharded driver, by
 * torch.distributed ranks that all-gather records themselves, and by bench.py to
 * time the kernel alone with events.
 */
int amm_argminmax_launch(int dtype, const void* dev_ptr, uint64_t n, int mode,
                         uint64_t global_offset, amm_workspace* ws, void* stream,
                         void* dev_record_out /* 64-B device buffer, NULL = workspace slot */);
int amm_workspace_wait(amm_workspace* ws, void* stream, int mode, uint64_t out[2]);
int amm_workspace_device_record(amm_workspace* ws, void** dev_record /* amm_record* on device */);
int amm_workspace_host_record(amm_workspace* ws, const amm_record** host_record);

/*
 * Deterministic first-index combine of `count` per-shard records given in
 * ascending shard order (host memory).  Strict compares, so the lowest shard wins
 * ties = global first occurrence -- the same rule as the reference's chunk merge
 * (src/simd/generic.rs:513-520) and prefix/tail merge (src/simd/task.rs:149-228).
 * Return-NaN: min over nan_idx wins.  No shard with a value (all NaN) -> (0,0).
 */
int amm_combine_records(const amm_record* records, int count, int mode, uint64_t out[2]);
/* The same fold as a one-thread kernel on `stream`: reads `count` records from DEVICE memory
   (e.g. the output of an NCCL all-gather) and publishes the folded record -- whose indices
   are the final answer -- to the workspace's device + pinned host slots, so that a
   multi-GPU step (scan -> all-gather -> combine) stays asynchronous.  amm_workspace_wait()
   then yields (argmin, argmax). */
int amm_combine_records_launch(const void* dev_records, int count, int mode, amm_workspace* ws,
                               void* stream);

/* ------------------------------- fused scan + cross-GPU exchange (one rank per GPU) */
/*
 * One kernel per rank does the scan AND the collective: warp 0 of its last CTA stores the rank's
 * 64-byte record into every peer's exchange buffer over NVLink (peer-mapped memory), waits
 * until all `world` records of the call have landed in its own buffer, folds them (ascending
 * rank order, strict compares) and publishes the GLOBAL record; amm_workspace_wait() then
 * returns the global (argmin, argmax) on every rank.  No NCCL call and no second launch sits
 * on the per-call path.  Setup, once per workspace (a workspace holds at most ONE exchange;
 * creating a second one fails with AMM_ERR_ARG until amm_exchange_destroy):
 *   amm_exchange_create   allocates this rank's buffer; `ipc_handle_out` (64 bytes, may be
 *                         NULL in a single process) receives its cudaIpcMemHandle_t;
 *   amm_exchange_connect_ipc   other processes: all ranks' handles (world x 64 B, rank order),
 *                         e.g. gathered once with torch.distributed / MPI; `device_ids`
 *                         (world x 16 bytes, rank order, may be NULL) are the ranks' GPU UUIDs
 *                         (amm_device_uuid): two ranks on one GPU are refused (AMM_ERR_ARG),
 *                         because their kernels would spin-wait for each other on one device;
 *   amm_exchange_connect_ptrs  same process: the peers' buffer pointers (peer access enabled).
 * Every rank must issue the same sequence of amm_argminmax_launch_exchange calls (it is a
 * collective); a shard may be empty (n == 0).
 * Failure is collective: a rank that waits longer than the timeout (amm_exchange_set_timeout,
 * default 60 000 ms, AMM_PEER_TIMEOUT_MS; 0 = wait for ever like NCCL) for a peer's record
 * poisons its own slot in every peer's buffer, so late ranks fail with AMM_ERR_PEER as well
 * instead of completing on their own.  After AMM_ERR_PEER the exchange is unusable (every
 * further launch returns AMM_ERR_PEER) until it is destroyed and created again on all ranks.
 * Teardown across processes: every rank calls amm_exchange_disconnect (closes the imported
 * handles), then -- after a barrier of the caller's -- amm_exchange_destroy (frees its own
 * buffer); amm_workspace_destroy does both for a single process.
 */
int amm_exchange_create(amm_workspace* ws, int world, int rank, void* ipc_handle_out);
int amm_exchange_connect_ipc(amm_workspace* ws, const void* ipc_handles, const void* device_ids);
int amm_exchange_connect_ptrs(amm_workspace* ws, void* const* peer_buffers);
int amm_exchange_local_buffer(amm_workspace* ws, void** dev_ptr);
int amm_exchange_set_timeout(amm_workspace* ws, uint64_t milliseconds);
int amm_exchange_disconnect(amm_workspace* ws);
int amm_exchange_destroy(amm_workspace* ws);
int amm_device_uuid(int device, void* uuid_out /* 16 bytes */);
int amm_argminmax_launch_exchange(int dtype, const void* dev_ptr, uint64_t n, int mode,
                                  uint64_t global_offset, amm_workspace* ws, void* stream);

/* -------------------------------------------------- host slices (step before) */
/*
 * `&[T]` / Vec<T> in HOST memory (src/lib.rs:679-712): chunked, double-buffered
 * H2D copies overlapped with the scan of the previous chunk; PCIe-bound.  `host_ptr`
 * should be pinned for full speed (pageable works).  Staging buffers are allocated
 * lazily inside the workspace.
 */
int amm_host_argminmax(int dtype, const void* host_ptr, uint64_t n, int mode, amm_workspace* ws,
                       uint64_t out[2]);
/* Same, but yields the 64-byte record (indices offset by `global_offset`) so that host
   shards held by several ranks can be combined with amm_combine_records. */
int amm_host_argminmax_record(int dtype, const void* host_ptr, uint64_t n, int mode,
                              uint64_t global_offset, amm_workspace* ws, amm_record* out);

/* ---------------------------------------------- one array over several GPUs */
/*
 * ShardedDeviceSlice<T>: one process drives `ngpu` devices of one box.  Shard g
 * is `shard_lens[g]` elements at `dev_ptrs[g]` on `devices[g]`; shards are
 * consecutive index ranges of one logical array (an empty shard is allowed).  All shards are
 * validated before anything is launched.  Data path, chosen at create time:
 *   "peer" (default when the devices are distinct and mutually P2P-capable): ONE kernel per
 *       device = scan + in-kernel exchange of the 64-byte records over NVLink + fold
 *       (amm_exchange_*, see above); no NCCL call, no second launch;
 *   "nccl" (north_star's baseline; AMM_SHARDED_EXCHANGE=nccl, and the fallback when peer
 *       mapping is unavailable): scan kernel -> ONE ncclAllGather of the records on the compute
 *       streams -> host fold as in amm_combine_records;
 *   "host" (no NCCL in the process): the records are read from each device's pinned slot.
 * If a launch fails half way through the devices of a peer-mode call, the kernels already
 * running are released (their missing peers' slots are filled with poisoned records), the
 * call returns the error and the handle continues on the nccl / host path.
 */
int amm_sharded_create(int ngpu, const int* devices, amm_sharded** sh);
int amm_sharded_destroy(amm_sharded* sh);
int amm_sharded_argminmax(amm_sharded* sh, int dtype, const void* const* dev_ptrs,
                          const uint64_t* shard_lens, int mode, uint64_t out[2]);
/* 1 if the records travel through NCCL, 0 if NCCL could not be loaded and the
   records are read from each device's pinned host record instead */
int amm_sharded_uses_nccl(const amm_sharded* sh);
/* 1 if the fused in-kernel peer exchange (amm_exchange_*) is used instead: chosen when all
   devices are distinct and mutually P2P-capable; AMM_SHARDED_EXCHANGE=nccl|host overrides */
int amm_sharded_uses_peer_exchange(const amm_sharded* sh);
/* the per-device workspace of shard g (e.g. to apply amm_workspace_set_tuning) */
int amm_sharded_workspace(amm_sharded* sh, int g, amm_workspace** ws);

/* ------------------------------------------- windowed / segmented form ("next" row) */
/*
 * Many sub-slices of one device array in ONE launch.  No reference entry point exists; the
 * semantics are the reference's, composed: for every window w,
 *     (dev_out[2w], dev_out[2w+1]) = arr[b_w .. e_w].argminmax() + b_w      (mode as above)
 * which is what the crate's down-sampling users compute by calling `argminmax` once per bin.
 * amm_argminmax_windows: b_w = dev_offsets[w], e_w = dev_offsets[w+1] (n_windows + 1 ascending
 * offsets in DEVICE memory, clamped to n).  amm_argminmax_fixed_windows: b_w = w * window,
 * ceil(n / window) windows, the last possibly short.  An empty window yields AMM_NO_INDEX
 * twice.  Asynchronous: results land in DEVICE memory (2 * n_windows uint64) in stream order.
 */
int amm_argminmax_windows(int dtype, const void* dev_ptr, uint64_t n, const uint64_t* dev_offsets,
                          uint64_t n_windows, int mode, amm_workspace* ws, void* stream,
                          uint64_t* dev_out);
int amm_argminmax_fixed_windows(int dtype, const void* dev_ptr, uint64_t n, uint64_t window,
                                int mode, amm_workspace* ws, void* stream, uint64_t* dev_out);

/* ------------------------------------------------------ device memory helpers */
/* What `DeviceSlice::<T>::from_slice(&[T])` needs so that a binding does not have to link the
   CUDA runtime itself: allocate / free HBM on `device`, blocking copies in both directions. */
int amm_device_count(int* count);
int amm_device_alloc(int device, size_t bytes, void** dev_ptr);
int amm_device_free(int device, void* dev_ptr);
int amm_copy_to_device(int device, void* dst_dev, const void* src_host, size_t bytes);
int amm_copy_to_host(int device, void* dst_host, const void* src_dev, size_t bytes);

/* ------------------------------------------------------------- diagnostics */
const char* amm_status_string(int status);
int amm_last_cuda_error(void);        /* cudaError_t of the last AMM_ERR_CUDA on this thread */
size_t amm_dtype_size(int dtype);     /* bytes per element, 0 for an unknown dtype */
int amm_version(void);                /* AMM_VERSION_MAJOR * 100 + AMM_VERSION_MINOR */

/* Scan geometry knobs (benchmarks / tuning only).  By default the library picks the geometry
   per call from the input size.  threads: CTA size (multiple of 32, 64..512); ctas_per_sm:
   persistent CTAs per SM; variant: load shape, see amm_variant_name() (the last one is the TMA-ring
   kernel, which has its own fixed geometry).  threads / ctas_per_sm <= 0 and variant < 0 select
   "automatic" for that knob: 256-bit loads, 256 threads, 4 CTAs per SM and a claimed tail of the
   tile sequence; the TMA ring from 2 GiB.  get_tuning reports what
   the most recent streaming launch actually used. */
int amm_workspace_set_tuning(amm_workspace* ws, int threads, int ctas_per_sm, int variant);
/* Pipelined mode for back-to-back scans on one stream (many resident arrays, repeated scans):
   the streaming phase of scan k+1 may start while scan k's last CTA is still folding /
   publishing (programmatic dependent launch).  The caller promises that the input of a scan is
   NOT produced by the kernel enqueued immediately before it on the same stream (memcpys and
   earlier kernels are fine).  Off by default; results and their order are unchanged.  In the
   default mode a scan waits for the preceding kernel to complete before its first load and
   only then lets its own dependents start, so a producer kernel that itself uses programmatic
   dependent launch (and triggers early) is safe. */
int amm_workspace_set_pipelined(amm_workspace* ws, int on);
/* Small inputs take the one-trip latency kernel (every thread holds its 1-4 vectors in registers:
   value-only fold, then the first matching element is located in registers; nothing is re-read)
   instead of the streaming kernel: by default up to 2^20 elements for 1/2/4-byte dtypes and 2^19
   for 8-byte dtypes (the measured crossovers), and never more than four 16-byte vectors per
   thread.  This call replaces the size rule by "inputs of at most `bytes`"; 0 disables the
   latency kernel. */
int amm_workspace_set_direct_threshold(amm_workspace* ws, uint64_t bytes);
int amm_workspace_get_tuning(const amm_workspace* ws, int* threads, int* ctas_per_sm, int* variant,
                             int* sm_count);
int amm_variant_count(void);
const char* amm_variant_name(int variant);
/* Host wall-clock latency of the synchronous amm_argminmax call, measured inside the library
   (no binding overhead): out_us = {median, min, p99} microseconds over `calls` calls. */
int amm_measure_call_latency(int dtype, const void* dev_ptr, uint64_t n, int mode,
                             amm_workspace* ws, void* stream, int calls, double out_us[3]);
/* The same for the collective call of the fused exchange (amm_argminmax_launch_exchange +
   amm_workspace_wait); every rank must call it with the same `calls`. */
int amm_measure_exchange_latency(int dtype, const void* dev_ptr, uint64_t n, int mode,
                                 uint64_t global_offset, amm_workspace* ws, void* stream, int calls,
                                 double out_us[3]);
/* Test hook (pure host logic, needs no GPU): the chunk plan the windowed form uses for few
   large adjacent windows.  chunk k = [cuts[k], cuts[k+1]); window w owns chunks
   [first[w], first[w+1]).  `first_out` holds n_windows + 1 entries.  Returns 1000 when the
   windows are not adjacent. */
int amm_debug_plan_window_chunks(const uint64_t* begins, const uint64_t* ends, uint64_t n_windows,
                                 size_t elem_size, int sm_count, int chunks_per_sm,
                                 uint64_t* cuts_out, uint64_t cuts_capacity, uint64_t* n_cuts,
                                 uint64_t* first_out);
/* Test hook (pure host logic, needs no GPU): the tile plan of the streaming kernel for an array
   of `n` elements of `elem_size` bytes at address `base_addr`, vectors of `vec_bytes`, `u` vectors
   per thread per tile, CTAs of `threads`, at most `grid_cap` CTAs.  out[0..11] = head_elems, n_vec,
   tail_start, per_vecs, grid, n_full_tiles, rem_vecs, rem_owner, dyn_tiles, dyn_koff, kind
   (0 grid-stride, 1 contiguous ranges, 2 grid-stride + claimed tail), status. */
int amm_debug_plan_scan(uint64_t base_addr, uint64_t n, int elem_size, int vec_bytes, int u, int threads,
                        uint64_t grid_cap, int sm_count, int pipelined, int scan_contig, int scan_dyn_pct,
                        uint64_t out[12]);
/* The same for the one-trip latency kernel (scan_launch.cuh: plan_direct; tests/test_direct_plan.py):
   out[0..6] = head_elems, n_vec, tail_start, grid, threads, vectors per thread, fits. */
int amm_debug_plan_direct(uint64_t base_addr, uint64_t n, int elem_size, int sm_count, int ctas_per_sm,
                          int threads, int min_vpt, int one_cta_vecs, uint64_t out[8]);
/* Debug twin built with -DAMM_PHASE_STAMPS only (tools/phase_stamps.py): %globaltimer stamps of
   the most recent scan, 8 words per CTA (entry, first tile folded, streaming done, CTA reduced,
   grid step entered; finishing CTA: grid folded, exact finish done, published).  The product library
   returns AMM_ERR_ARG. */
int amm_debug_read_stamps(amm_workspace* ws, uint64_t* out, int ctas);
/* grid size and kernels launched by the most recent amm_argminmax_launch on ws */
int amm_workspace_last_launch(const amm_workspace* ws, int* grid, int* threads, uint64_t* launches_total);

#ifdef __cplusplus
}
#endif
#endif /* ARGMINMAX_B200_H */
