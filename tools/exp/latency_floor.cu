// latency_floor.cu -- experiment: what is the floor of "launch one kernel, get 16 bytes back on the
// host" on this box?  Compares against the synchronous amm_argminmax call (8-12 us for n <= 1 Mi).
//   A  empty <<<1,32>>> kernel, result flag written to mapped pinned memory, host polls the flag
//   B  as A, launched through cudaLaunchKernelEx with the programmatic-stream-serialization attr
//   C  as A, host waits with cudaStreamSynchronize instead of polling
//   D  <<<256,256>>>: every CTA reads 16 KiB (4 MiB in total, L2-resident), warp+CTA reduce,
//      last-block-done ticket, flag -- the skeleton of the small-input scan without its bookkeeping
//   E  as D without the ticket: every CTA atomically folds its packed (value, index) into one word,
//      the last ticket holder only copies the word out
//   J* grid-step variants on the product's latency shapes (CTAs x 512 threads x vpt 16-byte vectors,
//      L2-resident): packed atomics + fence + ticket (what packed_grid_reduce does) against
//      (1) no fence, returning atomics; (2) one 16-byte {word, tag} slot per CTA, polled by warp 0 of
//      CTA 0; (3) the same polled by all of CTA 0; (4) one thread-block cluster, candidates through
//      distributed shared memory + cluster barrier; (5) fence.acq_rel.gpu instead of __threadfence
//      (= fence.sc); (6) the ordering folded into the ticket atomic (atom.acq_rel.gpu)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/bin/latency_floor tools/exp/latency_floor.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <vector>

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(1);                                                                   \
    }                                                                            \
  } while (0)

__global__ void k_empty(volatile uint64_t* flag, uint64_t seq) {
  if (threadIdx.x == 0) *flag = seq;
}

__global__ void __launch_bounds__(256) k_scan(const uint4* data, uint32_t* ticket, float* partial,
                                              volatile uint64_t* flag, uint64_t seq) {
  __shared__ float s[8];
  __shared__ bool last;
  float m = 3.0e38f;
  const uint4* p = data + (size_t)blockIdx.x * 1024 + threadIdx.x;
  uint4 v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = p[j * 256];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m = fminf(m, fminf(fminf(__uint_as_float(v[j].x), __uint_as_float(v[j].y)),
                       fminf(__uint_as_float(v[j].z), __uint_as_float(v[j].w))));
  }
  for (int o = 16; o; o >>= 1) m = fminf(m, __shfl_xor_sync(~0u, m, o));
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fminf(m, s[i]);
    partial[blockIdx.x] = m;
    __threadfence();
    last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  m = 3.0e38f;
  for (uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) m = fminf(m, __ldcg(&partial[i]));
  for (int o = 16; o; o >>= 1) m = fminf(m, __shfl_xor_sync(~0u, m, o));
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fminf(m, s[i]);
    partial[gridDim.x] = m;
    *flag = seq;
  }
}

__global__ void __launch_bounds__(256) k_scan_atomic(const uint4* data, uint32_t* ticket,
                                                     unsigned long long* word,
                                                     volatile uint64_t* flag, uint64_t seq) {
  __shared__ float s[8];
  float m = 3.0e38f;
  const uint4* p = data + (size_t)blockIdx.x * 1024 + threadIdx.x;
  uint4 v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = p[j * 256];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m = fminf(m, fminf(fminf(__uint_as_float(v[j].x), __uint_as_float(v[j].y)),
                       fminf(__uint_as_float(v[j].z), __uint_as_float(v[j].w))));
  }
  for (int o = 16; o; o >>= 1) m = fminf(m, __shfl_xor_sync(~0u, m, o));
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fminf(m, s[i]);
    const unsigned long long packed = ((unsigned long long)__float_as_uint(m) << 32) | blockIdx.x;
    atomicMin(word, packed);
    __threadfence();
    if (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1) {
      const unsigned long long r = atomicExch(word, ~0ull);  // read + re-arm
      *flag = seq + (r & 0);
    }
  }
}


namespace cg = cooperative_groups;

// ---- grid-step variants on the latency kernel's shape ------------------------------------------
enum { GS_PACKED = 0, GS_NOFENCE = 1, GS_POLL_WARP = 2, GS_POLL_CTA = 3, GS_CLUSTER = 4, GS_ACQREL_FENCE = 5, GS_ACQREL_ATOM = 6 };

__device__ __forceinline__ uint32_t cta_min_u32(uint32_t m, uint32_t* s /* [32] */) {
  m = __reduce_min_sync(~0u, m);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) s[warp] = m;
  __syncthreads();
  uint32_t r = lane < nw ? s[lane] : 0xFFFFFFFFu;
  return __reduce_min_sync(~0u, r);  // valid in every warp
}

template <int MODE>
__global__ void __launch_bounds__(512) k_step(const uint4* data, int vpt, uint32_t* ticket,
                                              unsigned long long* word, uint4* slots,
                                              volatile uint64_t* flag, uint64_t seq) {
  __shared__ uint32_t s[32];
  __shared__ unsigned long long s_cl[16];
  uint32_t m = 0xFFFFFFFFu;
  const uint4* p = data + (size_t)blockIdx.x * blockDim.x * vpt + threadIdx.x;
  uint4 v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j < vpt) v[j] = __ldcg(p + (size_t)j * blockDim.x);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j < vpt) m = min(m, min(min(v[j].x, v[j].y), min(v[j].z, v[j].w)));
  m = cta_min_u32(m, s);
  const unsigned long long packed = ((unsigned long long)m << 32) | blockIdx.x;
  const uint32_t tag = (uint32_t)seq;
  if (MODE == GS_ACQREL_FENCE || MODE == GS_ACQREL_ATOM) {
    // the same protocol with the weakest sufficient ordering: release before the ticket, acquire
    // after it (fence.acq_rel instead of __threadfence's fence.sc), or both folded into the atomic
    if (threadIdx.x == 0) {
      atomicMin(word, packed);
      uint32_t old;
      if (MODE == GS_ACQREL_FENCE) {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        old = atomicInc(ticket, gridDim.x - 1);
        if (old == gridDim.x - 1) asm volatile("fence.acq_rel.gpu;" ::: "memory");
      } else {
        asm volatile("atom.acq_rel.gpu.global.inc.u32 %0, [%1], %2;" : "=r"(old) : "l"(ticket), "r"(gridDim.x - 1) : "memory");
      }
      if (old == gridDim.x - 1) {
        const unsigned long long r = atomicExch(word, ~0ull);
        *flag = seq + (r & 0);
      }
    }
  } else if (MODE == GS_PACKED || MODE == GS_NOFENCE) {
    if (threadIdx.x == 0) {
      if (MODE == GS_PACKED) {
        atomicMin(word, packed);
        __threadfence();
      } else {
        const unsigned long long old = atomicMin(word, packed);  // returning: performed at L2 when it is back
        if (old == 1ull) return;                                 // (never; keeps the dependency)
      }
      if (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1) {
        const unsigned long long r = atomicExch(word, ~0ull);
        *flag = seq + (r & 0);
      }
    }
  } else if (MODE == GS_POLL_WARP || MODE == GS_POLL_CTA) {
    if (blockIdx.x != 0) {
      if (threadIdx.x == 0)
        asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(slots + blockIdx.x),
                     "r"((uint32_t)packed), "r"((uint32_t)(packed >> 32)), "r"(0u), "r"(tag)
                     : "memory");
      return;
    }
    unsigned long long best = packed;
    const uint32_t first = 1 + (MODE == GS_POLL_WARP ? (threadIdx.x & 31) : threadIdx.x);
    const uint32_t step = MODE == GS_POLL_WARP ? 32 : blockDim.x;
    if (MODE == GS_POLL_WARP && threadIdx.x >= 32) return;
    for (uint32_t b0 = first; b0 < gridDim.x; b0 += 4 * step) {
      uint4 r[4];
      bool ok;
      do {
        ok = true;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t b = b0 + j * step;
          if (b < gridDim.x) {
            asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(r[j].x), "=r"(r[j].y), "=r"(r[j].z), "=r"(r[j].w)
                         : "l"(slots + b)
                         : "memory");
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (b0 + j * step < gridDim.x) ok = ok && r[j].w == tag;
      } while (!ok);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (b0 + j * step < gridDim.x) {
          const unsigned long long w = ((unsigned long long)r[j].y << 32) | r[j].x;
          best = w < best ? w : best;
        }
    }
    uint32_t hi = (uint32_t)(best >> 32), lo = (uint32_t)best;
    if (MODE == GS_POLL_WARP) {
      const uint32_t mh = __reduce_min_sync(~0u, hi);
      const uint32_t ml = __reduce_min_sync(~0u, hi == mh ? lo : 0xFFFFFFFFu);
      if (threadIdx.x == 0) *flag = seq + (ml & 0);
    } else {
      __syncthreads();  // s[] is reused
      const uint32_t mh = cta_min_u32(hi, s);
      __syncthreads();
      const uint32_t ml = cta_min_u32(hi == mh ? lo : 0xFFFFFFFFu, s);
      if (threadIdx.x == 0) *flag = seq + (ml & 0);
    }
  } else {  // GS_CLUSTER: the grid is ONE cluster
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    if (threadIdx.x == 0) {
      unsigned long long* dst = cluster.map_shared_rank(s_cl, 0);
      dst[rank] = packed;
    }
    cluster.sync();
    if (rank == 0 && threadIdx.x < 32) {
      const unsigned long long w = threadIdx.x < cluster.num_blocks() ? s_cl[threadIdx.x] : ~0ull;
      const uint32_t hi = (uint32_t)(w >> 32), lo = (uint32_t)w;
      const uint32_t mh = __reduce_min_sync(~0u, hi);
      const uint32_t ml = __reduce_min_sync(~0u, hi == mh ? lo : 0xFFFFFFFFu);
      if (threadIdx.x == 0) *flag = seq + (ml & 0);
    }
  }
}

// publish variants: how much does the SHAPE of the result write to mapped host memory cost?
//   9 x 8-byte volatile stores (what publish_record did) vs 4 x 16-byte stores vs 2 x 32-byte
__global__ void k_publish(volatile uint64_t* slot, uint64_t seq, int shape) {
  if (threadIdx.x != 0) return;
  if (shape == 0) {
    for (int i = 0; i < 8; ++i) slot[i] = seq + 100 + i;
    slot[8] = seq;
  } else if (shape == 1) {
    uint32_t lo = (uint32_t)seq, hi = (uint32_t)(seq >> 32);
    for (int i = 0; i < 4; ++i) {
      const uint32_t tag_lo = i == 3 ? lo : lo + 7, tag_hi = hi;
      asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"((uint64_t*)slot + 2 * i),
                   "r"(lo + i), "r"(hi), "r"(tag_lo), "r"(tag_hi)
                   : "memory");
    }
  } else {
    uint32_t lo = (uint32_t)seq, hi = (uint32_t)(seq >> 32);
    for (int i = 0; i < 2; ++i) {
      const uint32_t tag_lo = i == 1 ? lo : lo + 7;
      asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"((uint64_t*)slot + 4 * i),
                   "r"(lo + i), "r"(hi), "r"(lo + 1), "r"(hi), "r"(lo + 2), "r"(hi), "r"(tag_lo), "r"(hi)
                   : "memory");
    }
  }
}

static uint64_t g_seq = 0;  // monotonic over all cases: a stale tag never matches

template <class F>
static void measure(const char* name, volatile uint64_t* hflag, bool poll, F launch) {
  std::vector<double> us;
  for (int i = 0; i < 2200; ++i) {
    const uint64_t seq = ++g_seq;
    const auto t0 = std::chrono::steady_clock::now();
    launch(seq);
    if (poll) {
      while (*hflag != seq) {
      }
    } else {
      CK(cudaStreamSynchronize(0));
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (i >= 200) us.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
  }
  CK(cudaDeviceSynchronize());
  std::sort(us.begin(), us.end());
  printf("{\"case\": \"%s\", \"us_median\": %.2f, \"us_min\": %.2f, \"us_p99\": %.2f}\n", name,
         us[us.size() / 2], us[0], us[us.size() * 99 / 100]);
  fflush(stdout);
}

int main() {
  uint64_t* hflag;
  CK(cudaHostAlloc((void**)&hflag, 64, cudaHostAllocMapped));
  *hflag = 0;
  uint64_t* dflag;
  CK(cudaHostGetDevicePointer((void**)&dflag, hflag, 0));
  uint4* data;
  CK(cudaMalloc(&data, 4 << 20));
  CK(cudaMemset(data, 0x3c, 4 << 20));
  uint32_t* ticket;
  CK(cudaMalloc(&ticket, 4));
  CK(cudaMemset(ticket, 0, 4));
  float* partial;
  CK(cudaMalloc(&partial, 4 * 2048));
  unsigned long long* word;
  CK(cudaMalloc(&word, 8));
  CK(cudaMemset(word, 0xff, 8));
  CK(cudaDeviceSynchronize());

  measure("A empty kernel + poll", hflag, true, [&](uint64_t s) { k_empty<<<1, 32>>>(dflag, s); });
  measure("B empty kernel (LaunchKernelEx + PDL attr) + poll", hflag, true, [&](uint64_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(32);
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, k_empty, (volatile uint64_t*)dflag, s));
  });
  measure("C empty kernel + cudaStreamSynchronize", hflag, false,
          [&](uint64_t s) { k_empty<<<1, 32>>>(dflag, s); });
  measure("D 4 MiB scan skeleton (256 CTAs, ticket) + poll", hflag, true,
          [&](uint64_t s) { k_scan<<<256, 256>>>(data, ticket, partial, dflag, s); });
  measure("E 4 MiB scan skeleton (256 CTAs, packed atomicMin) + poll", hflag, true,
          [&](uint64_t s) { k_scan_atomic<<<256, 256>>>(data, ticket, word, dflag, s); });
  measure("F 16 KiB one CTA + poll", hflag, true,
          [&](uint64_t s) { k_scan<<<1, 256>>>(data, ticket, partial, dflag, s); });
  // result-write shapes: the host polls the LAST word written (index 8 / 7 / 7)
  uint64_t* hslot;
  CK(cudaHostAlloc((void**)&hslot, 128, cudaHostAllocMapped));
  for (int i = 0; i < 16; ++i) hslot[i] = 0;
  uint64_t* dslot;
  CK(cudaHostGetDevicePointer((void**)&dslot, hslot, 0));
  measure("G publish 9 x 8 B, poll last word", hslot + 8, true,
          [&](uint64_t s) { k_publish<<<1, 32>>>(dslot, s, 0); });
  measure("H publish 4 x 16 B, poll last word", hslot + 7, true,
          [&](uint64_t s) { k_publish<<<1, 32>>>(dslot, s, 1); });
  measure("I publish 2 x 32 B, poll last word", hslot + 7, true,
          [&](uint64_t s) { k_publish<<<1, 32>>>(dslot, s, 2); });

  // ---- grid-step variants on the latency kernel's shapes (512-thread CTAs) ----
  uint4* slots;
  CK(cudaMalloc(&slots, 16 * 4096));
  CK(cudaMemset(slots, 0, 16 * 4096));
  CK(cudaFuncSetAttribute(k_step<GS_CLUSTER>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  struct Shape { int ctas, vpt; const char* what; };
  const Shape shapes[] = {{4, 4, "32 KiB"}, {16, 1, "128 KiB"}, {16, 4, "512 KiB/16"}, {64, 1, "512 KiB/64"},
                          {128, 1, "1 MiB"}, {128, 4, "4 MiB"}};
  char name[160];
  for (const Shape& sh : shapes) {
    snprintf(name, sizeof name, "J0 %s: %d CTAs x 512 x %d, packed atomic + fence + ticket", sh.what, sh.ctas, sh.vpt);
    measure(name, hflag, true, [&](uint64_t s) { k_step<GS_PACKED><<<sh.ctas, 512>>>(data, sh.vpt, ticket, word, slots, dflag, s); });
    snprintf(name, sizeof name, "J1 %s: %d CTAs x 512 x %d, returning atomic, no fence", sh.what, sh.ctas, sh.vpt);
    measure(name, hflag, true, [&](uint64_t s) { k_step<GS_NOFENCE><<<sh.ctas, 512>>>(data, sh.vpt, ticket, word, slots, dflag, s); });
    snprintf(name, sizeof name, "J2 %s: %d CTAs x 512 x %d, tagged slots polled by warp 0 of CTA 0", sh.what, sh.ctas, sh.vpt);
    measure(name, hflag, true, [&](uint64_t s) { k_step<GS_POLL_WARP><<<sh.ctas, 512>>>(data, sh.vpt, ticket, word, slots, dflag, s); });
    snprintf(name, sizeof name, "J3 %s: %d CTAs x 512 x %d, tagged slots polled by CTA 0", sh.what, sh.ctas, sh.vpt);
    measure(name, hflag, true, [&](uint64_t s) { k_step<GS_POLL_CTA><<<sh.ctas, 512>>>(data, sh.vpt, ticket, word, slots, dflag, s); });
    snprintf(name, sizeof name, "J5 %s: %d CTAs x 512 x %d, packed atomic + fence.acq_rel + ticket", sh.what, sh.ctas, sh.vpt);
    measure(name, hflag, true, [&](uint64_t s) { k_step<GS_ACQREL_FENCE><<<sh.ctas, 512>>>(data, sh.vpt, ticket, word, slots, dflag, s); });
    snprintf(name, sizeof name, "J6 %s: %d CTAs x 512 x %d, packed atomic + atom.acq_rel ticket", sh.what, sh.ctas, sh.vpt);
    measure(name, hflag, true, [&](uint64_t s) { k_step<GS_ACQREL_ATOM><<<sh.ctas, 512>>>(data, sh.vpt, ticket, word, slots, dflag, s); });
    if (sh.ctas <= 16) {
      snprintf(name, sizeof name, "J4 %s: one cluster of %d CTAs x 512 x %d, DSMEM + cluster barrier", sh.what, sh.ctas, sh.vpt);
      measure(name, hflag, true, [&](uint64_t s) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sh.ctas);
        cfg.blockDim = dim3(512);
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = sh.ctas;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, k_step<GS_CLUSTER>, (const uint4*)data, sh.vpt, ticket, word, slots, (volatile uint64_t*)dflag, s));
      });
    }
  }
  measure("F' 16 KiB one CTA + poll (again)", hflag, true,
          [&](uint64_t s) { k_scan<<<1, 256>>>(data, ticket, partial, dflag, s); });
  measure("A' empty kernel + poll (again)", hflag, true, [&](uint64_t s) { k_empty<<<1, 32>>>(dflag, s); });
  return 0;
}
