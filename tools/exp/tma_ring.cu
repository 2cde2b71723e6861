// tma_ring.cu -- experiment: does feeding the big-input scan through a
// cp.async.bulk -> shared-memory -> mbarrier ring beat the register ping-pong of the product
// kernel?  Bytes in flight are then bounded by shared memory (up to ~190 KiB per SM) instead of
// registers (128 KiB per SM at 4 CTAs x 256 threads x 128 B), the memory system sees 16 KiB
// sequential requests, and one producer thread per CTA can CLAIM tiles dynamically without a CTA
// barrier (the tile id travels through the ring) -- the hand-off variant of dynamic tile claims.
//
// Kernels (all: value-only f32 min/max fold, result sunk; read-only stream):
//   mode 0  register ping-pong, static grid-stride           (the product's loop, v32x2_pf)
//   mode 1  TMA ring, static grid-stride tiles               (isolates the load path)
//   mode 2  TMA ring, tiles claimed by the producer thread   (load path + dynamic schedule)
// Usage: tma_ring <mode> <MiB> <ctas_per_sm> <stages> <tile_KiB> <iters>
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/bin/tma_ring tools/exp/tma_ring.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(1);                                                                   \
    }                                                                            \
  } while (0)

constexpr int THREADS = 256;
constexpr int MAX_STAGES = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void ld8(const uint8_t* p, uint32_t (&w)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p));
}
__device__ __forceinline__ void lds4(uint32_t addr, uint32_t (&w)[4]) {
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr));
}

struct Args {
  const uint8_t* base;
  uint32_t n_tiles;
  uint32_t tile_bytes;
  uint32_t stages;
  uint32_t* counter;
  float* sink;
};

// ---------------------------------------------------------------- mode 0: register ping-pong
__global__ void __launch_bounds__(THREADS) reg_kernel(const Args a) {
  constexpr int U = 2;
  const uint64_t tile_bytes = (uint64_t)THREADS * U * 32;
  const uint32_t nt = a.n_tiles;
  const uint8_t* lane_base = a.base + (uint64_t)threadIdx.x * 32;
  const uint64_t js = (uint64_t)THREADS * 32;
  float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
  uint32_t va[U][8], vb[U][8];
  uint32_t t = blockIdx.x;
  auto fold = [&](const uint32_t (&v)[U][8]) {
#pragma unroll
    for (int j = 0; j < U; ++j)
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const float x = __uint_as_float(v[j][w]);
        mn = fminf(mn, x);
        mx = fmaxf(mx, x);
      }
  };
  if (t < nt) {
#pragma unroll
    for (int j = 0; j < U; ++j) ld8(lane_base + (uint64_t)t * tile_bytes + j * js, va[j]);
  }
  while (t < nt) {
    const uint32_t t1 = t + gridDim.x;
    if (t1 < nt) {
#pragma unroll
      for (int j = 0; j < U; ++j) ld8(lane_base + (uint64_t)t1 * tile_bytes + j * js, vb[j]);
    }
    fold(va);
    if (t1 >= nt) break;
    const uint32_t t2 = t1 + gridDim.x;
    if (t2 < nt) {
#pragma unroll
      for (int j = 0; j < U; ++j) ld8(lane_base + (uint64_t)t2 * tile_bytes + j * js, va[j]);
    }
    fold(vb);
    t = t2;
  }
  if (mn == 12345.678f && mx == 0.1f) a.sink[0] = mn + mx;  // never true: keeps the fold alive
}

// ---------------------------------------------------------------- modes 1 / 2: TMA ring
// Warp 0 lane 0 is the producer; all 256 threads consume (thread i reads bytes [i*16 + k*4096)).
template <bool DYNAMIC>
__global__ void __launch_bounds__(THREADS) ring_kernel(const Args a) {
  extern __shared__ __align__(128) uint8_t dsm[];
  __shared__ __align__(8) uint64_t full[MAX_STAGES], empty[MAX_STAGES];
  __shared__ uint32_t tile_of[MAX_STAGES];
  const uint32_t S = a.stages, TB = a.tile_bytes, nt = a.n_tiles;
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
  // producer state (thread 0 only)
  uint32_t p_next = DYNAMIC ? 0 : blockIdx.x;  // next tile to issue (static) / unused (dynamic)
  uint32_t p_issued = 0;                       // tiles issued so far by this CTA
  bool p_done = false;
  auto produce = [&](uint32_t upto) {  // issue until `upto` tiles are in flight beyond consumption
    while (!p_done && p_issued < upto) {
      uint32_t t;
      if (DYNAMIC) {
        t = atomicAdd(a.counter, 1u);
      } else {
        t = p_next;
        p_next += gridDim.x;
      }
      const uint32_t s = p_issued % S;
      if (p_issued >= S) mbar_wait(&empty[s], ((p_issued / S) - 1) & 1);  // stage free again
      if (t >= nt) {
        tile_of[s] = 0xFFFFFFFFu;  // end marker
        mbar_arrive(&full[s]);
        p_done = true;
      } else {
        tile_of[s] = t;
        mbar_expect_tx(&full[s], TB);
        tma_g2s(dsm + (size_t)s * TB, a.base + (uint64_t)t * TB, TB, &full[s]);
      }
      ++p_issued;
    }
  };
  if (threadIdx.x == 0) produce(S);  // fill the ring
  const uint32_t my = smem_u32(dsm) + threadIdx.x * 16;
  const uint32_t per_thread_vecs = TB / (THREADS * 16);
  for (uint32_t c = 0;; ++c) {
    const uint32_t s = c % S;
    mbar_wait(&full[s], (c / S) & 1);
    const uint32_t t = tile_of[s];
    if (t == 0xFFFFFFFFu) break;
    const uint32_t sb = my + s * TB;
    for (uint32_t k = 0; k < per_thread_vecs; k += 4) {
      uint32_t w[4][4];
#pragma unroll
      for (int q = 0; q < 4; ++q) lds4(sb + (k + q) * (THREADS * 16), w[q]);
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float x = __uint_as_float(w[q][e]);
          mn = fminf(mn, x);
          mx = fmaxf(mx, x);
        }
    }
    mbar_arrive(&empty[s]);
    if (threadIdx.x == 0) produce(c + 1 + S);  // refill the stage just released (waits for all 256 arrivals)
  }
  if (mn == 12345.678f && mx == 0.1f) a.sink[0] = mn + mx;
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const size_t mib = argc > 2 ? (size_t)atoll(argv[2]) : 4096;
  const int ctas_per_sm = argc > 3 ? atoi(argv[3]) : 4;
  const int stages = argc > 4 ? atoi(argv[4]) : 3;
  const int tile_kib = argc > 5 ? atoi(argv[5]) : 16;
  const int iters = argc > 6 ? atoi(argv[6]) : 20;
  const size_t bytes = mib << 20;
  // rotate over copies when the array would fit in L2
  const int copies = bytes >= ((size_t)1 << 30) ? 1 : (int)std::max<size_t>(2, ((size_t)512 << 20) / bytes);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  std::vector<uint8_t*> bufs(copies);
  for (auto& b : bufs) {
    CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(b, 0x3c, bytes));
  }
  uint32_t* counter;
  float* sink;
  CK(cudaMalloc(&counter, 4));
  CK(cudaMalloc(&sink, 4));
  Args a;
  a.tile_bytes = mode == 0 ? THREADS * 2 * 32 : tile_kib * 1024;
  a.n_tiles = (uint32_t)(bytes / a.tile_bytes);
  a.stages = stages;
  a.counter = counter;
  a.sink = sink;
  const int grid = prop.multiProcessorCount * ctas_per_sm;
  const size_t dyn = mode == 0 ? 0 : (size_t)stages * a.tile_bytes;
  if (mode == 1) CK(cudaFuncSetAttribute(ring_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  if (mode == 2) CK(cudaFuncSetAttribute(ring_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  auto launch = [&](int i) {
    a.base = bufs[i % copies];
    if (mode == 2) CK(cudaMemsetAsync(counter, 0, 4));
    if (mode == 0) reg_kernel<<<grid, THREADS>>>(a);
    if (mode == 1) ring_kernel<false><<<grid, THREADS, dyn>>>(a);
    if (mode == 2) ring_kernel<true><<<grid, THREADS, dyn>>>(a);
  };
  for (int i = 0; i < 3; ++i) launch(i);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) launch(i);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  CK(cudaGetLastError());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double us = ms * 1e3 / iters;
  printf("{\"mode\": %d, \"MiB\": %zu, \"ctas_per_sm\": %d, \"stages\": %d, \"tile_KiB\": %d, \"grid\": %d, "
         "\"smem_per_cta\": %zu, \"us_per_launch\": %.2f, \"GBps\": %.1f}\n",
         mode, mib, ctas_per_sm, stages, mode == 0 ? 16 : tile_kib, grid, dyn, us, (double)bytes / us / 1e3);
  return 0;
}
