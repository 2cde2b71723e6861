// latency_floor.cu -- experiment: what is the floor of "launch one kernel, get 16 bytes back on the
// host" on this box?  Compares against the synchronous amm_argminmax call (8-12 us for n <= 1 Mi).
//   A  empty <<<1,32>>> kernel, result flag written to mapped pinned memory, host polls the flag
//   B  as A, launched through cudaLaunchKernelEx with the programmatic-stream-serialization attr
//   C  as A, host waits with cudaStreamSynchronize instead of polling
//   D  <<<256,256>>>: every CTA reads 16 KiB (4 MiB in total, L2-resident), warp+CTA reduce,
//      last-block-done ticket, flag -- the skeleton of the small-input scan without its bookkeeping
//   E  as D without the ticket: every CTA atomically folds its packed (value, index) into one word,
//      the last ticket holder only copies the word out
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/bin/latency_floor tools/exp/latency_floor.cu
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

template <class F>
static void measure(const char* name, volatile uint64_t* hflag, bool poll, F launch) {
  std::vector<double> us;
  uint64_t seq = 0;
  for (int i = 0; i < 2200; ++i) {
    ++seq;
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
  measure("A' empty kernel + poll (again)", hflag, true, [&](uint64_t s) { k_empty<<<1, 32>>>(dflag, s); });
  return 0;
}
