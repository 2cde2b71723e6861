// sched_ab.cu -- experiment: does the order in which CTAs take tiles matter for a read-only
// streaming min/max on B200?  Same load shape as the library's big-input kernel
// (256-bit ld.global.nc.L1::no_allocate.L2::256B, U = 2 vectors per thread per tile, ping-pong
// register buffers, 256 threads), three schedules:
//   0  static grid-stride over tiles (what argminmax_scan_kernel does)
//   1  dynamic: CTAs claim batches of B tiles from a global counter (claim issued a batch ahead)
//   2  static contiguous: CTA c streams tiles [c*per, (c+1)*per)
//   3  dynamic per WARP, no CTA barrier: a warp-tile is 32 lanes x U vectors (2 KiB); every warp
//      claims batches of B warp-tiles with its lane 0 (claim issued one batch ahead, broadcast by
//      shuffle) -- the barrier-free variant left open by profiles/r01d_schedule_experiment.md
//   4  static per warp: warp w streams warp-tiles w, w + W, ... (control for the access pattern of 3)
// Prints kernel time (CUDA events) and the spread of per-CTA finish times (globaltimer).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/bin/sched_ab tools/exp/sched_ab.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));   \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

constexpr int U = 2, NW = 8, VEC = 32;
constexpr uint32_t NONE = 0xFFFFFFFFu;

__device__ __forceinline__ void ld8(const uint8_t* p, uint32_t (&w)[NW]) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]),
                 "=r"(w[6]), "=r"(w[7])
               : "l"(p));
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void fold(const uint32_t (&v)[U][NW], float& mn, float& mx) {
#pragma unroll
  for (int j = 0; j < U; ++j)
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const float x = __uint_as_float(v[j][w]);
      mn = fminf(mn, x);
      mx = fmaxf(mx, x);
    }
}

struct Args {
  const uint8_t* base;
  uint32_t n_tiles;
  uint32_t batch;      // tiles per claim (mode 1)
  uint32_t* counter;   // modes 1 and 3, zeroed before the launch
  uint32_t wrap;       // 0xFFFFFFFF: bound of atom.inc (== add 1); a runtime value so that ptxas
                       // cannot turn it into an add and warp-aggregate it (ATOMG + SHFL of the
                       // result at issue time would stall the warp's next loads)
  unsigned long long* t0;
  unsigned long long* t1;
  float* sink;
};

template <int MODE>
__global__ void __launch_bounds__(256) stream_kernel(const Args a) {
  __shared__ uint32_t s_next[2];
  const unsigned long long tb = gtime();
  const uint32_t tile_vecs = blockDim.x * U;
  const uint64_t tile_bytes = (uint64_t)tile_vecs * VEC;
  const uint8_t* lane_base = a.base + (uint64_t)threadIdx.x * VEC;
  const uint64_t jstride = (uint64_t)blockDim.x * VEC;
  float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
  const uint32_t nt = a.n_tiles;

  // tile sequence generator
  uint32_t b_cur = 0, b_next = NONE, i_in = 0, pending = NONE, par = 0, t_end = 0, t = NONE;
  const uint32_t B = a.batch;
  const uint32_t nb = (nt + B - 1) / B;
  if (MODE == 0) {
    t = blockIdx.x < nt ? blockIdx.x : NONE;
  } else if (MODE == 2) {
    const uint32_t per = (nt + gridDim.x - 1) / gridDim.x;
    t = blockIdx.x * per;
    t_end = min(nt, t + per);
    if (t >= t_end) t = NONE;
  } else {
    b_cur = blockIdx.x;
    if (threadIdx.x == 0) s_next[0] = gridDim.x + atomicAdd(a.counter, 1u);
    __syncthreads();
    b_next = s_next[0];
    par = 1;
    if (threadIdx.x == 0 && b_next < nb) pending = gridDim.x + atomicAdd(a.counter, 1u);
    t = b_cur < nb ? b_cur * B : NONE;
  }
  auto next_tile = [&](uint32_t cur) -> uint32_t {
    if (MODE == 0) {
      const uint32_t n = cur + gridDim.x;
      return n < nt ? n : NONE;
    } else if (MODE == 2) {
      const uint32_t n = cur + 1;
      return n < t_end ? n : NONE;
    } else {
      ++i_in;
      const uint32_t n = cur + 1;
      if (i_in < B && n < nt) return n;
      // batch boundary: move on to the batch claimed earlier, hand the next claim round
      if (b_next >= nb) return NONE;
      b_cur = b_next;
      i_in = 0;
      if (threadIdx.x == 0) s_next[par] = pending;  // claimed one batch ago: long since returned
      __syncthreads();
      b_next = s_next[par];
      par ^= 1;
      if (threadIdx.x == 0 && b_next < nb) pending = gridDim.x + atomicAdd(a.counter, 1u);
      return b_cur * B;
    }
  };

  uint32_t va[U][NW], vb[U][NW];
  if (t != NONE) {
    const uint8_t* q = lane_base + (uint64_t)t * tile_bytes;
#pragma unroll
    for (int j = 0; j < U; ++j) ld8(q + j * jstride, va[j]);
  }
  while (t != NONE) {
    const uint32_t t1 = next_tile(t);
    if (t1 != NONE) {
      const uint8_t* q = lane_base + (uint64_t)t1 * tile_bytes;
#pragma unroll
      for (int j = 0; j < U; ++j) ld8(q + j * jstride, vb[j]);
    }
    fold(va, mn, mx);
    if (t1 == NONE) break;
    const uint32_t t2 = next_tile(t1);
    if (t2 != NONE) {
      const uint8_t* q = lane_base + (uint64_t)t2 * tile_bytes;
#pragma unroll
      for (int j = 0; j < U; ++j) ld8(q + j * jstride, va[j]);
    }
    fold(vb, mn, mx);
    t = t2;
  }
  const unsigned long long te = gtime();
  if (threadIdx.x == 0) {
    a.t0[blockIdx.x] = tb;
    a.t1[blockIdx.x] = te;
  }
  if (mn == 123.456f && mx == 654.321f) a.sink[0] = mn;  // keep the folds alive
}

__device__ __forceinline__ uint32_t claim_raw(const Args& a) {
  uint32_t r;
  asm volatile("atom.global.inc.u32 %0, [%1], %2;" : "=r"(r) : "l"(a.counter), "r"(a.wrap));
  return r;
}

// modes 3 / 4: the unit of scheduling is a WARP (tile = 32 lanes x U x 32 B = 2 KiB)
template <int MODE>
__global__ void __launch_bounds__(256) stream_warp_kernel(const Args a) {
  const unsigned long long tb = gtime();
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
  constexpr uint64_t WT_BYTES = 32ull * U * VEC;  // bytes per warp-tile
  const uint32_t nwt = a.n_tiles * (blockDim.x >> 5);  // warp-tiles in the array (CTA tile = 8 of them)
  const uint8_t* lane_base = a.base + (uint64_t)lane * VEC;
  const uint64_t jstride = 32ull * VEC;
  float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
  const uint32_t B = a.batch;
  const uint32_t nb = (nwt + B - 1) / B;
  uint32_t b_cur = warp_global, b_next = warp_global + n_warps, i_in = 0, pending = 0;
  if (MODE == 3 && lane == 0 && b_next < nb) pending = claim_raw(a);  // consumed one batch later
  uint32_t t = b_cur < nb ? b_cur * B : NONE;
  auto next_tile = [&](uint32_t cur) -> uint32_t {
    ++i_in;
    const uint32_t n = cur + 1;
    if (i_in < B && n < nwt) return n;
    if (b_next >= nb) return NONE;
    b_cur = b_next;
    i_in = 0;
    if (MODE == 3) {
      const uint32_t claimed = 2u * n_warps + __shfl_sync(0xFFFFFFFFu, pending, 0);
      b_next = claimed;
      if (lane == 0 && b_next < nb) pending = claim_raw(a);
    } else {
      b_next = b_cur + n_warps;
    }
    return b_cur * B;
  };
  uint32_t va[U][NW], vb[U][NW];
  if (t != NONE) {
    const uint8_t* q = lane_base + (uint64_t)t * WT_BYTES;
#pragma unroll
    for (int j = 0; j < U; ++j) ld8(q + j * jstride, va[j]);
  }
  while (t != NONE) {
    const uint32_t t1 = next_tile(t);
    if (t1 != NONE) {
      const uint8_t* q = lane_base + (uint64_t)t1 * WT_BYTES;
#pragma unroll
      for (int j = 0; j < U; ++j) ld8(q + j * jstride, vb[j]);
    }
    fold(va, mn, mx);
    if (t1 == NONE) break;
    const uint32_t t2 = next_tile(t1);
    if (t2 != NONE) {
      const uint8_t* q = lane_base + (uint64_t)t2 * WT_BYTES;
#pragma unroll
      for (int j = 0; j < U; ++j) ld8(q + j * jstride, va[j]);
    }
    fold(vb, mn, mx);
    t = t2;
  }
  const unsigned long long te = gtime();
  if (threadIdx.x == 0) {
    a.t0[blockIdx.x] = tb;
    a.t1[blockIdx.x] = te;
  }
  if (mn == 123.456f && mx == 654.321f) a.sink[0] = mn;
}

template <int MODE>
static void run(const char* name, const Args& a0, int grid, int reps, uint64_t bytes) {
  Args a = a0;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::vector<float> ms;
  for (int r = 0; r < reps + 3; ++r) {
    CK(cudaMemsetAsync(a.counter, 0, 4));
    CK(cudaEventRecord(e0));
    if (MODE >= 3) {
      stream_warp_kernel<MODE><<<grid, 256>>>(a);
    } else {
      stream_kernel<(MODE >= 3 ? 0 : MODE)><<<grid, 256>>>(a);
    }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float m;
    CK(cudaEventElapsedTime(&m, e0, e1));
    if (r >= 3) ms.push_back(m);
  }
  CK(cudaGetLastError());
  std::sort(ms.begin(), ms.end());
  std::vector<unsigned long long> t0(grid), t1(grid);
  CK(cudaMemcpy(t0.data(), a.t0, 8 * grid, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(t1.data(), a.t1, 8 * grid, cudaMemcpyDeviceToHost));
  const unsigned long long s = *std::min_element(t0.begin(), t0.end());
  std::vector<double> end(grid);
  double start_max = 0, sum = 0;
  for (int i = 0; i < grid; ++i) {
    end[i] = (double)(t1[i] - s) / 1e3;
    start_max = std::max(start_max, (double)(t0[i] - s) / 1e3);
    sum += end[i];
  }
  std::sort(end.begin(), end.end());
  const double med = ms[ms.size() / 2];
  printf("{\"sched\": \"%s\", \"grid\": %d, \"batch\": %u, \"ms_median\": %.4f, \"ms_min\": %.4f, "
         "\"GBps_median\": %.1f, \"cta_end_us\": {\"min\": %.1f, \"p10\": %.1f, \"median\": %.1f, "
         "\"p90\": %.1f, \"max\": %.1f, \"mean\": %.1f}, \"cta_start_max_us\": %.1f}\n",
         name, grid, a.batch, med, ms[0], (double)bytes / med / 1e6, end[0], end[grid / 10],
         end[grid / 2], end[grid * 9 / 10], end[grid - 1], sum / grid, start_max);
  fflush(stdout);
}

int main(int argc, char** argv) {
  const uint64_t bytes = argc > 1 ? strtoull(argv[1], nullptr, 10) : (1ull << 32);
  const int reps = argc > 2 ? atoi(argv[2]) : 20;
  uint8_t* buf;
  CK(cudaMalloc(&buf, bytes));
  CK(cudaMemset(buf, 0x3c, bytes));
  Args a;
  a.base = buf;
  a.wrap = 0xFFFFFFFFu;
  a.n_tiles = (uint32_t)(bytes / (256ull * U * VEC));
  CK(cudaMalloc(&a.counter, 4));
  CK(cudaMalloc(&a.t0, 8 * 4096));
  CK(cudaMalloc(&a.t1, 8 * 4096));
  CK(cudaMalloc(&a.sink, 4));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  for (int per_sm : {4, 5, 2}) {
    const int grid = sms * per_sm;
    a.batch = 1;
    run<0>("static_stride", a, grid, reps, bytes);
    run<2>("static_contiguous", a, grid, reps, bytes);
    for (uint32_t b : {1u, 2u, 4u, 8u, 32u}) {
      a.batch = b;
      run<1>("dynamic", a, grid, reps, bytes);
    }
    for (uint32_t b : {8u, 32u, 128u}) {  // warp-tiles (2 KiB) per claim: 16 / 64 / 256 KiB
      a.batch = b;
      run<4>("static_per_warp", a, grid, reps, bytes);
      run<3>("dynamic_per_warp", a, grid, reps, bytes);
    }
  }
  return 0;
}
