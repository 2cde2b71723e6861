// C++ host program over include/argminmax_b200.hpp: the reference's documented examples and
// known-answer vectors through the DeviceSlice<T> mirror of the trait surface, plus a
// C-level latency sweep (no Python in the loop).  Built with plain g++ (no nvcc, no CUDA
// headers): everything it needs is in the C ABI.
//   device_slice_demo check      -> prints "OK <n checks>" or throws
//   device_slice_demo latency    -> JSON lines: us/call (median, min) for f32 n = 2^10..2^24
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

#include "argminmax_b200.hpp"

using amm::DeviceSlice;

static int g_checks = 0;
#define EXPECT_EQ_PAIR(got, a, b)                                                       \
  do {                                                                                  \
    auto _g = (got);                                                                    \
    if (_g.first != (size_t)(a) || _g.second != (size_t)(b)) {                          \
      std::fprintf(stderr, "FAIL line %d: got (%zu,%zu) want (%zu,%zu)\n", __LINE__,    \
                   _g.first, _g.second, (size_t)(a), (size_t)(b));                      \
      return 1;                                                                         \
    }                                                                                   \
    ++g_checks;                                                                         \
  } while (0)

template <typename T>
static int monotonic(amm::Workspace& ws, size_t max_index) {
  // tests/argminmax_test.rs:100-140: a[i] = i % max_index, n = 100_000 -> (0, max_index-1)
  std::vector<T> v(100000);
  for (size_t i = 0; i < v.size(); ++i) v[i] = (T)(i % max_index);
  auto d = DeviceSlice<T>::from_host(v.data(), v.size(), ws);
  EXPECT_EQ_PAIR(d.argminmax(), 0, max_index - 1);
  if (d.argmin() != 0 || d.argmax() != max_index - 1) return 1;
  return 0;
}

static int run_checks() {
  amm::Workspace ws(0);
  {  // src/lib.rs:50-53
    std::vector<int32_t> v = {0, 1, 2, 3, 4, 5};
    auto d = DeviceSlice<int32_t>::from_host(v.data(), v.size(), ws);
    EXPECT_EQ_PAIR(d.argminmax(), 0, 5);
  }
  {  // src/lib.rs:61-67
    const float nan = std::numeric_limits<float>::quiet_NaN();
    std::vector<float> v = {nan, 1, nan, 3, 4, 5};
    auto d = DeviceSlice<float>::from_host(v.data(), v.size(), ws);
    EXPECT_EQ_PAIR(d.argminmax(), 1, 5);
    EXPECT_EQ_PAIR(d.nanargminmax(), 0, 0);
  }
  {  // README.md:53-55
    std::vector<int32_t> v(200000);
    std::iota(v.begin(), v.end(), 0);
    auto d = DeviceSlice<int32_t>::from_host(v.data(), v.size(), ws);
    EXPECT_EQ_PAIR(d.argminmax(), 0, 199999);
  }
  if (monotonic<int8_t>(ws, 127) || monotonic<uint8_t>(ws, 255) || monotonic<int16_t>(ws, 32767) ||
      monotonic<uint16_t>(ws, 65535) || monotonic<int32_t>(ws, 100000) ||
      monotonic<uint32_t>(ws, 100000) || monotonic<int64_t>(ws, 100000) ||
      monotonic<uint64_t>(ws, 100000) || monotonic<float>(ws, 100000) ||
      monotonic<double>(ws, 100000))
    return 1;
  {  // src/simd/test_utils.rs:125-150: 64 x one, MIN @5,13,41, MAX @7,17,31 -> (5,7)
    std::vector<double> v(64, 1.0);
    v[5] = v[13] = v[41] = std::numeric_limits<double>::lowest();
    v[7] = v[17] = v[31] = std::numeric_limits<double>::max();
    auto d = DeviceSlice<double>::from_host(v.data(), v.size(), ws);
    EXPECT_EQ_PAIR(d.argminmax(), 5, 7);
    EXPECT_EQ_PAIR(d.nanargminmax(), 5, 7);
  }
  {  // f16 as raw bits: 1027 x NaN -> (0,0) both modes (test_utils.rs:508-527, 668-687)
    std::vector<amm::f16> v(1027, amm::f16{0x7E00});
    auto d = DeviceSlice<amm::f16>::from_host(v.data(), v.size(), ws);
    EXPECT_EQ_PAIR(d.argminmax(), 0, 0);
    EXPECT_EQ_PAIR(d.nanargminmax(), 0, 0);
  }
  {  // host slice path + empty input error (reference: panic)
    std::vector<float> v(3000001);
    for (size_t i = 0; i < v.size(); ++i) v[i] = std::sin((float)i);
    v[1234567] = -5.f;
    v[2345678] = 7.f;
    EXPECT_EQ_PAIR(amm::host_argminmax(v.data(), v.size(), ws), 1234567, 2345678);
    bool threw = false;
    try {
      DeviceSlice<float> e(nullptr, 0, ws);
      (void)e.argminmax();
    } catch (const amm::Error& err) {
      threw = true;
    }
    if (!threw) return 1;
    ++g_checks;
  }
  {  // widened row: one launch for many bins == argminmax() per bin (MinMax down-sampling)
    std::vector<float> v(100003);
    for (size_t i = 0; i < v.size(); ++i) v[i] = std::sin(0.37f * (float)i) + 1e-3f * (float)(i % 97);
    auto d = DeviceSlice<float>::from_host(v.data(), v.size(), ws);
    for (size_t window : {64u, 1000u, 30000u}) {
      const auto got = d.argminmax_windows(window);
      for (size_t w = 0; w < got.size(); ++w) {
        const size_t b = w * window, e = std::min(v.size(), b + window);
        size_t imin = b, imax = b;
        for (size_t i = b; i < e; ++i) {  // first occurrence, strict compares
          if (v[i] < v[imin]) imin = i;
          if (v[i] > v[imax]) imax = i;
        }
        if (got[w].first != imin || got[w].second != imax) {
          std::fprintf(stderr, "FAIL windows(%zu) bin %zu\n", window, w);
          return 1;
        }
      }
      ++g_checks;
    }
  }
  std::printf("OK %d\n", g_checks);
  return 0;
}

static int run_latency() {
  amm::Workspace ws(0);
  const size_t nmax = (size_t)1 << 24;
  std::vector<float> v(nmax);
  uint32_t s = 12345;
  for (auto& x : v) {
    s = s * 1664525u + 1013904223u;
    x = (float)(int32_t)s;
  }
  auto d = DeviceSlice<float>::from_host(v.data(), nmax, ws);
  for (int lg = 10; lg <= 24; ++lg) {
    const size_t n = (size_t)1 << lg;
    DeviceSlice<float> part(d.as_ptr(), n, ws);
    for (int i = 0; i < 100; ++i) (void)part.argminmax();
    std::vector<double> us(2000);
    for (auto& t : us) {
      auto t0 = std::chrono::steady_clock::now();
      (void)part.argminmax();
      t = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    }
    std::sort(us.begin(), us.end());
    std::printf("{\"n\": %zu, \"us_call_median\": %.2f, \"us_call_min\": %.2f, \"us_call_p99\": %.2f}\n", n,
                us[us.size() / 2], us[0], us[us.size() * 99 / 100]);
  }
  return 0;
}

int main(int argc, char** argv) {
  try {
    if (argc > 1 && std::strcmp(argv[1], "latency") == 0) return run_latency();
    return run_checks();
  } catch (const std::exception& e) {
    std::fprintf(stderr, "exception: %s\n", e.what());
    return 2;
  }
}
