// argminmax_b200.hpp -- header-only C++17 mirror of the reference's trait surface over the
// C ABI (argminmax_b200.h).  It is what the Rust shim (rust/src/cuda/mod.rs) is, written in
// the language this image can compile: same method names, argument meaning and error
// behaviour as `trait ArgMinMax` / `trait NaNArgMinMax` (reference src/lib.rs:134-244):
//   argminmax() / argmin() / argmax()           ints + floats (floats ignore NaN)
//   nanargminmax() / nanargmin() / nanargmax()  floats only (first NaN wins)
// Errors (the reference panics, e.g. on an empty slice) surface as amm::Error.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "argminmax_b200.h"

namespace amm {

struct f16 {  // IEEE binary16 carried as raw bits (half::f16 on the Rust side)
  uint16_t bits;
};

template <typename T> struct dtype_of;
template <> struct dtype_of<int8_t> { static constexpr int code = AMM_I8; static constexpr bool is_float = false; };
template <> struct dtype_of<int16_t> { static constexpr int code = AMM_I16; static constexpr bool is_float = false; };
template <> struct dtype_of<int32_t> { static constexpr int code = AMM_I32; static constexpr bool is_float = false; };
template <> struct dtype_of<int64_t> { static constexpr int code = AMM_I64; static constexpr bool is_float = false; };
template <> struct dtype_of<uint8_t> { static constexpr int code = AMM_U8; static constexpr bool is_float = false; };
template <> struct dtype_of<uint16_t> { static constexpr int code = AMM_U16; static constexpr bool is_float = false; };
template <> struct dtype_of<uint32_t> { static constexpr int code = AMM_U32; static constexpr bool is_float = false; };
template <> struct dtype_of<uint64_t> { static constexpr int code = AMM_U64; static constexpr bool is_float = false; };
template <> struct dtype_of<f16> { static constexpr int code = AMM_F16; static constexpr bool is_float = true; };
template <> struct dtype_of<float> { static constexpr int code = AMM_F32; static constexpr bool is_float = true; };
template <> struct dtype_of<double> { static constexpr int code = AMM_F64; static constexpr bool is_float = true; };

class Error : public std::runtime_error {
 public:
  Error(int status, const char* where)
      : std::runtime_error(std::string(where) + ": " + amm_status_string(status)), status_(status) {}
  int status() const { return status_; }

 private:
  int status_;
};

inline void check(int status, const char* where) {
  if (status != AMM_OK) throw Error(status, where);
}

class Workspace {
 public:
  explicit Workspace(int device = 0) : device_(device) {
    check(amm_workspace_create(device, &ws_), "amm_workspace_create");
  }
  ~Workspace() { amm_workspace_destroy(ws_); }
  Workspace(const Workspace&) = delete;
  Workspace& operator=(const Workspace&) = delete;
  amm_workspace* get() const { return ws_; }
  int device() const { return device_; }

 private:
  amm_workspace* ws_ = nullptr;
  int device_;
};

// `n` elements of T in the HBM of one B200.  Borrowing by default (like `&[T]`);
// from_host() makes an owning device copy.
template <typename T>
class DeviceSlice {
 public:
  DeviceSlice(const T* dev_ptr, size_t n, Workspace& ws, void* stream = nullptr)
      : ptr_(dev_ptr), n_(n), ws_(&ws), stream_(stream) {}
  DeviceSlice(DeviceSlice&& o) noexcept
      : ptr_(o.ptr_), n_(o.n_), ws_(o.ws_), stream_(o.stream_), owned_(o.owned_) {
    o.owned_ = false;
  }
  DeviceSlice(const DeviceSlice&) = delete;
  ~DeviceSlice() {
    if (owned_) amm_device_free(ws_->device(), const_cast<T*>(ptr_));
  }
  static DeviceSlice from_host(const T* host, size_t n, Workspace& ws) {
    void* d = nullptr;
    check(amm_device_alloc(ws.device(), n * sizeof(T), &d), "amm_device_alloc");
    if (n) check(amm_copy_to_device(ws.device(), d, host, n * sizeof(T)), "amm_copy_to_device");
    DeviceSlice s(static_cast<const T*>(d), n, ws);
    s.owned_ = true;
    return s;
  }
  size_t len() const { return n_; }
  const T* as_ptr() const { return ptr_; }

  // trait ArgMinMax
  std::pair<size_t, size_t> argminmax() const { return run(AMM_IGNORE_NAN); }
  size_t argmin() const { return run(AMM_IGNORE_NAN).first; }
  size_t argmax() const { return run(AMM_IGNORE_NAN).second; }
  // trait NaNArgMinMax (float dtypes only, as in the reference)
  template <typename U = T, typename = std::enable_if_t<dtype_of<U>::is_float>>
  std::pair<size_t, size_t> nanargminmax() const { return run(AMM_RETURN_NAN); }
  template <typename U = T, typename = std::enable_if_t<dtype_of<U>::is_float>>
  size_t nanargmin() const { return run(AMM_RETURN_NAN).first; }
  template <typename U = T, typename = std::enable_if_t<dtype_of<U>::is_float>>
  size_t nanargmax() const { return run(AMM_RETURN_NAN).second; }

  // Widened row (not part of the reference's traits): what a down-sampler gets by
  // calling argminmax() once per bin -- (argmin, argmax) of every consecutive window of `window`
  // elements (the last may be shorter), absolute indices, ONE launch.  Blocking; default-stream
  // slices only (the result is copied back with the C ABI's synchronous copy).
  std::vector<std::pair<size_t, size_t>> argminmax_windows(size_t window, int mode = AMM_IGNORE_NAN) const {
    if (stream_ != nullptr) throw Error(AMM_ERR_ARG, "argminmax_windows: default-stream slices only");
    if (window == 0) throw Error(AMM_ERR_ARG, "argminmax_windows: window == 0");
    const size_t nwin = (n_ + window - 1) / window;
    std::vector<std::pair<size_t, size_t>> res(nwin);
    if (nwin == 0) check(AMM_ERR_EMPTY, "amm_argminmax_fixed_windows");
    void* d_out = nullptr;
    check(amm_device_alloc(ws_->device(), nwin * 2 * sizeof(uint64_t), &d_out), "amm_device_alloc");
    std::vector<uint64_t> h(2 * nwin);
    const int rc = amm_argminmax_fixed_windows(dtype_of<T>::code, ptr_, n_, window, mode, ws_->get(),
                                               nullptr, static_cast<uint64_t*>(d_out));
    const int rc2 = rc ? rc : amm_copy_to_host(ws_->device(), h.data(), d_out, h.size() * sizeof(uint64_t));
    amm_device_free(ws_->device(), d_out);
    check(rc2, "amm_argminmax_fixed_windows");
    for (size_t w = 0; w < nwin; ++w) res[w] = {static_cast<size_t>(h[2 * w]), static_cast<size_t>(h[2 * w + 1])};
    return res;
  }

 private:
  std::pair<size_t, size_t> run(int mode) const {
    uint64_t out[2];
    check(amm_argminmax(dtype_of<T>::code, ptr_, n_, mode, ws_->get(), stream_, out), "amm_argminmax");
    return {static_cast<size_t>(out[0]), static_cast<size_t>(out[1])};
  }
  const T* ptr_;
  size_t n_;
  Workspace* ws_;
  void* stream_;
  bool owned_ = false;
};

// `&[T]` / Vec<T> in host memory: chunked H2D overlapped with the scan (PCIe-bound).
template <typename T>
std::pair<size_t, size_t> host_argminmax(const T* host, size_t n, Workspace& ws, int mode = AMM_IGNORE_NAN) {
  uint64_t out[2];
  check(amm_host_argminmax(dtype_of<T>::code, host, n, mode, ws.get(), out), "amm_host_argminmax");
  return {static_cast<size_t>(out[0]), static_cast<size_t>(out[1])};
}

// One logical array as consecutive shards on several GPUs of one box (one host thread).
template <typename T>
class ShardedDeviceSlice {
 public:
  struct Shard {
    int device;
    const T* dev_ptr;
    size_t len;
  };
  explicit ShardedDeviceSlice(std::vector<Shard> shards) : shards_(std::move(shards)) {
    std::vector<int> devs;
    for (auto& s : shards_) devs.push_back(s.device);
    check(amm_sharded_create(static_cast<int>(devs.size()), devs.data(), &sh_), "amm_sharded_create");
  }
  ~ShardedDeviceSlice() { amm_sharded_destroy(sh_); }
  ShardedDeviceSlice(const ShardedDeviceSlice&) = delete;
  bool uses_nccl() const { return amm_sharded_uses_nccl(sh_) != 0; }
  std::pair<size_t, size_t> argminmax() const { return run(AMM_IGNORE_NAN); }
  size_t argmin() const { return run(AMM_IGNORE_NAN).first; }
  size_t argmax() const { return run(AMM_IGNORE_NAN).second; }
  template <typename U = T, typename = std::enable_if_t<dtype_of<U>::is_float>>
  std::pair<size_t, size_t> nanargminmax() const { return run(AMM_RETURN_NAN); }

 private:
  std::pair<size_t, size_t> run(int mode) const {
    std::vector<const void*> ptrs;
    std::vector<uint64_t> lens;
    for (auto& s : shards_) {
      ptrs.push_back(s.dev_ptr);
      lens.push_back(s.len);
    }
    uint64_t out[2];
    check(amm_sharded_argminmax(sh_, dtype_of<T>::code, ptrs.data(), lens.data(), mode, out),
          "amm_sharded_argminmax");
    return {static_cast<size_t>(out[0]), static_cast<size_t>(out[1])};
  }
  std::vector<Shard> shards_;
  amm_sharded* sh_ = nullptr;
};

}  // namespace amm
