// scan_inst.cu -- instantiates the scan kernels for ONE group of policies; compiled once per
// group with -DAMM_GROUP=<k> so the groups build in parallel (see argminmax_b200/_build.py).
#include "scan_launch.cuh"

namespace amm {

#ifndef AMM_GROUP
#error "compile with -DAMM_GROUP=<0..6>"
#endif

#if AMM_GROUP == 0
int launch_group0(int slot, amm_workspace* ws, ScanParams& p, cudaStream_t s) {
  return slot == 0 ? launch_variant<PolInt8<true>>(ws, p, 0, s) : launch_variant<PolInt8<false>>(ws, p, 4, s);
}
int windows_group0(int slot, amm_workspace* ws, const WindowParams& p, uint64_t mean, cudaStream_t s) {
  return slot == 0 ? launch_windows<PolInt8<true>>(ws, p, mean, s) : launch_windows<PolInt8<false>>(ws, p, mean, s);
}
#elif AMM_GROUP == 1
int launch_group1(int slot, amm_workspace* ws, ScanParams& p, cudaStream_t s) {
  return slot == 1 ? launch_variant<PolInt16<true>>(ws, p, 1, s) : launch_variant<PolInt16<false>>(ws, p, 5, s);
}
int windows_group1(int slot, amm_workspace* ws, const WindowParams& p, uint64_t mean, cudaStream_t s) {
  return slot == 1 ? launch_windows<PolInt16<true>>(ws, p, mean, s) : launch_windows<PolInt16<false>>(ws, p, mean, s);
}
#elif AMM_GROUP == 2
int launch_group2(int slot, amm_workspace* ws, ScanParams& p, cudaStream_t s) {
  return slot == 2 ? launch_variant<PolInt32<true>>(ws, p, 2, s) : launch_variant<PolInt32<false>>(ws, p, 6, s);
}
int windows_group2(int slot, amm_workspace* ws, const WindowParams& p, uint64_t mean, cudaStream_t s) {
  return slot == 2 ? launch_windows<PolInt32<true>>(ws, p, mean, s) : launch_windows<PolInt32<false>>(ws, p, mean, s);
}
#elif AMM_GROUP == 3
int launch_group3(int slot, amm_workspace* ws, ScanParams& p, cudaStream_t s) {
  return slot == 3 ? launch_variant<PolInt64<true>>(ws, p, 3, s) : launch_variant<PolInt64<false>>(ws, p, 7, s);
}
int windows_group3(int slot, amm_workspace* ws, const WindowParams& p, uint64_t mean, cudaStream_t s) {
  return slot == 3 ? launch_windows<PolInt64<true>>(ws, p, mean, s) : launch_windows<PolInt64<false>>(ws, p, mean, s);
}
#elif AMM_GROUP == 4
int launch_group4(int slot, amm_workspace* ws, ScanParams& p, cudaStream_t s) {
  return slot == 8 ? launch_variant<PolF16<MODE_IGNORE>>(ws, p, 8, s) : launch_variant<PolF16<MODE_RETURN>>(ws, p, 9, s);
}
int windows_group4(int slot, amm_workspace* ws, const WindowParams& p, uint64_t mean, cudaStream_t s) {
  return slot == 8 ? launch_windows<PolF16<MODE_IGNORE>>(ws, p, mean, s) : launch_windows<PolF16<MODE_RETURN>>(ws, p, mean, s);
}
#elif AMM_GROUP == 5
int launch_group5(int slot, amm_workspace* ws, ScanParams& p, cudaStream_t s) {
  return slot == 10 ? launch_variant<PolF32<MODE_IGNORE>>(ws, p, 10, s) : launch_variant<PolF32<MODE_RETURN>>(ws, p, 11, s);
}
int windows_group5(int slot, amm_workspace* ws, const WindowParams& p, uint64_t mean, cudaStream_t s) {
  return slot == 10 ? launch_windows<PolF32<MODE_IGNORE>>(ws, p, mean, s) : launch_windows<PolF32<MODE_RETURN>>(ws, p, mean, s);
}
#elif AMM_GROUP == 6
int launch_group6(int slot, amm_workspace* ws, ScanParams& p, cudaStream_t s) {
  return slot == 12 ? launch_variant<PolF64<MODE_IGNORE>>(ws, p, 12, s) : launch_variant<PolF64<MODE_RETURN>>(ws, p, 13, s);
}
int windows_group6(int slot, amm_workspace* ws, const WindowParams& p, uint64_t mean, cudaStream_t s) {
  return slot == 12 ? launch_windows<PolF64<MODE_IGNORE>>(ws, p, mean, s) : launch_windows<PolF64<MODE_RETURN>>(ws, p, mean, s);
}
#endif

}  // namespace amm
