//! `mod cuda` -- the B200 (sm_100a) backend, a sibling of `src/simd/` reached from the same
//! trait surface (`ArgMinMax`, `NaNArgMinMax`, reference src/lib.rs:134-244) on two new receiver
//! types: `DeviceSlice<T>` (one array in the HBM of one GPU) and `ShardedDeviceSlice<T>` (one
//! array split over the GPUs of one box).  Gated by the `cuda` Cargo feature; `build.rs` runs
//! `nvcc -gencode arch=compute_100a,code=sm_100a` and links `libargminmax_b200`.
//!
//! NOT COMPILED IN THE BUILD CONTAINER (the image has no rustc/cargo).  This file is the
//! reference-side binding a maintainer adds; it mirrors `include/argminmax_b200.h` 1:1 and holds
//! no logic beyond turning a non-zero status into a panic (the reference's own error convention:
//! `assert!(!arr.is_empty())`, src/scalar/generic.rs:182).  The C++ mirror of exactly this
//! surface (`include/argminmax_b200.hpp`) is what the test-suite compiles and runs.
use std::ffi::c_void;
use std::marker::PhantomData;
use std::os::raw::c_int;

use crate::{ArgMinMax, NaNArgMinMax};

#[repr(C)]
pub struct AmmWorkspace {
    _private: [u8; 0],
}
#[repr(C)]
pub struct AmmSharded {
    _private: [u8; 0],
}

pub const AMM_IGNORE_NAN: c_int = 0; // Int / FloatIgnoreNaN (src/dtype_strategy.rs:6,15)
pub const AMM_RETURN_NAN: c_int = 1; // FloatReturnNaN       (src/dtype_strategy.rs:24)

extern "C" {
    fn amm_workspace_create(device: c_int, ws: *mut *mut AmmWorkspace) -> c_int;
    fn amm_workspace_destroy(ws: *mut AmmWorkspace) -> c_int;
    fn amm_argminmax(
        dtype: c_int,
        dev_ptr: *const c_void,
        n: u64,
        mode: c_int,
        ws: *mut AmmWorkspace,
        stream: *mut c_void,
        out: *mut u64,
    ) -> c_int;
    fn amm_host_argminmax(
        dtype: c_int,
        host_ptr: *const c_void,
        n: u64,
        mode: c_int,
        ws: *mut AmmWorkspace,
        out: *mut u64,
    ) -> c_int;
    fn amm_sharded_create(ngpu: c_int, devices: *const c_int, sh: *mut *mut AmmSharded) -> c_int;
    fn amm_sharded_destroy(sh: *mut AmmSharded) -> c_int;
    fn amm_sharded_argminmax(
        sh: *mut AmmSharded,
        dtype: c_int,
        dev_ptrs: *const *const c_void,
        shard_lens: *const u64,
        mode: c_int,
        out: *mut u64,
    ) -> c_int;
    fn amm_argminmax_fixed_windows(
        dtype: c_int,
        dev_ptr: *const c_void,
        n: u64,
        window: u64,
        mode: c_int,
        ws: *mut AmmWorkspace,
        stream: *mut c_void,
        dev_out: *mut u64,
    ) -> c_int;
    fn amm_copy_to_host(device: c_int, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    fn amm_device_alloc(device: c_int, bytes: usize, dev_ptr: *mut *mut c_void) -> c_int;
    fn amm_device_free(device: c_int, dev_ptr: *mut c_void) -> c_int;
    fn amm_copy_to_device(device: c_int, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    fn amm_status_string(status: c_int) -> *const std::os::raw::c_char;
}

fn check(status: c_int, what: &str) {
    if status != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(amm_status_string(status)) };
        panic!("{what}: {}", msg.to_string_lossy()); // the reference panics as well
    }
}

/// Element types the CUDA backend understands: the eleven dtypes of README.md:22.
/// `CODE` is `amm_dtype` in include/argminmax_b200.h.
pub trait CudaDType: Copy {
    const CODE: c_int;
}
macro_rules! impl_cuda_dtype {
    ($($t:ty => $c:expr),*) => { $(impl CudaDType for $t { const CODE: c_int = $c; })* };
}
impl_cuda_dtype!(i8 => 0, i16 => 1, i32 => 2, i64 => 3, u8 => 4, u16 => 5, u32 => 6, u64 => 7,
                 f32 => 9, f64 => 10);
#[cfg(feature = "half")]
impl CudaDType for half::f16 {
    const CODE: c_int = 8;
}

/// Per-device scratch (per-CTA result slots, ticket and packed words, pinned result slot).  One per concurrent caller.
pub struct Workspace {
    raw: *mut AmmWorkspace,
    device: i32,
}
impl Workspace {
    pub fn new(device: i32) -> Self {
        let mut raw = std::ptr::null_mut();
        check(unsafe { amm_workspace_create(device, &mut raw) }, "amm_workspace_create");
        Workspace { raw, device }
    }
}
impl Drop for Workspace {
    fn drop(&mut self) {
        unsafe { amm_workspace_destroy(self.raw) };
    }
}

/// `n` elements of `T` resident in the HBM of one B200.
pub struct DeviceSlice<'w, T: CudaDType> {
    ptr: *const T,
    len: usize,
    ws: &'w Workspace,
    owned: bool,
    _t: PhantomData<T>,
}

impl<'w, T: CudaDType> DeviceSlice<'w, T> {
    /// Borrow memory that already lives on the device (e.g. produced by another kernel).
    /// # Safety: `ptr` must be a device pointer valid for `len` elements on `ws`'s device.
    pub unsafe fn from_raw_parts(ptr: *const T, len: usize, ws: &'w Workspace) -> Self {
        DeviceSlice { ptr, len, ws, owned: false, _t: PhantomData }
    }
    /// Upload a host slice (blocking copy).
    pub fn from_slice(data: &[T], ws: &'w Workspace) -> Self {
        let bytes = std::mem::size_of_val(data);
        let mut d: *mut c_void = std::ptr::null_mut();
        check(unsafe { amm_device_alloc(ws.device, bytes, &mut d) }, "amm_device_alloc");
        check(
            unsafe { amm_copy_to_device(ws.device, d, data.as_ptr() as *const c_void, bytes) },
            "amm_copy_to_device",
        );
        DeviceSlice { ptr: d as *const T, len: data.len(), ws, owned: true, _t: PhantomData }
    }
    pub fn len(&self) -> usize {
        self.len
    }
    pub fn is_empty(&self) -> bool {
        self.len == 0
    }
    /// Not part of the reference's traits: what a down-sampler gets by calling
    /// `argminmax()` once per bin -- `(argmin, argmax)` of every consecutive window of `window`
    /// elements (the last may be shorter), absolute indices, in ONE launch.  `return_nan`
    /// selects the `NaNArgMinMax` semantics per window.
    pub fn argminmax_windows(&self, window: usize, return_nan: bool) -> Vec<(usize, usize)> {
        assert!(window > 0 && self.len > 0); // the reference panics on empty input as well
        let nwin = (self.len + window - 1) / window;
        let bytes = nwin * 2 * std::mem::size_of::<u64>();
        let mut d: *mut c_void = std::ptr::null_mut();
        check(unsafe { amm_device_alloc(self.ws.device, bytes, &mut d) }, "amm_device_alloc");
        let mode = if return_nan { AMM_RETURN_NAN } else { AMM_IGNORE_NAN };
        let rc = unsafe {
            amm_argminmax_fixed_windows(T::CODE, self.ptr as *const c_void, self.len as u64,
                                        window as u64, mode, self.ws.raw, std::ptr::null_mut(),
                                        d as *mut u64)
        };
        let mut host = vec![0u64; 2 * nwin];
        let rc2 = if rc != 0 { rc } else {
            unsafe { amm_copy_to_host(self.ws.device, host.as_mut_ptr() as *mut c_void, d, bytes) }
        };
        unsafe { amm_device_free(self.ws.device, d) };
        check(rc2, "amm_argminmax_fixed_windows");
        host.chunks_exact(2).map(|p| (p[0] as usize, p[1] as usize)).collect()
    }
    fn run(&self, mode: c_int) -> (usize, usize) {
        let mut out = [0u64; 2];
        check(
            unsafe {
                amm_argminmax(T::CODE, self.ptr as *const c_void, self.len as u64, mode, self.ws.raw,
                              std::ptr::null_mut(), out.as_mut_ptr())
            },
            "amm_argminmax",
        );
        (out[0] as usize, out[1] as usize)
    }
}
impl<T: CudaDType> Drop for DeviceSlice<'_, T> {
    fn drop(&mut self) {
        if self.owned {
            unsafe { amm_device_free(self.ws.device, self.ptr as *mut c_void) };
        }
    }
}

// ---- the traits, verbatim from src/lib.rs:134-244, on the new receiver type ----------------
impl<T: CudaDType> ArgMinMax for DeviceSlice<'_, T> {
    fn argminmax(&self) -> (usize, usize) {
        self.run(AMM_IGNORE_NAN)
    }
    fn argmin(&self) -> usize {
        self.run(AMM_IGNORE_NAN).0
    }
    fn argmax(&self) -> usize {
        self.run(AMM_IGNORE_NAN).1
    }
}

macro_rules! impl_nan_device_slice {
    ($($t:ty),*) => { $(
        impl NaNArgMinMax for DeviceSlice<'_, $t> {
            fn nanargminmax(&self) -> (usize, usize) { self.run(AMM_RETURN_NAN) }
            fn nanargmin(&self) -> usize { self.run(AMM_RETURN_NAN).0 }
            fn nanargmax(&self) -> usize { self.run(AMM_RETURN_NAN).1 }
        }
    )* };
}
impl_nan_device_slice!(f32, f64);
#[cfg(feature = "half")]
impl_nan_device_slice!(half::f16);

/// A host slice routed through the GPU (chunked H2D overlapped with the scan; PCIe-bound).
pub fn host_argminmax<T: CudaDType>(data: &[T], ws: &Workspace, mode: c_int) -> (usize, usize) {
    let mut out = [0u64; 2];
    check(
        unsafe {
            amm_host_argminmax(T::CODE, data.as_ptr() as *const c_void, data.len() as u64, mode, ws.raw,
                               out.as_mut_ptr())
        },
        "amm_host_argminmax",
    );
    (out[0] as usize, out[1] as usize)
}

/// One logical array split into consecutive index ranges over the GPUs of one box.  Each GPU
/// scans its shard; by default the SAME kernel then exchanges the 64-byte records with its peers
/// over NVLink (peer-mapped buffers) and folds them with a deterministic first-index tie-break,
/// so a call is one launch per GPU and no NCCL call.  `AMM_SHARDED_EXCHANGE=nccl` (and boxes
/// without peer access) use north_star's baseline instead: scan kernels, ONE `ncclAllGather` of
/// the records, host fold.  All of it happens inside `amm_sharded_argminmax`; every shard is
/// validated before anything is launched.
pub struct ShardedDeviceSlice<T: CudaDType> {
    raw: *mut AmmSharded,
    ptrs: Vec<*const c_void>,
    lens: Vec<u64>,
    _t: PhantomData<T>,
}
impl<T: CudaDType> ShardedDeviceSlice<T> {
    /// `shards[g] = (device, device pointer, length)`, consecutive ranges of one array.
    /// # Safety: every pointer must be valid on its device for its length.
    pub unsafe fn from_raw_shards(shards: &[(i32, *const T, usize)]) -> Self {
        let devices: Vec<c_int> = shards.iter().map(|s| s.0).collect();
        let mut raw = std::ptr::null_mut();
        check(amm_sharded_create(devices.len() as c_int, devices.as_ptr(), &mut raw), "amm_sharded_create");
        ShardedDeviceSlice {
            raw,
            ptrs: shards.iter().map(|s| s.1 as *const c_void).collect(),
            lens: shards.iter().map(|s| s.2 as u64).collect(),
            _t: PhantomData,
        }
    }
    fn run(&self, mode: c_int) -> (usize, usize) {
        let mut out = [0u64; 2];
        check(
            unsafe {
                amm_sharded_argminmax(self.raw, T::CODE, self.ptrs.as_ptr(), self.lens.as_ptr(), mode,
                                      out.as_mut_ptr())
            },
            "amm_sharded_argminmax",
        );
        (out[0] as usize, out[1] as usize)
    }
}
impl<T: CudaDType> Drop for ShardedDeviceSlice<T> {
    fn drop(&mut self) {
        unsafe { amm_sharded_destroy(self.raw) };
    }
}
impl<T: CudaDType> ArgMinMax for ShardedDeviceSlice<T> {
    fn argminmax(&self) -> (usize, usize) {
        self.run(AMM_IGNORE_NAN)
    }
    fn argmin(&self) -> usize {
        self.run(AMM_IGNORE_NAN).0
    }
    fn argmax(&self) -> usize {
        self.run(AMM_IGNORE_NAN).1
    }
}
macro_rules! impl_nan_sharded {
    ($($t:ty),*) => { $(
        impl NaNArgMinMax for ShardedDeviceSlice<$t> {
            fn nanargminmax(&self) -> (usize, usize) { self.run(AMM_RETURN_NAN) }
            fn nanargmin(&self) -> usize { self.run(AMM_RETURN_NAN).0 }
            fn nanargmax(&self) -> usize { self.run(AMM_RETURN_NAN).1 }
        }
    )* };
}
impl_nan_sharded!(f32, f64);
#[cfg(feature = "half")]
impl_nan_sharded!(half::f16);
