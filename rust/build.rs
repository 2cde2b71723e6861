// build.rs -- compiled only with `--features cuda`.
//
// Drives nvcc for sm_100a over the CUDA sources of this repository (argminmax_b200/csrc) and
// links the resulting shared library.  NOT COMPILED IN THE BUILD CONTAINER (no cargo/rustc in
// the image); kept 1:1 with argminmax_b200/_build.py, which is what the tests actually run.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    if env::var("CARGO_FEATURE_CUDA").is_err() {
        return;
    }
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    // The CUDA sources live next to the crate when it is vendored into the reference tree as
    // `cuda/`; AMM_CSRC overrides.
    let csrc = env::var("AMM_CSRC")
        .map(PathBuf::from)
        .unwrap_or_else(|_| root.join("cuda").join("csrc"));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    let common = [
        "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "-Xcompiler", "-fPIC",
    ];
    let mut objs: Vec<PathBuf> = Vec::new();
    let mut compile = |src: &str, obj: &str, extra: &[&str]| {
        let o = out.join(obj);
        let status = Command::new(&nvcc)
            .args(common)
            .args(extra)
            .arg("-c")
            .arg(csrc.join(src))
            .arg("-o")
            .arg(&o)
            .status()
            .expect("failed to run nvcc");
        assert!(status.success(), "nvcc failed on {src}");
        objs.push(o);
    };
    compile("libamm.cu", "libamm.o", &[]);
    compile("host_slice.cu", "host_slice.o", &[]);
    compile("sharded.cu", "sharded.o", &[]);
    for g in 0..7 {
        // one translation unit per policy group: scan_inst.cu with -DAMM_GROUP=g
        let def = format!("-DAMM_GROUP={g}");
        compile("scan_inst.cu", &format!("scan_inst_g{g}.o"), &[def.as_str()]);
    }
    let lib = out.join("libargminmax_b200.so");
    let status = Command::new(&nvcc)
        .args(["-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o"])
        .arg(&lib)
        .args(&objs)
        .arg("-ldl")
        .status()
        .expect("failed to link");
    assert!(status.success(), "link failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=argminmax_b200");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
