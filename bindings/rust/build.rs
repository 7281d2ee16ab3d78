// build.rs -- compiles the CUDA library for sm_100a with nvcc and links it (north star: "Rust host code calls
// hand-written sm_100a CUDA through a thin extern \"C\" FFI layer built by build.rs with nvcc").
//
//   NVCC                 path of nvcc (default: nvcc on PATH, then /usr/local/cuda/bin/nvcc)
//   TEKKEN_B200_SRC      directory that holds tk_kernels.cu ... (default: ../../tekken_rs_b200/csrc)
//   TEKKEN_B200_LIB_DIR  with feature "prebuilt": directory of an already built libtekken_b200.so
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").expect("OUT_DIR"));
    if env::var("CARGO_FEATURE_PREBUILT").is_ok() {
        let dir = env::var("TEKKEN_B200_LIB_DIR").expect("feature `prebuilt` needs TEKKEN_B200_LIB_DIR");
        println!("cargo:rustc-link-search=native={dir}");
    } else {
        let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
        let src = env::var("TEKKEN_B200_SRC").map(PathBuf::from).unwrap_or_else(|_| manifest.join("../../tekken_rs_b200/csrc"));
        let files = ["tk_kernels.cu", "tk_decode.cu", "tk_api.cu", "tk_host.cpp"];
        for f in files.iter().chain(["tk_common.h", "tk_device.cuh", "tk_pretok.h", "tk_pretok_cfg.h", "tk_small.cuh", "tk_kernels.h", "tk_host.h"].iter()) {
            println!("cargo:rerun-if-changed={}", src.join(f).display());
        }
        let nvcc = env::var("NVCC").unwrap_or_else(|_| {
            if Command::new("nvcc").arg("--version").output().is_ok() { "nvcc".into() } else { "/usr/local/cuda/bin/nvcc".into() }
        });
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o"])
            .arg(out.join("libtekken_b200.so"))
            .args(files.iter().map(|f| src.join(f)))
            .status()
            .unwrap_or_else(|e| panic!("cannot run {nvcc}: {e} (the B200 path has no CPU fallback; nvcc is required)"));
        assert!(status.success(), "nvcc failed");
        println!("cargo:rustc-link-search=native={}", out.display());
        // let the test binaries find the library without LD_LIBRARY_PATH
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
    }
    println!("cargo:rustc-link-lib=dylib=tekken_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
}
