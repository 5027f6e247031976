// build.rs -- builds the CUDA engine with nvcc for sm_100a and links it (see rust/README.md: not built in this image).
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").expect("OUT_DIR"));
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").expect("CARGO_MANIFEST_DIR"));
    let src = root.join("aeonflux_b200/csrc/afx_b200.cu");
    let lib = out.join("libaeonflux_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    let status = Command::new(nvcc)
        .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-o"])
        .arg(&lib)
        .arg(&src)
        .status()
        .expect("failed to run nvcc");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=aeonflux_b200");
    println!("cargo:rerun-if-changed={}", root.join("aeonflux_b200/csrc").display());
    println!("cargo:rerun-if-changed={}", root.join("include/aeonflux_b200.h").display());
}
