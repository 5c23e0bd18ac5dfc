// build.rs — compiles the CUDA sources for sm_100a with nvcc and links the result (plus cudart) into the crate.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let csrc = root.join("examples_b200").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libflk.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let status = Command::new(nvcc)
        .current_dir(&csrc)
        .args([
            "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
            "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
            "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-o",
        ])
        .arg(&lib)
        .args(["flk_kernels.cu", "flk_slots.cu", "flk_api.cu", "flk_dist.cu"])
        .status()
        .expect("nvcc not found: libflk has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=flk");
    // the library stays in OUT_DIR: give binaries and examples of this crate an rpath to it
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
    for f in ["flk_kernels.cu", "flk_slots.cu", "flk_api.cu", "flk_dist.cu", "flk_device.cuh", "flk_kernels.cuh", "flk_step.cuh",
              "flk_slots.cuh", "flk_internal.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include").join("flk.h").display());
}
