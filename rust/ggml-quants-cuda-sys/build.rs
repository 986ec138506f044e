// build.rs — compile the hand-written sm_100a kernels of this repository into libggq.a and link it, or (feature
// "prebuilt") link the gguf_b200/libggq.so the repository's Makefile produced.  The source list and flags mirror
// gguf_b200/csrc/Makefile; -fmad=false is REQUIRED (bit parity with the reference's unfused Rust arithmetic).
// CARGO_MANIFEST_DIR is rust/ggml-quants-cuda-sys, the repository root two levels up.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..").canonicalize().unwrap();
    let csrc = root.join("gguf_b200/csrc");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include/ggq.h").display());
    if env::var_os("CARGO_FEATURE_PREBUILT").is_some() {
        println!("cargo:rustc-link-search=native={}", root.join("gguf_b200").display());
        println!("cargo:rustc-link-lib=dylib=ggq");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", root.join("gguf_b200").display());
        return;
    }
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let mut objs = vec![];
    for s in ["api.cu", "dequant.cu", "quant_legacy.cu", "quant_k.cu", "rearrange.cu"] {
        let o = out.join(s).with_extension("o");
        let st = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo"])
            .args(["-fmad=false", "--ftz=false", "--prec-div=true", "--prec-sqrt=true"])
            .args(["-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(s))
            .arg("-o")
            .arg(&o)
            .status()
            .expect("nvcc not found (set NVCC or use --features prebuilt)");
        assert!(st.success(), "nvcc failed on {s}");
        objs.push(o);
    }
    for s in ["convert.cpp", "host_copy.cpp"] {
        let o = out.join(s).with_extension("o");
        let st = Command::new(env::var("CXX").unwrap_or_else(|_| "g++".into()))
            .args(["-O2", "-std=c++17", "-fPIC", "-c"])
            .arg(csrc.join(s))
            .arg("-o")
            .arg(&o)
            .status()
            .expect("g++ not found");
        assert!(st.success(), "g++ failed on {s}");
        objs.push(o);
    }
    let lib = out.join("libggq.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=ggq");
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}
