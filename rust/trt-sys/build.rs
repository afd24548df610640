// build.rs — compiles the CUDA sources for sm_100a with nvcc and links the result, replacing the reference's
// raytracer/build.rs (which shells out to `xcrun metal`, raytracer/build.rs:15-33).
// The list of translation units is tiny-raytracer_b200/csrc/sources.txt — the same file csrc/build.sh reads, so the two
// builds cannot drift apart (tests/test_capi_host.py::test_rust_build_lists_every_translation_unit).
// SOURCE ONLY in this repository: the image has no Rust toolchain, so this file has never been run.
use std::{env, fs, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("tiny-raytracer_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    // the numeric contract: no FMA contraction, IEEE division and square root, denormals kept — on both sides
    let dev = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
               "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-pthread"];
    let list = csrc.join("sources.txt");
    println!("cargo:rerun-if-changed={}", list.display());
    let mut objs = vec![];
    for line in fs::read_to_string(&list).expect("csrc/sources.txt").lines() {
        let line = line.trim();
        if line.is_empty() || line.starts_with('#') { continue; }
        let mut it = line.split_whitespace();
        let (src, kind) = (it.next().unwrap(), it.next().unwrap_or("cu"));
        let obj = out.join(format!("{src}.o"));
        let mut c = Command::new(&nvcc);
        c.args(dev);
        if kind == "cxx" { c.args(["-x", "cu"]); }          // host-only C++ goes through nvcc's host compiler with the same flags
        let st = c.arg("-c").arg(csrc.join(src)).arg("-o").arg(&obj).status().expect("nvcc not found");
        assert!(st.success(), "nvcc failed on {src}");
        objs.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(src).display());
    }
    for h in ["trt_device.cuh", "trt_layout.h", "trt_kernels.h", "trt_host_glue.cuh", "trt_scene.hpp"] {
        println!("cargo:rerun-if-changed={}", csrc.join(h).display());
    }
    let lib = out.join("libtrt_b200.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=trt_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=pthread");
}
