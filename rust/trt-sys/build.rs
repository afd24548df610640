// build.rs — compiles the CUDA sources for sm_100a with nvcc and links the result, replacing the reference's
// raytracer/build.rs (which shells out to `xcrun metal`, raytracer/build.rs:15-33).
// SOURCE ONLY in this repository: the image has no Rust toolchain, so this file has never been run.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("tiny-raytracer_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let dev = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
               "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-Xcompiler", "-fPIC,-ffp-contract=off"];
    let mut objs = vec![];
    for (src, is_cu) in [("trt_kernels.cu", true), ("trt_wavefront.cu", true), ("trt_capi.cpp", false), ("trt_scene.cpp", false)] {
        let obj = out.join(format!("{src}.o"));
        let mut c = Command::new(&nvcc);
        c.args(dev);
        if !is_cu { c.args(["-x", "cu"]); }
        let st = c.arg("-c").arg(csrc.join(src)).arg("-o").arg(&obj).status().expect("nvcc not found");
        assert!(st.success(), "nvcc failed on {src}");
        objs.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(src).display());
    }
    let lib = out.join("libtrt_b200.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=trt_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}
