// build.rs -- compiles the CUDA sources of the batch engine with nvcc for sm_100a and links them statically.
// SOURCE-ONLY DELIVERABLE (no cargo in the build image).  Expects the engine sources under `gpu/csrc`
// (= capycrypt_b200/csrc of this repository) and the header under `gpu/include`.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let srcs = ["ctx.cu", "sha3_api.cu", "ed448_api.cu", "ed448_fixed.cu", "ed448_var.cu", "ae_api.cu"];
    let mut objs = vec![];
    for s in srcs {
        let obj = out.join(s.replace(".cu", ".o"));
        let mut cmd = Command::new(&nvcc);
        cmd.args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                  "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-c"]);
        let status = cmd.arg(format!("gpu/csrc/{s}")).arg("-o").arg(&obj).status().expect("nvcc not found");
        assert!(status.success(), "nvcc failed on {s} (there is no CPU fallback)");
        objs.push(obj);
        println!("cargo:rerun-if-changed=gpu/csrc/{s}");
    }
    let lib = out.join("libcapycrypt_gpu.a");
    let status = Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap();
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=capycrypt_gpu");
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rustc-link-lib=stdc++");
}
