// Builds the CUDA engine for sm_100a with nvcc and links it statically (the recipe of csrc/Makefile).
// NOT COMPILED in the engine's own build image (no cargo / rustc there); the same objects are built by the Makefile.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let src = PathBuf::from(env::var("BENDY_B200_CSRC").unwrap_or_else(|_| "csrc".into()));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut objs = vec![];
    // kernels.cu is compiled twice: the fast arithmetic flavour (FMA contraction on) and the bit-exact one
    // (-fmad=false -DBT_EXACT_SCAN); everything else is built -fmad=false
    for (f, o, fmad, extra) in [
        ("engine.cu", "engine.o", "-fmad=false", None),
        ("kernels.cu", "kernels.o", "-fmad=true", None),
        ("kernels.cu", "kernels_exact.o", "-fmad=false", Some("-DBT_EXACT_SCAN")),
        ("scene.cpp", "scene.o", "-fmad=false", None),
    ] {
        let o = out.join(o);
        let mut cmd = Command::new(&nvcc);
        cmd.args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", fmad]);
        cmd.args(["-Xcompiler", "-fPIC,-ffp-contract=off", "-x", "cu", "-c"]);
        if let Some(d) = extra {
            cmd.arg(d);
        }
        let ok = cmd.arg(src.join(f)).arg("-o").arg(&o).status().expect("nvcc not found").success();
        assert!(ok, "nvcc failed on {f}");
        objs.push(o);
        println!("cargo:rerun-if-changed={}", src.join(f).display());
    }
    for h in ["device.cuh", "render_pool.cuh", "kernels.h", "layout.h", "scene.hpp"] {
        println!("cargo:rerun-if-changed={}", src.join(h).display());
    }
    let lib = out.join("libbendy_b200.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=bendy_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rustc-link-lib=z");
    println!("cargo:rustc-link-lib=stdc++");
}
