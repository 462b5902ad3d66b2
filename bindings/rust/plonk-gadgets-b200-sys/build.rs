// SOURCE ONLY (never compiled here).  Builds libpg_b200.so with nvcc for sm_100a and links it.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../..");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libpg_b200.so");
    let status = Command::new("nvcc")
        .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .arg(root.join("plonk_gadgets_b200/csrc/engine.cu"))
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=pg_b200");
    println!("cargo:rerun-if-changed={}", root.join("plonk_gadgets_b200/csrc").display());
    println!("cargo:rerun-if-changed={}", root.join("include/pg_b200.h").display());
}
