// Compiles the CUDA sources of this repository (rusterix_b200/csrc) for sm_100a and links them statically, or links a
// prebuilt librxcuda.so when RXCUDA_LIB_DIR points at one (what `python __graft_entry__.py` builds in-tree).
use std::{env, path::PathBuf};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    if let Ok(dir) = env::var("RXCUDA_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=rxcuda");
        return;
    }
    let csrc = root.join("rusterix_b200/csrc");
    cc::Build::new()
        .cuda(true)
        .flag("-gencode").flag("arch=compute_100a,code=sm_100a")
        .flag("-std=c++17").flag("-O3").flag("-lineinfo")
        .flag("-fmad=false") // Rust never contracts a*b+c; the kernels use explicit FMAs where the reference has mul_add
        .include(root.join("include"))
        .include(&csrc)
        .file(csrc.join("rx_kernels.cu"))
        .file(csrc.join("rx_api.cu"))
        .compile("rxcuda");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include/rxcuda.h").display());
}
