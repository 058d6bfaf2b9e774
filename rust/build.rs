// Compiles the CUDA kernels and the C ABI (../maxpro_b200/csrc) with nvcc for sm_100a and links
// the shared library.  No cudarc, no bindgen: the FFI surface is the dozen functions of src/ffi.rs.
use std::path::PathBuf;
use std::process::Command;

const SOURCES: &[&str] = &[
    "api.cu", "config.cu", "search_small.cu", "search_cyclic.cu", "search_screen.cu", "search_cta.cu", "search_warp.cu",
    "pair_generic.cu", "generate.cu", "anneal.cu", "anneal_pipe.cu", "anneal_maximin.cu", "order.cu", "order_cluster.cu", "order_cluster_async.cu", "probe.cu",
];

fn main() {
    let out = PathBuf::from(std::env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap()).join("../maxpro_b200/csrc");
    let nvcc = std::env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let arch = ["-gencode", "arch=compute_100a,code=sm_100a"];
    let mut objects = Vec::new();
    for src in SOURCES {
        let obj = out.join(format!("{src}.o"));
        let status = Command::new(&nvcc)
            .args(arch)
            .args(["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-c"])
            .arg(csrc.join(src))
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("nvcc not found (set NVCC)");
        assert!(status.success(), "nvcc failed on {src}");
        println!("cargo:rerun-if-changed={}", csrc.join(src).display());
        objects.push(obj);
    }
    let lib = out.join("libmaxpro_b200.so");
    let status = Command::new(&nvcc).args(arch).arg("-shared").arg("-o").arg(&lib).args(&objects).status().unwrap();
    assert!(status.success(), "nvcc failed to link libmaxpro_b200.so");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=maxpro_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
}
