// Links libshimmer_b200.so.  Two ways to get it:
//   * default: SHIMMER_B200_LIB_DIR points at a directory holding a prebuilt library
//     (python -c 'import __graft_entry__ as g; g.build()' leaves it in raytracinginoneweekendinrust_b200/lib);
//   * feature `build-cuda`: nvcc compiles the three sources for sm_100a into OUT_DIR.
// NOTE: written without a Rust toolchain at hand (the build image has none); see rust/README.md.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let repo = manifest.join("../..");
    let header = repo.join("include/shimmer_b200.h");
    println!("cargo:rerun-if-changed={}", header.display());
    println!("cargo:rerun-if-env-changed=SHIMMER_B200_LIB_DIR");

    let lib_dir = if cfg!(feature = "build-cuda") {
        let out = PathBuf::from(env::var("OUT_DIR").unwrap());
        let csrc = repo.join("raytracinginoneweekendinrust_b200/csrc");
        let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
        let status = Command::new(nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17"])
            .args(["-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden,-pthread", "-shared", "-o"])
            .arg(out.join("libshimmer_b200.so"))
            .arg(csrc.join("shim_api.cu"))
            .arg(csrc.join("shim_builder.cpp"))
            .arg(csrc.join("shim_scene.cpp"))
            .status()
            .expect("failed to run nvcc");
        assert!(status.success(), "nvcc failed");
        for f in ["shim_api.cu", "shim_builder.cpp", "shim_scene.cpp", "shim_device.h", "shim_kernels.cuh", "shim_types.h", "shim_scene.h"] {
            println!("cargo:rerun-if-changed={}", csrc.join(f).display());
        }
        out
    } else {
        PathBuf::from(env::var("SHIMMER_B200_LIB_DIR").unwrap_or_else(|_| {
            repo.join("raytracinginoneweekendinrust_b200/lib").to_string_lossy().into_owned()
        }))
    };
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=shimmer_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir.display());

    #[cfg(feature = "bindgen")]
    {
        let bindings = bindgen::Builder::default()
            .header(header.to_string_lossy())
            .allowlist_function("shim_.*")
            .allowlist_type("shim_.*")
            .allowlist_var("SHIM_.*")
            .generate()
            .expect("bindgen");
        bindings
            .write_to_file(PathBuf::from(env::var("OUT_DIR").unwrap()).join("bindings.rs"))
            .unwrap();
    }
}
