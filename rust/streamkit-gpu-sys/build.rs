// UNCOMPILED -- see rust/README.md.
// Links the C ABI libraries of streamkit_b200 (built by `python -c "import __graft_entry__ as g; g.build()"`):
//   libskgpu.so         include/skgpu_batch.h   (CUDA kernels + batch ABI)
//   libskgpu_hub.so     include/skgpu_hub.h     (frame-batching layer)
//   libskgpu_router.so  include/skgpu_router.h  (single-process multi-GPU router)
// SKGPU_LIB_DIR points at streamkit_b200/csrc of a built checkout; the libraries carry rpath $ORIGIN for each other.
use std::env;

fn main() {
    println!("cargo:rerun-if-env-changed=SKGPU_LIB_DIR");
    if let Ok(dir) = env::var("SKGPU_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    for lib in ["skgpu", "skgpu_hub", "skgpu_router"] {
        println!("cargo:rustc-link-lib=dylib={lib}");
    }
}
