// Links libracer_cuda.so.  RACER_CUDA_LIB_DIR points at the directory that holds it
// (racer_tracer_b200/ in this repository).
fn main() {
    let dir = std::env::var("RACER_CUDA_LIB_DIR").unwrap_or_else(|_| "../../".into());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=racer_cuda");
    println!("cargo:rerun-if-env-changed=RACER_CUDA_LIB_DIR");
}
