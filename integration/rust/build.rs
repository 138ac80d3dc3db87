// UNBUILT: no Rust toolchain exists in the image this was written in (cargo / rustc absent).  The text below is the
// code block of INTEGRATION.md, kept in step with it by tests/test_integration_shim.py, which also checks every
// `galah_b200_*` symbol named here against include/galah_b200.h and the built library.
fn main() {
    // libgalah_b200.so is built by `python -m galah_b200.build` (nvcc, sm_100a)
    println!("cargo:rustc-link-search=native={}", std::env::var("GALAH_B200_LIB_DIR").unwrap());
    println!("cargo:rustc-link-lib=dylib=galah_b200");
}
