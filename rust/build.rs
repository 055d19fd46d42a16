// Links liblinfa_b200.so (built by `python -c "import __graft_entry__ as g; g.build()"` -> linfa_linalg_b200/lib/).
// LINFA_B200_LIB_DIR overrides the search path.  libnccl.so.2 is NOT linked: the library dlopen()s it only when a
// multi-device handle is created (lfb_create_multi).
fn main() {
    let dir = std::env::var("LINFA_B200_LIB_DIR").unwrap_or_else(|_| "../linfa_linalg_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=linfa_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=LINFA_B200_LIB_DIR");
}
