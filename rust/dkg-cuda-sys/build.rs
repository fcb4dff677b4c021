// Links against libdkgv.so (built by `python __graft_entry__.py` in the verifier's repository).
// DKGV_LIB_DIR = the directory that holds libdkgv.so (default: ../../dvt_circuits_b200 relative to this crate).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("DKGV_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../dvt_circuits_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=dkgv");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=DKGV_LIB_DIR");
}
