// Links the in-tree shared library (python -m loupiote_b200._build puts it under
// loupiote_b200/_lib).  LOUPIOTE_B200_LIB_DIR overrides the search directory.
use std::path::PathBuf;

fn main() {
    let dir = std::env::var("LOUPIOTE_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env!("CARGO_MANIFEST_DIR")).join("../../../loupiote_b200/_lib")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=loupiote_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=LOUPIOTE_B200_LIB_DIR");
}
