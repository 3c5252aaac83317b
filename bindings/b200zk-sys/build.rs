// Links the prebuilt shared library.  B200ZK_LIB_DIR points at the directory holding libb200zk.so (built by
// `python -c "import __graft_entry__ as g; g.build()"` or the nvcc line in README.md); default: the in-tree location.
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("B200ZK_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../zkvm_prover_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=b200zk");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=B200ZK_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/b200zk.h");
}
