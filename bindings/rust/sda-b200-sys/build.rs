// Link against libsda_b200.so.  SDA_B200_LIB_DIR points at the directory that holds it (the repository's
// `sda_b200/` after `make -C sda_b200/csrc`); the library itself loads CUDA and, on demand, libnccl.so.2.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("SDA_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        // bindings/rust/sda-b200-sys -> <repo>/sda_b200
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../../sda_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=sda_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=SDA_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=build.rs");
}
