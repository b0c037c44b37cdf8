// Links libzkfhe_b200.so (built by `python zk-fhe_b200/build.py`).  ZKFHE_B200_ROOT points at the repository root.
fn main() {
    let root = std::env::var("ZKFHE_B200_ROOT").unwrap_or_else(|_| "../..".to_string());
    println!("cargo:rustc-link-search=native={root}/zk-fhe_b200/lib");
    println!("cargo:rustc-link-lib=dylib=zkfhe_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{root}/zk-fhe_b200/lib");
    println!("cargo:rerun-if-env-changed=ZKFHE_B200_ROOT");
    println!("cargo:rerun-if-changed=../../include/zkfhe_b200.h");
}
