// Links libzkb200.so built by `python zktls_b200/build.py` (nvcc, sm_100a).  ZKB200_LIB_DIR points at the directory holding it.
fn main() {
    let dir = std::env::var("ZKB200_LIB_DIR").unwrap_or_else(|_| "../../../zktls_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=zkb200");
    println!("cargo:rerun-if-env-changed=ZKB200_LIB_DIR");
}
