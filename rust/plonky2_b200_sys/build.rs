// Links libp2b.so (built by `make -C city_rollup_b200/csrc`); P2B_LIB_DIR points at the directory
// holding it.
fn main() {
    let dir = std::env::var("P2B_LIB_DIR").unwrap_or_else(|_| "../../city_rollup_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=p2b");
    println!("cargo:rerun-if-env-changed=P2B_LIB_DIR");
}
