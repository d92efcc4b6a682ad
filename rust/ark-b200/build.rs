// Links the prebuilt CUDA library (algebra_b200/libalgebra_b200.so, built by `make -C algebra_b200/csrc`).
fn main() {
    let dir = std::env::var("ALGEBRA_B200_LIB_DIR").unwrap_or_else(|_| "../../algebra_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=algebra_b200");
    println!("cargo:rerun-if-env-changed=ALGEBRA_B200_LIB_DIR");
}
