// Links libmole_b200.so (built by mole_b200/csrc/build.sh).  MOLE_B200_LIB_DIR overrides the search path.
fn main() {
    let dir = std::env::var("MOLE_B200_LIB_DIR").unwrap_or_else(|_| "../../mole_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=mole_b200");
    println!("cargo:rerun-if-env-changed=MOLE_B200_LIB_DIR");
}
