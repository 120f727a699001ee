// Links the prebuilt libspittle_b200.so (python -m spittle_b200.build).  SPITTLE_B200_LIB_DIR points at
// the directory holding it.
fn main() {
    let dir = std::env::var("SPITTLE_B200_LIB_DIR").unwrap_or_else(|_| "../../spittle_b200".into());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=spittle_b200");
}
