// Link against libqfall_b200.so built by `python -m tools_b200.build` (nvcc, sm_100a).
// QFALL_B200_LIB_DIR = directory that holds the library (default: ../../tools_b200).
fn main() {
    let dir = std::env::var("QFALL_B200_LIB_DIR").unwrap_or_else(|_| {
        let here = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{here}/../../tools_b200")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=qfall_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=QFALL_B200_LIB_DIR");
}
