// Links the prebuilt native library. Point STRINGWARS_B200_LIB_DIR at <repo>/stringwars_b200/lib
// (built by `python -m stringwars_b200.build`, i.e. one nvcc invocation for sm_100a).
fn main() {
    let dir = std::env::var("STRINGWARS_B200_LIB_DIR").expect("set STRINGWARS_B200_LIB_DIR to the directory holding libstringwars_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=stringwars_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=STRINGWARS_B200_LIB_DIR");
}
