// Links libqlcuda.so (built by q-learning_b200/build.py with nvcc for sm_100a). QLCUDA_LIB_DIR points at its directory.
fn main() {
    let dir = std::env::var("QLCUDA_LIB_DIR").expect("set QLCUDA_LIB_DIR to the directory holding libqlcuda.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=qlcuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=QLCUDA_LIB_DIR");
}
