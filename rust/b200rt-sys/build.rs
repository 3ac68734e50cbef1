// libb200rt.so is built by `make -C onnx_rusty_inference_engine_b200/csrc` (nvcc, sm_100a only).
fn main() {
    let dir = std::env::var("B200RT_LIB_DIR").expect("set B200RT_LIB_DIR to the directory holding libb200rt.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=b200rt");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=B200RT_LIB_DIR");
}
