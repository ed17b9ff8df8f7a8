// Links libdensity_b200.so (built by `python -m density_b200.build`, nvcc -gencode arch=compute_100a,code=sm_100a).
fn main() {
    let dir = std::env::var("DENSITY_B200_LIB_DIR").expect("set DENSITY_B200_LIB_DIR to the directory that holds libdensity_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=density_b200");
    println!("cargo:rerun-if-env-changed=DENSITY_B200_LIB_DIR");
}
