// Links libchrom_cuda.so. Set CHROM_CUDA_LIB_DIR to the directory that holds it
// (in this repository: chromatography_b200/).
fn main() {
    if let Ok(dir) = std::env::var("CHROM_CUDA_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=chrom_cuda");
    println!("cargo:rerun-if-env-changed=CHROM_CUDA_LIB_DIR");
}
