// Links the pyo3 module against libssqcuda.so (built by `python -c "import __graft_entry__ as g; g.build()"`).
fn main() {
    let dir = std::env::var("SSQCUDA_DIR")
        .expect("set SSQCUDA_DIR to the directory holding libssqcuda.so (ssqueeze_rs_b200/)");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=ssqcuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=SSQCUDA_DIR");
}
