fn main() {
    // libeuc_b200.so is built by `python -c "import __graft_entry__ as g; g.build()"` into euc_b200/csrc/
    if let Ok(dir) = std::env::var("EUC_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=euc_b200");
}
