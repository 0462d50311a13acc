// Links libstratum_b200.so.  STRATUM_B200_LIB_DIR = directory holding the library
// (stratum_dsp_b200/_build after `python -c "import __graft_entry__ as g; g.build()"`).
fn main() {
    if let Ok(dir) = std::env::var("STRATUM_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=stratum_b200");
    println!("cargo:rerun-if-env-changed=STRATUM_B200_LIB_DIR");
}
