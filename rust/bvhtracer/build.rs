// build.rs -- link libbvht_cuda.so (built by `python -m bvhtracer_b200.build` in the bvht-b200 repository) when the `cuda`
// feature is on.  Point BVHT_LIB_DIR at <bvht-b200>/bvhtracer_b200/lib.  Without the feature this script does nothing, so the
// pure-CPU crate (and `--example dump_hits`) builds exactly as before.
fn main() {
    println!("cargo:rerun-if-env-changed=BVHT_LIB_DIR");
    if std::env::var_os("CARGO_FEATURE_CUDA").is_none() {
        return;
    }
    let dir = std::env::var("BVHT_LIB_DIR").expect("set BVHT_LIB_DIR to the directory holding libbvht_cuda.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=bvht_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
}
