// Links libb200zk.so (built by `make -C zcash-gpu-thesis_b200/csrc`).  B200ZK_LIB_DIR = the directory that holds it.
fn main() {
    let dir = std::env::var("B200ZK_LIB_DIR").expect("set B200ZK_LIB_DIR to the directory of libb200zk.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=b200zk");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=B200ZK_LIB_DIR");
}
