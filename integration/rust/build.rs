// Links libpqv.so (built by `make -C pq_vector_b200/csrc`, sm_100a).  PQV_LIB_DIR = the directory holding it.
fn main() {
    let dir = std::env::var("PQV_LIB_DIR").expect("set PQV_LIB_DIR to the directory of libpqv.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=pqv");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=PQV_LIB_DIR");
}
