// build.rs of the chrono-photo crate once it links libchrono_b200.so.
// The library is built by `python -m chrono_photo_b200.build` (nvcc, sm_100a); CHRONO_B200_LIB_DIR points at the directory
// that holds it (chrono_photo_b200/ of this repository).
fn main() {
    let dir = std::env::var("CHRONO_B200_LIB_DIR").expect("set CHRONO_B200_LIB_DIR to the directory of libchrono_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=chrono_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=CHRONO_B200_LIB_DIR");
}
