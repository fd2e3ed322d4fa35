// Builds libretto_b200.so with the repository's Makefile (nvcc, sm_100a) unless RETTO_B200_LIB_DIR points at a prebuilt one,
// then tells rustc where to find it.  src/lib.rs is GENERATED from include/retto_b200.h by tools/gen_rust_sys.py.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let repo = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let lib_dir = match env::var("RETTO_B200_LIB_DIR") {
        Ok(d) => PathBuf::from(d),
        Err(_) => {
            let status = Command::new("make").arg("-C").arg(repo.join("retto_b200/csrc")).arg("-j8").status().expect("make not found");
            assert!(status.success(), "building libretto_b200.so failed (needs nvcc with sm_100a support)");
            repo.join("retto_b200")
        }
    };
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=retto_b200");
    println!("cargo:rerun-if-changed={}", repo.join("include/retto_b200.h").display());
    println!("cargo:rerun-if-env-changed=RETTO_B200_LIB_DIR");
}
