//! B200-native backend behind the reference's module tree (src/lib.rs:1-2 upstream).  NOT COMPILED in this repository's
//! image (no Rust toolchain): see rust/README.md.
pub mod device;
pub mod inference_engine;
pub mod inference_fp32_ops;
