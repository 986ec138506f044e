//! Rust binding of libggq (`include/ggq.h`): the B200 implementation of the `ggml-quants` slice API.
//!
//! * [`ffi`] — `extern "C"` declarations, one per entry point of `include/ggq.h` (crate ggml-quants-cuda-sys).
//! * [`CudaQuantExt`] — the signatures of `ggml_quants::QuantExt` (ggml-quants/src/lib.rs:98-104), implemented
//!   for every block type of the reference by forwarding to libggq.  It exists beside the reference's blanket
//!   impl so that one binary can run both and compare them (`tests/parity.rs`).
//! * `drop-in.patch` (next to this crate's Cargo.toml) — the change to `ggml-quants/src/lib.rs` that makes
//!   the GPU path THE `QuantExt` impl behind a `cuda` feature: the existing blanket impl (lib.rs:116-148)
//!   gets `#[cfg(not(feature = "cuda"))]` — two blanket impls of one trait would collide — and the impl
//!   below is added under `#[cfg(feature = "cuda")]`.
#![deny(warnings)]

pub use ggml_quants_cuda_sys as ffi;

use ffi::*;
pub use ggml_quants::{bf16, f16, QuantizeError};
use ggml_quants::{Q2K, Q3K, Q4K, Q4_0, Q4_1, Q5K, Q5_0, Q5_1, Q6K, Q8K, Q8_0, Q8_1};

/// `GGmlType` discriminant of the float side `T` of `Quantize<T, N>` (ggus/src/tensor.rs:15-50).
pub trait FloatSide: Copy + Send + Sync + 'static {
    const FDT: u32;
}
impl FloatSide for f32 {
    const FDT: u32 = 0;
}
impl FloatSide for f16 {
    const FDT: u32 = 1;
}
impl FloatSide for bf16 {
    const FDT: u32 = 30;
}

/// `GGmlType` discriminant and element count of a block type.
pub trait GgmlBlock: Sized + Send + Sync + 'static {
    const TY: u32;
    const COUNT: usize;
}
macro_rules! block {
    ($($t:ty = $v:expr, $n:expr;)*) => { $(impl GgmlBlock for $t { const TY: u32 = $v; const COUNT: usize = $n; })* };
}
block! {
    f16 = 1, 1; bf16 = 30, 1;
    Q4_0 = 2, 32; Q4_1 = 3, 32; Q5_0 = 6, 32; Q5_1 = 7, 32; Q8_0 = 8, 32; Q8_1 = 9, 32;
    Q2K = 10, 256; Q3K = 11, 256; Q4K = 12, 256; Q5K = 13, 256; Q6K = 14, 256; Q8K = 15, 256;
}

fn status(rc: i32) -> Result<(), QuantizeError> {
    match rc {
        0 => Ok(()),
        1 => Err(QuantizeError::Indivisible),    // lib.rs:122-124 / 136-138: checked first
        2 => Err(QuantizeError::LengthMismatch), // lib.rs:125-127 / 139-141
        e => panic!("libggq: status {e}: {}", last_error()),
    }
}

/// The calling thread's last libggq error message.
pub fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(ggq_last_error()) }.to_string_lossy().into_owned()
}

/// `QuantExt<T, N>` (ggml-quants/src/lib.rs:98-104) on the GPU.  Same contract: host slices, synchronous, the two
/// length checks in the reference's order, every element of `dst` written on return.
pub trait CudaQuantExt<T>: Sized {
    fn quantize_slice(dst: &mut [Self], src: &[T]) -> Result<(), QuantizeError>;
    fn dequantize_slice(dst: &mut [T], src: &[Self]) -> Result<(), QuantizeError>;
}

impl<Blk: GgmlBlock, T: FloatSide> CudaQuantExt<T> for Blk {
    fn quantize_slice(dst: &mut [Self], src: &[T]) -> Result<(), QuantizeError> {
        debug_assert_eq!(std::mem::size_of::<Blk>(), block_bytes(Blk::TY));
        status(unsafe { ggq_quantize_slice(Blk::TY, T::FDT, dst.as_mut_ptr().cast(), dst.len(), src.as_ptr().cast(), src.len()) })
    }
    fn dequantize_slice(dst: &mut [T], src: &[Self]) -> Result<(), QuantizeError> {
        status(unsafe { ggq_dequantize_slice(Blk::TY, T::FDT, dst.as_mut_ptr().cast(), dst.len(), src.as_ptr().cast(), src.len()) })
    }
}

/// `size_of::<Blk>()` as libggq sees it (`ggq_block_info`).
pub fn block_bytes(ty: u32) -> usize {
    let (mut e, mut b) = (0u32, 0u32);
    assert_eq!(unsafe { ggq_block_info(ty, &mut e, &mut b) }, 0, "{}", last_error());
    b as usize
}

/// Bytes of a slice of blocks (`repr(C)` structs without padding: structs/*.rs).
pub fn bytes_of<B>(blocks: &[B]) -> &[u8] {
    unsafe { std::slice::from_raw_parts(blocks.as_ptr().cast(), std::mem::size_of_val(blocks)) }
}
