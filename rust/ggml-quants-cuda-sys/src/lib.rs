//! `extern "C"` view of include/ggq.h — keep in step with gguf_b200/_lib.py: SYMBOLS (tests/test_abi.py checks that
//! list against the header and the exported symbols of libggq.so).
use std::ffi::{c_char, c_int, c_void};

#[repr(C)]
pub struct GgqSliceJob {
    pub ty: u32,
    pub fdt: u32,
    pub quantize: c_int,
    pub dst: *mut c_void,
    pub dst_len: usize,
    pub src: *const c_void,
    pub src_len: usize,
}

#[repr(C)]
pub struct GgqShardPiece {
    pub job: u32,
    pub device: c_int,
    pub elem_begin: usize,
    pub elem_end: usize,
}

#[repr(C)]
pub struct GgqLayout {
    pub ndim: u32,
    pub shape: [u64; 4],
    pub strides: [i64; 4],
    pub offset: i64,
}

#[repr(C)]
#[derive(Default)]
pub struct GgqConvertStats {
    pub n_tensors: u64,
    pub n_cast_tensors: u64,
    pub cast_elems: u64,
    pub bytes_in: u64,
    pub bytes_out: u64,
    pub seconds_plan: f64,
    pub seconds_convert: f64,
    pub seconds_sync: f64,
    pub n_devices: c_int,
    pub n_out_files: c_int,
    pub n_rearranged_tensors: u64,
    pub n_workers: c_int,
    pub worker_seconds_read: f64,
    pub worker_seconds_write: f64,
    pub worker_seconds_gpu_wait: f64,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
    pub n_direct_inputs: c_int,
}

#[repr(C)]
#[derive(Default)]
pub struct GgqConvertOptions {
    pub n_devices: c_int,
    pub max_tensors: u64,
    pub max_bytes: u64,
    pub no_tensor_first: c_int,
    pub no_data: c_int,
    pub direct_io: c_int,
}

extern "C" {
    pub fn ggq_block_info(ty: u32, elems: *mut u32, bytes: *mut u32) -> c_int;
    pub fn ggq_last_error() -> *const c_char;
    pub fn ggq_device_count() -> c_int;
    pub fn ggq_set_device(device: c_int) -> c_int;
    pub fn ggq_set_shard_devices(n_devices: c_int) -> c_int;
    pub fn ggq_quantize_slice(ty: u32, fdt: u32, dst: *mut c_void, dst_blocks: usize, src: *const c_void, src_elems: usize) -> c_int;
    pub fn ggq_dequantize_slice(ty: u32, fdt: u32, dst: *mut c_void, dst_elems: usize, src: *const c_void, src_blocks: usize) -> c_int;
    pub fn ggq_slices(jobs: *const GgqSliceJob, n_jobs: usize) -> c_int;
    pub fn ggq_plan_shards(jobs: *const GgqSliceJob, n_jobs: usize, n_devices: c_int, out: *mut GgqShardPiece, cap: usize) -> usize;
    pub fn ggq_quantize_slice_device(ty: u32, fdt: u32, dst: *mut c_void, dst_blocks: usize, src: *const c_void, src_elems: usize, stream: *mut c_void) -> c_int;
    pub fn ggq_dequantize_slice_device(ty: u32, fdt: u32, dst: *mut c_void, dst_elems: usize, src: *const c_void, src_blocks: usize, stream: *mut c_void) -> c_int;
    pub fn ggq_slices_device(jobs: *const GgqSliceJob, n_jobs: usize, stream: *mut c_void) -> c_int;
    pub fn ggq_cast(types: *const u32, n_types: c_int, dst: *mut c_void, src: *const c_void, n_elems: usize) -> c_int;
    pub fn ggq_type_nbytes(ty: u32, n_elems: usize) -> usize;
    pub fn ggq_rearrange_device(dst: *mut c_void, dst_layout: *const GgqLayout, src: *const c_void, src_layout: *const GgqLayout, unit: usize, stream: *mut c_void) -> c_int;
    pub fn ggq_rearrange(dst: *mut c_void, dst_layout: *const GgqLayout, src: *const c_void, src_layout: *const GgqLayout, unit: usize) -> c_int;
    pub fn ggq_convert_gguf(in_path: *const c_char, out_path: *const c_char, steps: *const c_char, n_devices: c_int, stats: *mut GgqConvertStats) -> c_int;
    pub fn ggq_convert_gguf_ex(in_paths: *const *const c_char, n_in: usize, out_path: *const c_char, steps: *const c_char, opts: *const GgqConvertOptions, stats: *mut GgqConvertStats) -> c_int;
    pub fn ggq_convert_last_error() -> *const c_char;
    pub fn ggq_host_alloc(bytes: usize) -> *mut c_void;
    pub fn ggq_host_free(p: *mut c_void);
    pub fn ggq_shutdown();
    pub fn ggq_launch_count() -> u64;
    pub fn ggq_version() -> *const c_char;
}
