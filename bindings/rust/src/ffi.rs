//! Raw binding of include/chrono_b200.h (the C ABI of libchrono_b200.so): the file a maintainer drops into the reference as
//! `src/ffi.rs`. Written against the header; NOT compiled in this repository's image (no cargo / rustc there). The
//! `const _: () = assert!(..)` lines pin every struct's size and field offsets to the values tests/test_abi_layout.py derives
//! from the C header with gcc and from the ctypes mirror -- a layout drift fails either `cargo build` or that test.
#![allow(non_camel_case_types, dead_code)]
use std::mem::size_of;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct ChbCtx { _p: [u8; 0] }
#[repr(C)] pub struct ChbStack { _p: [u8; 0] }
#[repr(C)] pub struct ChbShake { _p: [u8; 0] }

pub const CHB_OK: c_int = 0;
pub const CHB_ERR_INVALID: c_int = 1;
pub const CHB_ERR_CUDA: c_int = 2;
pub const CHB_ERR_UNSUPPORTED: c_int = 3;
pub const CHB_ERR_STATE: c_int = 4;

/// chb_fade (Fade, src/options.rs:59-66: the already-built LUT)
#[repr(C)]
#[derive(Clone, Copy)]
pub struct ChbFade {
    pub is_none: u8,
    pub mode: u8,
    pub absolute: u8,
    pub _pad: u8,
    pub offset: i32,
    pub n_values: i32,
    pub values: *const f32,
}

/// chb_outlier_params (arguments of OutlierProcessor::new, src/chrono.rs:46-54)
#[repr(C)]
#[derive(Clone, Copy)]
pub struct ChbOutlierParams {
    pub thr_absolute: u8,
    pub background: u8,
    pub outlier: u8,
    pub _pad: u8,
    pub thr_min: f32,
    pub thr_max: f32,
    pub thr_scale: f32,
    pub weights: [f32; 4],
    pub fade: ChbFade,
    pub sample_count: i32,
    pub seed: u64,
    pub pixel_offset: u64,
    /// interleaved row-block shards: pixels per block and pixels between the ends and starts of consecutive owned blocks (0 / 0: one band)
    pub block_pixels: u64,
    pub block_skip: u64,
}

/// chb_simple_params (arguments of SimpleProcessor::new, src/simple.rs:18)
#[repr(C)]
#[derive(Clone, Copy)]
pub struct ChbSimpleParams {
    pub darker: u8,
    pub _pad: [u8; 3],
    pub weights: [f32; 4],
    pub fade: ChbFade,
}

/// chb_debug_planes
#[repr(C)]
pub struct ChbDebugPlanes {
    pub median: *mut f32,
    pub q1: *mut f32,
    pub q3: *mut f32,
    pub n_outliers: *mut i32,
}

// ---- layout pins (LP64): sizes and offsets as include/chrono_b200.h lays them out with gcc on x86-64 / aarch64
macro_rules! offset_of {
    ($t:ty, $f:ident) => {{
        let u = std::mem::MaybeUninit::<$t>::uninit();
        let base = u.as_ptr();
        unsafe { (std::ptr::addr_of!((*base).$f) as *const u8).offset_from(base as *const u8) as usize }
    }};
}
const _: () = assert!(size_of::<ChbFade>() == 24);
const _: () = assert!(size_of::<ChbOutlierParams>() == 96);
const _: () = assert!(size_of::<ChbSimpleParams>() == 48);
const _: () = assert!(size_of::<ChbDebugPlanes>() == 32);
#[cfg(test)]
mod layout {
    use super::*;
    #[test]
    fn offsets_match_the_c_header() {
        assert_eq!(offset_of!(ChbFade, offset), 4);
        assert_eq!(offset_of!(ChbFade, n_values), 8);
        assert_eq!(offset_of!(ChbFade, values), 16);
        assert_eq!(offset_of!(ChbOutlierParams, thr_min), 4);
        assert_eq!(offset_of!(ChbOutlierParams, weights), 16);
        assert_eq!(offset_of!(ChbOutlierParams, fade), 32);
        assert_eq!(offset_of!(ChbOutlierParams, sample_count), 56);
        assert_eq!(offset_of!(ChbOutlierParams, seed), 64);
        assert_eq!(offset_of!(ChbOutlierParams, pixel_offset), 72);
        assert_eq!(offset_of!(ChbOutlierParams, block_pixels), 80);
        assert_eq!(offset_of!(ChbOutlierParams, block_skip), 88);
        assert_eq!(offset_of!(ChbSimpleParams, weights), 4);
        assert_eq!(offset_of!(ChbSimpleParams, fade), 24);
    }
}

extern "C" {
    pub fn chb_last_error() -> *const c_char;
    pub fn chb_version() -> c_int;
    pub fn chb_ctx_create(device_ids: *const c_int, n_dev: c_int, out: *mut *mut ChbCtx) -> c_int;
    pub fn chb_ctx_destroy(ctx: *mut ChbCtx) -> c_int;
    pub fn chb_ctx_device_count(ctx: *const ChbCtx) -> c_int;
    pub fn chb_ctx_mem_info(ctx: *mut ChbCtx, dev_slot: c_int, free_bytes: *mut usize, total_bytes: *mut usize) -> c_int;
    pub fn chb_stack_create(ctx: *mut ChbCtx, width: c_int, height: c_int, channels: c_int, n_frames: c_int, out: *mut *mut ChbStack) -> c_int;
    pub fn chb_stack_destroy(stack: *mut ChbStack) -> c_int;
    pub fn chb_stack_upload(stack: *mut ChbStack, frame_idx: c_int, host_pixels: *const u8, row_pitch: usize, crop_x: c_int, crop_y: c_int) -> c_int;
    pub fn chb_stack_upload_pinned(stack: *mut ChbStack, frame_idx: c_int, pinned_pixels: *const u8, row_pitch: usize, crop_x: c_int, crop_y: c_int) -> c_int;
    pub fn chb_stack_sync(stack: *mut ChbStack) -> c_int;
    pub fn chb_outlier(stack: *mut ChbStack, params: *const ChbOutlierParams, indices: *const i32, n_indices: c_int,
                       out_image: *mut u8, out_mask: *mut u8, n_warnings: *mut u64) -> c_int;
    pub fn chb_outlier_debug(stack: *mut ChbStack, params: *const ChbOutlierParams, indices: *const i32, n_indices: c_int,
                             out_image: *mut u8, out_mask: *mut u8, n_warnings: *mut u64, dbg: *const ChbDebugPlanes) -> c_int;
    pub fn chb_simple(stack: *mut ChbStack, params: *const ChbSimpleParams, indices: *const i32, n_indices: c_int, out_image: *mut u8) -> c_int;
    pub fn chb_outlier_video(stack: *mut ChbStack, params: *const ChbOutlierParams, first_start: c_int, window_len: c_int, n_windows: c_int,
                             out_images: *mut u8, out_masks: *mut u8, n_warnings: *mut u64) -> c_int;
    pub fn chb_shake_create(ctx: *mut ChbCtx, width: c_int, height: c_int, channels: c_int, anchors_xy: *const i32, n_anchors: c_int,
                            anchor_radius: c_int, search_radius: c_int, first_frame: *const u8, row_pitch: usize, out: *mut *mut ChbShake) -> c_int;
    pub fn chb_shake_offset(analyzer: *mut ChbShake, frame: *const u8, row_pitch: usize, out_dx: *mut i32, out_dy: *mut i32, diffs: *mut i32) -> c_int;
    pub fn chb_shake_destroy(analyzer: *mut ChbShake) -> c_int;
    pub fn chb_crop_create(offsets_xy: *const i32, n: c_int, width: c_int, height: c_int, out_xy: *mut i32, out_w: *mut i32, out_h: *mut i32) -> c_int;
    // JPEG at either end of the path (nvJPEG): ImageStream::next -> image::open (src/streams.rs:63-69) and save_image's JPEG
    // branch (src/main.rs:520-571)
    pub fn chb_stack_upload_jpeg(stack: *mut ChbStack, frame_idx: c_int, jpeg: *const u8, n_bytes: usize, crop_x: c_int, crop_y: c_int) -> c_int;
    pub fn chb_decode_jpeg(ctx: *mut ChbCtx, jpeg: *const u8, n_bytes: usize, out: *mut u8, out_cap: usize, row_pitch: usize,
                           out_width: *mut c_int, out_height: *mut c_int) -> c_int;
    pub fn chb_encode_jpeg(ctx: *mut ChbCtx, rgb: *const u8, width: c_int, height: c_int, row_pitch: usize, quality: c_int, out: *mut u8,
                           out_cap: usize, out_size: *mut usize) -> c_int;
}

pub unsafe fn last_error() -> String {
    std::ffi::CStr::from_ptr(chb_last_error()).to_string_lossy().into_owned()
}
pub type Void = c_void;
