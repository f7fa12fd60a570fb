//! RAII wrapper of chb_ctx + chb_stack: what replaces the temp-file time slices (TimeSlicer::write_time_slices,
//! src/slicer.rs:106-231) in the reference. Written against include/chrono_b200.h; not compiled in this repository's image.
use crate::ffi;
use std::io;

fn check(rc: i32) -> io::Result<()> {
    if rc == ffi::CHB_OK { Ok(()) } else { Err(io::Error::new(io::ErrorKind::Other, unsafe { ffi::last_error() })) }
}

pub struct GpuContext { raw: *mut ffi::ChbCtx }
impl GpuContext {
    /// `devices`: CUDA device ordinals; every stack of the context is row-sharded over them.
    pub fn new(devices: &[i32]) -> io::Result<Self> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::chb_ctx_create(if devices.is_empty() { std::ptr::null() } else { devices.as_ptr() }, devices.len() as i32, &mut raw) })?;
        Ok(GpuContext { raw })
    }
    pub fn raw(&self) -> *mut ffi::ChbCtx { self.raw }
}
impl Drop for GpuContext { fn drop(&mut self) { unsafe { ffi::chb_ctx_destroy(self.raw); } } }
unsafe impl Send for GpuContext {}
unsafe impl Sync for GpuContext {}

pub struct GpuStack { raw: *mut ffi::ChbStack, pub width: u32, pub height: u32, pub channels: u32, pub frames: usize }
impl GpuStack {
    pub fn new(ctx: &GpuContext, width: u32, height: u32, channels: u32, frames: usize) -> io::Result<Self> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::chb_stack_create(ctx.raw(), width as i32, height as i32, channels as i32, frames as i32, &mut raw) })?;
        Ok(GpuStack { raw, width, height, channels, frames })
    }
    /// One decoded frame (`samples` = `as_flat_samples_u8().samples`, `row_pitch` = `layout.height_stride`), cropped at
    /// (crop_x, crop_y) = the frame's `Crop` origin (src/shake.rs:136-176). Asynchronous; callable from the decode threads.
    pub fn upload(&self, frame: usize, samples: &[u8], row_pitch: usize, crop_x: u32, crop_y: u32) -> io::Result<()> {
        let need = (crop_y as usize + self.height as usize - 1) * row_pitch + (crop_x as usize + self.width as usize) * self.channels as usize;
        if samples.len() < need { return Err(io::Error::new(io::ErrorKind::InvalidInput, "frame smaller than the cropped stack")); }
        check(unsafe { ffi::chb_stack_upload(self.raw, frame as i32, samples.as_ptr(), row_pitch, crop_x as i32, crop_y as i32) })
    }
    pub fn sync(&self) -> io::Result<()> { check(unsafe { ffi::chb_stack_sync(self.raw) }) }
    pub fn raw(&self) -> *mut ffi::ChbStack { self.raw }
}
impl Drop for GpuStack { fn drop(&mut self) { unsafe { ffi::chb_stack_destroy(self.raw); } } }
// chb_outlier / chb_simple may be entered from the rayon pool of create_video (src/main.rs:260-261, :378-379)
unsafe impl Send for GpuStack {}
unsafe impl Sync for GpuStack {}
