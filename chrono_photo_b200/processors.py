"""Host-side mirror of the reference's processor interface for the compositing path, on top of the C ABI.

    FrameStack         replaces TimeSlicer::write_time_slices + the temp slice files (src/slicer.rs:106, src/main.rs:574)
    OutlierProcessor   src/chrono.rs:32-206   new(threshold, bg_mode, outlier_mode, weights, fade, compression, sample_count)
    SimpleProcessor    src/simple.rs:11-168   new(weights, fade, darker)
    Crop / crop_create src/shake.rs:121-181
    video_windows      src/main.rs:230-286 / :349-404

`process` takes the HBM-resident FrameStack where the reference takes the list of slice / image files; every other
argument keeps its meaning. All arithmetic happens in libchrono_b200.so (CUDA); nothing here computes pixels.
"""
import ctypes as C

import numpy as np

from . import _lib
from .options import BackgroundMode, Fade, OutlierSelectionMode, Threshold


def _check_out(arr, shape, name):
    """User-supplied result buffers go straight to the C library: a wrong dtype, shape or stride would make it write out of bounds."""
    if not isinstance(arr, np.ndarray) or arr.dtype != np.uint8 or tuple(arr.shape) != tuple(shape) or not arr.flags["C_CONTIGUOUS"]:
        raise ValueError(f"{name} must be a C-contiguous uint8 array of shape {tuple(shape)}")
    if not arr.flags["WRITEABLE"]:
        raise ValueError(f"{name} must be writeable")
    return arr


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


class Context:
    """chb_ctx: the GPUs this process drives (row shards, one band per device)."""

    def __init__(self, device_ids=None):
        self._h = C.c_void_p()
        if device_ids is None:
            _lib.check(_lib.lib().chb_ctx_create(None, 0, C.byref(self._h)))
        else:
            arr = (C.c_int * len(device_ids))(*device_ids)
            _lib.check(_lib.lib().chb_ctx_create(arr, len(device_ids), C.byref(self._h)))

    @property
    def device_count(self):
        return _lib.lib().chb_ctx_device_count(self._h)

    def set_stream(self, dev_slot, cuda_stream):
        _lib.check(_lib.lib().chb_ctx_set_stream(self._h, dev_slot, C.c_void_p(cuda_stream)))

    def close(self):
        if self._h:
            _lib.lib().chb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FrameStack:
    """Device-resident frame stack (`chb_stack`). width/height are the (cropped) output size."""

    def __init__(self, ctx, width, height, channels, n_frames):
        self.ctx = ctx
        self.width, self.height, self.channels, self.n_frames = int(width), int(height), int(channels), int(n_frames)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().chb_stack_create(ctx._h, self.width, self.height, self.channels, self.n_frames, C.byref(self._h)))

    def upload(self, frame_idx, pixels, crop_xy=(0, 0), pinned=False):
        """pixels: (rows, cols, channels) uint8 host image; crop_xy: Crop origin (src/shake.rs:178-180)."""
        px = pixels
        if px.dtype != np.uint8 or px.ndim != 3 or px.shape[2] != self.channels or not px.flags["C_CONTIGUOUS"]:
            raise ValueError("pixels must be a C-contiguous (rows, cols, channels) uint8 array")
        if px.shape[0] < crop_xy[1] + self.height or px.shape[1] < crop_xy[0] + self.width:
            raise ValueError("crop window outside the host image")
        fn = _lib.lib().chb_stack_upload_pinned if pinned else _lib.lib().chb_stack_upload
        _lib.check(fn(self._h, int(frame_idx), C.c_void_p(px.ctypes.data), px.strides[0], int(crop_xy[0]), int(crop_xy[1])))

    def upload_raw(self, frame_idx, ptr, row_pitch, crop_xy=(0, 0), pinned=True):
        fn = _lib.lib().chb_stack_upload_pinned if pinned else _lib.lib().chb_stack_upload
        _lib.check(fn(self._h, int(frame_idx), C.c_void_p(ptr), int(row_pitch), int(crop_xy[0]), int(crop_xy[1])))

    def upload_jpeg(self, frame_idx, data, crop_xy=(0, 0)):
        """One JPEG-compressed frame (bytes-like): decoded on the GPU into the stack (chb_stack_upload_jpeg)."""
        buf = np.frombuffer(data, dtype=np.uint8)
        _lib.check(_lib.lib().chb_stack_upload_jpeg(self._h, int(frame_idx), C.c_void_p(buf.ctypes.data), buf.size, int(crop_xy[0]), int(crop_xy[1])))

    def upload_all(self, frames, crops=None):
        for i in range(self.n_frames):
            self.upload(i, frames[i], crops[i] if crops is not None else (0, 0))
        self.sync()

    def download(self, frame_idx, out=None):
        """Frame frame_idx as a (height, width, channels) uint8 array (inverse of upload)."""
        if out is None:
            out = np.empty((self.height, self.width, self.channels), dtype=np.uint8)
        else:
            _check_out(out, (self.height, self.width, self.channels), "out")
        _lib.check(_lib.lib().chb_stack_download(self._h, int(frame_idx), C.c_void_p(out.ctypes.data), out.strides[0]))
        return out

    def download_raw(self, frame_idx, ptr, row_pitch):
        _lib.check(_lib.lib().chb_stack_download(self._h, int(frame_idx), C.c_void_p(ptr), int(row_pitch)))

    def sync(self):
        _lib.check(_lib.lib().chb_stack_sync(self._h))

    def wait(self):
        """Waits for enqueued compositing launches; returns (device ms of the last launch, its warning count)."""
        ms, warn = C.c_float(0), C.c_uint64(0)
        _lib.check(_lib.lib().chb_stack_wait(self._h, C.byref(ms), C.byref(warn)))
        return ms.value, warn.value

    def fill_synthetic(self, kind, seed=42, row0_global=0, full_height=None, block_rows=0, block_skip_rows=0):
        """block_rows / block_skip_rows: the stack is an interleaved row-block shard (sharding.InterleavedShard)."""
        _lib.check(_lib.lib().chb_stack_fill_synthetic_blocks(self._h, int(kind), int(seed), int(row0_global),
                                                             int(self.height if full_height is None else full_height),
                                                             int(block_rows), int(block_skip_rows)))

    def device_bytes(self, dev_slot=0):
        return _lib.lib().chb_stack_device_bytes(self._h, dev_slot)

    def close(self):
        if self._h:
            _lib.lib().chb_stack_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synth_frame_host(kind, seed, frame_idx, n_frames, width, full_height, channels, row0=0, rows=None):
    rows = full_height if rows is None else rows
    out = np.empty((rows, width, channels), dtype=np.uint8)
    _lib.check(_lib.lib().chb_synth_frame_host(int(kind), int(seed), int(frame_idx), int(n_frames), int(width), int(full_height),
                                               int(channels), int(row0), int(rows), C.c_void_p(out.ctypes.data)))
    return out


def _indices(image_indices):
    if image_indices is None:
        return None, 0, None
    arr = np.ascontiguousarray(image_indices, dtype=np.int32)
    return arr.ctypes.data_as(C.POINTER(C.c_int32)), len(arr), arr


class OutlierProcessor:
    """OutlierProcessor::new (src/chrono.rs:46-54). `compression` is accepted and ignored (there are no temp files)."""

    def __init__(self, threshold, bg_mode, outlier_mode, weights=(1.0, 1.0, 1.0, 1.0), fade=None, compression=None,
                 sample_count=None, seed=0, pixel_offset=0, block_pixels=0, block_skip=0):
        if not isinstance(threshold, Threshold):
            raise TypeError("threshold must be a Threshold")
        self.threshold = threshold
        self.background = BackgroundMode(bg_mode)
        self.outlier = OutlierSelectionMode(outlier_mode)
        self.weights = [float(w) for w in weights] + [1.0] * (4 - len(weights))
        self.fade = fade if fade is not None else Fade.none()
        self.sample_count = sample_count
        self.seed = int(seed)
        self.pixel_offset = int(pixel_offset)
        self.block_pixels, self.block_skip = int(block_pixels), int(block_skip)  # interleaved row-block shards (chrono_b200.h)
        self.warnings = 0
        self.kernel_ms = None

    def _params(self):
        p = _lib.OutlierParams()
        p.thr_absolute = 1 if self.threshold.absolute else 0
        p.background = int(self.background)
        p.outlier = int(self.outlier)
        p.thr_min, p.thr_max, p.thr_scale = self.threshold.min, self.threshold.max, self.threshold.scale
        for i in range(4):
            p.weights[i] = self.weights[i]
        p.fade = self.fade.to_c()
        p.sample_count = -1 if self.sample_count is None else int(self.sample_count)
        p.seed = self.seed
        p.pixel_offset = self.pixel_offset
        p.block_pixels, p.block_skip = self.block_pixels, self.block_skip
        return p

    def process(self, stack, image_indices=None, want_mask=True, debug=False, out=None, mask_out=None):
        """OutlierProcessor::process (src/chrono.rs:73-206) -> (buffer, is_outlier) as (H, W, C) uint8 arrays.
        debug=True additionally returns a dict of per-pixel sub-results (median, q1, q3, n_outliers).
        out / mask_out: optional preallocated (e.g. pinned) C-contiguous uint8 arrays to receive the results."""
        shape = (stack.height, stack.width, stack.channels)
        out = np.empty(shape, dtype=np.uint8) if out is None else _check_out(out, shape, "out")
        mask = (np.empty(shape, dtype=np.uint8) if mask_out is None else _check_out(mask_out, shape, "mask_out")) if want_mask else None
        ip, n, _keep = _indices(image_indices)
        warn = C.c_uint64(0)
        p = self._params()
        if debug:
            P = stack.height * stack.width
            d = {"median": np.zeros((P, 4), np.float32), "q1": np.zeros((P, 4), np.float32), "q3": np.zeros((P, 4), np.float32),
                 "n_outliers": np.zeros(P, np.int32)}
            planes = _lib.DebugPlanes(d["median"].ctypes.data_as(C.POINTER(C.c_float)), d["q1"].ctypes.data_as(C.POINTER(C.c_float)),
                                      d["q3"].ctypes.data_as(C.POINTER(C.c_float)), d["n_outliers"].ctypes.data_as(C.POINTER(C.c_int32)))
            _lib.check(_lib.lib().chb_outlier_debug(stack._h, C.byref(p), ip, n, C.c_void_p(out.ctypes.data),
                                                    C.c_void_p(mask.ctypes.data) if want_mask else None, C.byref(warn), C.byref(planes)))
            self.warnings = warn.value
            return out, mask, d
        _lib.check(_lib.lib().chb_outlier(stack._h, C.byref(p), ip, n, C.c_void_p(out.ctypes.data),
                                          C.c_void_p(mask.ctypes.data) if want_mask else None, C.byref(warn)))
        self.warnings = warn.value
        return out, mask

    def process_device(self, stack, image_indices=None, want_mask=True):
        """Kernel-only variant: results stay on the device; returns the launch's device time in ms."""
        ip, n, _keep = _indices(image_indices)
        ms = C.c_float(0)
        p = self._params()
        _lib.check(_lib.lib().chb_outlier_device(stack._h, C.byref(p), ip, n, 1 if want_mask else 0, C.byref(ms)))
        self.kernel_ms = ms.value
        return ms.value


    def enqueue_device(self, stack, image_indices=None, want_mask=True):
        """Launch only (no host wait); pair with FrameStack.wait()."""
        ip, n, _keep = _indices(image_indices)
        p = self._params()
        _lib.check(_lib.lib().chb_outlier_enqueue(stack._h, C.byref(p), ip, n, 1 if want_mask else 0))


    # ---- chrono-video: runs of sliding windows (src/main.rs:254-331) ------------------------------------------------
    MAX_VIDEO_WINDOW = 64

    def process_video_run(self, stack, first_start, window_len, n_windows, want_mask=True, out=None, mask_out=None):
        """n_windows windows of window_len consecutive frames, window i = frames [first_start + i, ... + window_len):
        one sliding-window launch sequence (chb_outlier_video). Returns (images, masks, warnings) with images / masks of
        shape (n_windows, H, W, C). out / mask_out: optional preallocated (e.g. pinned) arrays of that shape."""
        shape = (n_windows, stack.height, stack.width, stack.channels)
        out = np.empty(shape, dtype=np.uint8) if out is None else _check_out(out, shape, "out")
        mask = (np.empty(shape, dtype=np.uint8) if mask_out is None else _check_out(mask_out, shape, "mask_out")) if want_mask else None
        warn = (C.c_uint64 * n_windows)()
        p = self._params()
        _lib.check(_lib.lib().chb_outlier_video(stack._h, C.byref(p), int(first_start), int(window_len), int(n_windows), C.c_void_p(out.ctypes.data),
                                                C.c_void_p(mask.ctypes.data) if want_mask else None, warn))
        return out, mask, [int(w) for w in warn]

    def process_video_run_device(self, stack, first_start, window_len, n_windows, want_mask=True):
        """Kernel-only variant of process_video_run: returns the device time of the run in ms."""
        ms = C.c_float(0)
        p = self._params()
        _lib.check(_lib.lib().chb_outlier_video_device(stack._h, C.byref(p), int(first_start), int(window_len), int(n_windows),
                                                       1 if want_mask else 0, C.byref(ms)))
        return ms.value

    def video_runs(self, windows):
        """Splits create_video's window list [(number, indices)] into maximal runs the sliding kernel takes -- consecutive
        windows of equal length <= MAX_VIDEO_WINDOW, each a contiguous frame range starting one frame after the previous
        one -- and single windows. Returns [(first_position_in_windows, count)], count > 1 only for runs."""
        def slidable(idx):
            return (1 <= len(idx) <= self.MAX_VIDEO_WINDOW and idx[-1] - idx[0] + 1 == len(idx)
                    and (self.sample_count is None or self.sample_count >= len(idx)))
        runs, i = [], 0
        while i < len(windows):
            idx = windows[i][1]
            j = i + 1
            if slidable(idx):
                while (j < len(windows) and len(windows[j][1]) == len(idx) and slidable(windows[j][1])
                       and windows[j][1][0] == windows[j - 1][1][0] + 1):
                    j += 1
            runs.append((i, j - i))
            i = j
        return runs

    def process_video(self, stack, windows, want_mask=True):
        """Composites every window of `windows` (as returned by video_windows): runs through chb_outlier_video, the
        remaining windows one by one through chb_outlier. Yields (number, image, mask, warnings) in window order."""
        for pos, count in self.video_runs(windows):
            if count > 1:
                idx0 = windows[pos][1]
                imgs, masks, warns = self.process_video_run(stack, idx0[0], len(idx0), count, want_mask)
                for k in range(count):
                    yield windows[pos + k][0], imgs[k], (masks[k] if want_mask else None), warns[k]
            else:
                img, mask = self.process(stack, windows[pos][1], want_mask=want_mask)
                yield windows[pos][0], img, mask, self.warnings


class SimpleProcessor:
    """SimpleProcessor::new(weights, fade, darker) (src/simple.rs:18)."""

    def __init__(self, weights=(1.0, 1.0, 1.0, 1.0), fade=None, darker=True):
        self.weights = [float(w) for w in weights] + [1.0] * (4 - len(weights))
        self.fade = fade if fade is not None else Fade.none()
        self.darker = bool(darker)
        self.kernel_ms = None

    def _params(self):
        p = _lib.SimpleParams()
        p.darker = 1 if self.darker else 0
        for i in range(4):
            p.weights[i] = self.weights[i]
        p.fade = self.fade.to_c()
        return p

    def process(self, stack, image_indices=None):
        """SimpleProcessor::process (src/simple.rs:26-168) -> (H, W, C) uint8 buffer."""
        out = np.empty((stack.height, stack.width, stack.channels), dtype=np.uint8)
        ip, n, _keep = _indices(image_indices)
        p = self._params()
        _lib.check(_lib.lib().chb_simple(stack._h, C.byref(p), ip, n, C.c_void_p(out.ctypes.data)))
        return out

    def process_device(self, stack, image_indices=None):
        ip, n, _keep = _indices(image_indices)
        ms = C.c_float(0)
        p = self._params()
        _lib.check(_lib.lib().chb_simple_device(stack._h, C.byref(p), ip, n, C.byref(ms)))
        self.kernel_ms = ms.value
        return ms.value


def encode_jpeg(ctx, image, quality=95):
    """save_image's JPEG branch (src/main.rs:520-571) on the GPU: (H, W, 3) uint8 -> bytes."""
    img = np.ascontiguousarray(image, dtype=np.uint8)
    if img.ndim != 3 or img.shape[2] != 3:
        raise ValueError("image must be (H, W, 3) uint8")
    h, w, _ = img.shape
    cap = w * h * 3 + 65536
    out = np.empty(cap, np.uint8)
    size = C.c_size_t(0)
    _lib.check(_lib.lib().chb_encode_jpeg(ctx._h, C.c_void_p(img.ctypes.data), w, h, img.strides[0], int(quality), C.c_void_p(out.ctypes.data), cap, C.byref(size)))
    return out[:size.value].tobytes()


def decode_jpeg(ctx, data):
    """JPEG bytes -> (H, W, 3) uint8 host image, decoded on the GPU (chb_decode_jpeg): the host copy the shake analysis reads
    (image::open per frame, src/shake.rs:248-283)."""
    buf = np.frombuffer(data, dtype=np.uint8)
    w, h = C.c_int(0), C.c_int(0)
    _lib.check(_lib.lib().chb_decode_jpeg(ctx._h, C.c_void_p(buf.ctypes.data), buf.size, None, 0, 0, C.byref(w), C.byref(h)))
    out = np.empty((h.value, w.value, 3), dtype=np.uint8)
    _lib.check(_lib.lib().chb_decode_jpeg(ctx._h, C.c_void_p(buf.ctypes.data), buf.size, C.c_void_p(out.ctypes.data), out.nbytes, w.value * 3,
                                          C.byref(w), C.byref(h)))
    return out


def set_tuning(key, value):
    """Tuning / test knobs of the library (chb_set_tuning): 'force_variant', 'hist', 'pdl', 'video_queue_cap', 'inline_min'."""
    _lib.check(_lib.lib().chb_set_tuning(str(key).encode(), int(value)))


def fetch_last_device(stack, d_image_ptr, d_mask_ptr=None, dev_slot=0):
    """The last call's band of device slot dev_slot copied into caller-owned device buffers (raw device pointers), on the
    context's compute stream -- the input of a band gather over NVLink (chb_fetch_last_device)."""
    _lib.check(_lib.lib().chb_fetch_last_device(stack._h, int(dev_slot), C.c_void_p(int(d_image_ptr)),
                                                C.c_void_p(int(d_mask_ptr)) if d_mask_ptr else None))


def fetch_last(stack, want_mask=False):
    out = np.empty((stack.height, stack.width, stack.channels), dtype=np.uint8)
    mask = np.empty_like(out) if want_mask else None
    warn = C.c_uint64(0)
    _lib.check(_lib.lib().chb_fetch_last(stack._h, C.c_void_p(out.ctypes.data), C.c_void_p(mask.ctypes.data) if want_mask else None, C.byref(warn)))
    return out, mask, warn.value


class ShakeAnalyzer:
    """ShakeAnalyzer::analyze (src/shake.rs:190-305) over decoded frames: the first frame supplies the anchor windows, every
    further frame gets its (dx, dy) from the GPU (sums of squared differences over the search square, first minimum)."""

    def __init__(self, ctx, first_frame, anchors, anchor_radius, search_radius):
        f = np.ascontiguousarray(first_frame, dtype=np.uint8)
        if f.ndim != 3:
            raise ValueError("frame must be (H, W, C) uint8")
        self.shape = f.shape
        anc = np.ascontiguousarray([a.anchor if hasattr(a, "anchor") else a for a in anchors], dtype=np.int32).reshape(-1, 2)
        self.search_size = 2 * int(search_radius) + 1
        h = C.c_void_p()
        _lib.check(_lib.lib().chb_shake_create(ctx._h, f.shape[1], f.shape[0], f.shape[2], anc.ctypes.data_as(C.POINTER(C.c_int32)), len(anc),
                                               int(anchor_radius), int(search_radius), C.c_void_p(f.ctypes.data), f.shape[1] * f.shape[2], C.byref(h)))
        self._h = h

    def offset(self, frame, want_diffs=False):
        """One frame -> (dx, dy) [, diff table of shape (2s+1, 2s+1), rows = dy]."""
        f = np.ascontiguousarray(frame, dtype=np.uint8)
        if f.shape != self.shape:
            raise ValueError("Image layout does not fit!")
        dx, dy = C.c_int32(), C.c_int32()
        diffs = np.zeros((self.search_size, self.search_size), np.int32) if want_diffs else None
        _lib.check(_lib.lib().chb_shake_offset(self._h, C.c_void_p(f.ctypes.data), f.shape[1] * f.shape[2], C.byref(dx), C.byref(dy),
                                               diffs.ctypes.data_as(C.POINTER(C.c_int32)) if want_diffs else None))
        return ((dx.value, dy.value), diffs) if want_diffs else (dx.value, dy.value)

    @classmethod
    def analyze(cls, ctx, frames, anchors, anchor_radius, search_radius):
        """-> [(0, 0), (dx1, dy1), ...] for an iterable of (H, W, C) frames, like the reference's return value."""
        it = iter(frames)
        first = next(it)
        an = cls(ctx, first, anchors, anchor_radius, search_radius)
        try:
            return [(0, 0)] + [an.offset(f) for f in it]
        finally:
            an.close()

    def close(self):
        if self._h:
            _lib.lib().chb_shake_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def crop_create(offsets, width, height):
    """Crop::create (src/shake.rs:136-176): per-frame crop origins and the common size, or None if all offsets are zero."""
    off = np.ascontiguousarray(offsets, dtype=np.int32).reshape(-1, 2)
    xy = np.zeros_like(off)
    w, h = C.c_int32(), C.c_int32()
    r = _lib.lib().chb_crop_create(off.ctypes.data_as(C.POINTER(C.c_int32)), len(off), int(width), int(height),
                                   xy.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(w), C.byref(h))
    if r == 0:
        return None
    return xy, w.value, h.value


def video_windows(image_count, video_in, video_out):
    """Window index lists of create_video / create_video_simple (src/main.rs:230-286): returns [(number, indices)],
    skipping the frames the reference skips (empty windows)."""
    cap = max(1, 4 * image_count + 16)
    ws, we, num = (np.zeros(cap, np.int32) for _ in range(3))
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    n = _lib.lib().chb_video_windows(int(image_count),
                                     video_in.start is not None, video_in.start or 0, video_in.end is not None, video_in.end or 0, video_in.step,
                                     video_out.start is not None, video_out.start or 0, video_out.end is not None, video_out.end or 0, video_out.step,
                                     p(ws), p(we), p(num), cap)
    if n < 0:
        raise ValueError("invalid frame ranges")
    out = []
    for i in range(min(n, cap)):
        idx = list(range(int(ws[i]), int(we[i]), video_in.step))
        if idx:
            out.append((int(num[i]), idx))
    return out


def sample_positions(seed, n, cnt):
    out = np.zeros(cnt, np.int32)
    _lib.check(_lib.lib().chb_sample_positions(int(seed), int(n), int(cnt), out.ctypes.data_as(C.POINTER(C.c_int32))))
    return out
