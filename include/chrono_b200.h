/*
 * chrono_b200.h -- C ABI of libchrono_b200.so: the B200 (sm_100a) implementation of chrono-photo's
 * frame-stack compositing path (`--mode outlier | darker | lighter`).
 *
 * There is no FFI/plugin interface in the reference; the seam this library replaces is two Rust methods
 * (paths relative to mlange-42/chrono-photo v0.6.5):
 *   OutlierProcessor::new / ::process   src/chrono.rs:46-54, :73-81   called by create_frame        src/main.rs:454-492
 *   SimpleProcessor::new  / ::process   src/simple.rs:18,   :26-32    called by create_frame_simple src/main.rs:495-517
 * plus the storage they sit on, which disappears:
 *   TimeSlicer::write_time_slices       src/slicer.rs:106-231  -> chb_stack_create + chb_stack_upload (HBM-resident stack)
 *   PixelInputStream::read_chunk        src/streams.rs:169-202 -> the kernels read the stack directly
 *
 * Conventions: plain pointers and sizes only; every function returns an int status (CHB_OK == 0) and never
 * throws or aborts across the boundary; chb_last_error() returns a thread-local message for the last
 * failing call on the calling thread. The caller owns every host buffer; the library owns device memory.
 * There is no CPU fallback: without a CUDA device chb_ctx_create fails with CHB_ERR_CUDA.
 * chb_outlier / chb_simple may be entered concurrently from several host threads on one stack (the
 * reference's video path calls the processors from a rayon pool, src/main.rs:260-261, :378-379): a stack
 * owns four call slots (stream, output planes, tables, queues), so up to four such calls overlap their
 * launches, tier kernels and D2H copies; further callers wait for a slot. The device-side entry points
 * (chb_*_device, chb_outlier_enqueue, chb_stack_wait, chb_fetch_last*, chb_outlier_video*) are defined on
 * "the last call" of the stack and serialise on slot 0. Compositing launches order themselves after the
 * uploads issued before them (no chb_stack_sync needed in between).
 */
#ifndef CHRONO_B200_H
#define CHRONO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHB_VERSION 200 /* 0.2.0 */

enum chb_status {
    CHB_OK = 0,
    CHB_ERR_INVALID = 1,     /* bad argument (the reference would panic or index out of range) */
    CHB_ERR_CUDA = 2,        /* CUDA runtime failure, no device, or out of device memory */
    CHB_ERR_UNSUPPORTED = 3, /* valid in the reference but outside what this build handles (see message) */
    CHB_ERR_STATE = 4        /* call order (e.g. compositing before every frame of the window was uploaded) */
};

/* BackgroundMode, src/options.rs:319-328 */
enum chb_background { CHB_BG_FIRST = 0, CHB_BG_RANDOM = 1, CHB_BG_AVERAGE = 2, CHB_BG_MEDIAN = 3 };
/* OutlierSelectionMode, src/options.rs:284-297 */
enum chb_outlier_mode {
    CHB_OUT_FIRST = 0, CHB_OUT_LAST = 1, CHB_OUT_EXTREME = 2, CHB_OUT_AVERAGE = 3, CHB_OUT_FORWARD = 4, CHB_OUT_BACKWARD = 5
};
/* FadeMode, src/options.rs:35-41 */
enum chb_fade_mode { CHB_FADE_CLAMP = 0, CHB_FADE_REPEAT = 1 };

/* Fade, src/options.rs:59-66: the already-built LUT (Fade::new, :69-94) is passed as is. */
typedef struct chb_fade {
    uint8_t is_none;  /* Fade::none(): every lookup returns 1.0 */
    uint8_t mode;     /* chb_fade_mode */
    uint8_t absolute; /* 1: lookup frame_offset + position, 0: lookup total - position - 1 (src/chrono.rs:496-502) */
    uint8_t _pad;
    int32_t offset;   /* first frame of the LUT */
    int32_t n_values; /* <= CHB_MAX_FADE_VALUES */
    const float *values;
} chb_fade;
#define CHB_MAX_FADE_VALUES 2048

/* Arguments of OutlierProcessor::new (src/chrono.rs:46-54). Threshold fields are Threshold's internal units
 * (src/options.rs:197-213: abs thresholds already multiplied by 255, scale precomputed in f32). */
typedef struct chb_outlier_params {
    uint8_t thr_absolute;
    uint8_t background; /* chb_background */
    uint8_t outlier;    /* chb_outlier_mode */
    uint8_t _pad;
    float thr_min, thr_max, thr_scale;
    float weights[4];
    chb_fade fade;
    int32_t sample_count; /* --sample: median/IQR on this many randomly chosen frames of the window; <0 = all */
    uint64_t seed;        /* counter-based RNG seed for --background random and --sample (the reference uses thread_rng) */
    uint64_t pixel_offset; /* global index of this stack's pixel 0; lets row shards in separate processes draw the same numbers */
    /* Interleaved row-block shards (GPU g of G owns the blocks of B rows with index = g mod G, so that objects spread evenly over
     * the GPUs): the stack then holds blocks of block_pixels = B * image_width pixels that lie block_pixels + block_skip apart
     * in the whole image (block_skip = (G - 1) * block_pixels) and pixel_offset = g * block_pixels. The global index of local
     * pixel p is pixel_offset + p + (p / block_pixels) * block_skip. 0 / 0 = one contiguous band (single-device stacks only). */
    uint64_t block_pixels;
    uint64_t block_skip;
} chb_outlier_params;

/* Arguments of SimpleProcessor::new (src/simple.rs:18). */
typedef struct chb_simple_params {
    uint8_t darker; /* 1 = --mode darker, 0 = --mode lighter */
    uint8_t _pad[3];
    float weights[4];
    chb_fade fade;
} chb_simple_params;

/* Optional per-pixel sub-results (device -> host) for bit-exact parity checks; any pointer may be NULL. */
typedef struct chb_debug_planes {
    float *median;       /* [H*W][4] */
    float *q1;           /* [H*W][4]  (rel thresholds) */
    float *q3;           /* [H*W][4] */
    int32_t *n_outliers; /* [H*W] */
} chb_debug_planes;

typedef struct chb_ctx chb_ctx;
typedef struct chb_stack chb_stack;

const char *chb_last_error(void);
int chb_version(void);

/* One context spans the GPUs a process drives. device_ids == NULL, n_dev == 0: device 0 only.
 * With n_dev > 1 every stack is row-sharded over the devices (no collective; bands are gathered by D2H copies). */
int chb_ctx_create(const int *device_ids, int n_dev, chb_ctx **out);
int chb_ctx_destroy(chb_ctx *ctx);
int chb_ctx_device_count(const chb_ctx *ctx);
/* Free / total HBM of device slot dev_slot: what a caller sizes its row bands by when a frame series does not fit at once (the
 * reference is out-of-core by construction, SliceLength::bytes, src/slicer.rs:19-41; here the caller composites the image in
 * row bands, each band a stack of its own: chb_stack_upload's crop origin selects the band's rows, pixel_offset keeps the
 * per-pixel RNG aligned). */
int chb_ctx_mem_info(chb_ctx *ctx, int dev_slot, size_t *free_bytes, size_t *total_bytes);
/* Use the caller's CUDA stream (cudaStream_t as void*) for device `dev_slot`'s compute work, so the caller can
 * bracket launches with its own events (torch.cuda.Event only sees torch's current stream). NULL = internal stream. */
int chb_ctx_set_stream(chb_ctx *ctx, int dev_slot, void *cuda_stream);

/* Replaces TimeSlicer::write_time_slices (src/slicer.rs:106): allocates the HBM-resident stack for n_frames
 * frames of width x height x channels (channels = SampleLayout.width_stride: 3 = RGB8 or 4 = RGBA8, the two layouts
 * save_image knows, src/main.rs:550-567), row-sharded over the context's devices. */
int chb_stack_create(chb_ctx *ctx, int width, int height, int channels, int n_frames, chb_stack **out);
int chb_stack_destroy(chb_stack *stack);
/* bytes of HBM held by the stack on device slot dev_slot */
size_t chb_stack_device_bytes(const chb_stack *stack, int dev_slot);

/* Upload one decoded frame (interleaved u8, row pitch in bytes). crop_x/crop_y: top-left corner of the
 * width x height window inside the host image (Crop::crop, src/shake.rs:178-180; slicer.rs:138-140). The copy is
 * asynchronous (pinned staging, double-buffered cudaMemcpyAsync + a device re-layout kernel); callable from several
 * threads for distinct frame_idx. The host buffer may be reused once the call returns. */
int chb_stack_upload(chb_stack *stack, int frame_idx, const uint8_t *host_pixels, size_t row_pitch, int crop_x, int crop_y);
/* Same, from a pinned (page-locked) host buffer the caller keeps alive until chb_stack_sync: no staging copy. */
int chb_stack_upload_pinned(chb_stack *stack, int frame_idx, const uint8_t *pinned_pixels, size_t row_pitch, int crop_x, int crop_y);
/* Read frame frame_idx back from the stack (interleaved u8, tightly packed rows of width*channels bytes at row_pitch):
 * the inverse of chb_stack_upload, for checks and for filling host buffers from a device-generated series. */
int chb_stack_download(chb_stack *stack, int frame_idx, uint8_t *host_pixels, size_t row_pitch);
/* Wait until every upload issued so far has landed in HBM. */
int chb_stack_sync(chb_stack *stack);

/* Fill the stack on the device with the synthetic series S<kind> of DESIGN.md (counter-based generator,
 * identical bytes to chb_synth_frame_host). row0_global: image row of this stack's first row (for shards that
 * live in separate processes); full_height: height of the whole image. */
int chb_stack_fill_synthetic(chb_stack *stack, int kind, uint64_t seed, int row0_global, int full_height);
/* Same for an interleaved row-block shard (chb_outlier_params.block_pixels; single-device stacks): local row r of the stack
 * is row row0_global + r + (r / block_rows) * block_skip_rows of the whole image. 0 / 0 = one contiguous band. */
int chb_stack_fill_synthetic_blocks(chb_stack *stack, int kind, uint64_t seed, int row0_global, int full_height, int block_rows,
                                    int block_skip_rows);
/* Host twin of the generator: writes frame `frame_idx` rows [row0, row0+rows) of a width x full_height image. */
int chb_synth_frame_host(int kind, uint64_t seed, int frame_idx, int n_frames, int width, int full_height, int channels,
                         int row0, int rows, uint8_t *out_pixels);

/* OutlierProcessor::process (src/chrono.rs:73-206). indices: ascending frame indices of the window
 * (image_indices; NULL = every frame); frame_offset is taken as indices[0] like the reference (chrono.rs:103).
 * out_image / out_mask: height*width*channels bytes each, tightly packed; out_mask may be NULL (--output-blend
 * absent). n_warnings: pixels that consist of only outliers (chrono.rs:198-203), may be NULL. */
int chb_outlier(chb_stack *stack, const chb_outlier_params *params, const int32_t *indices, int n_indices,
                uint8_t *out_image, uint8_t *out_mask, uint64_t *n_warnings);
/* Same plus per-pixel sub-results. */
int chb_outlier_debug(chb_stack *stack, const chb_outlier_params *params, const int32_t *indices, int n_indices,
                      uint8_t *out_image, uint8_t *out_mask, uint64_t *n_warnings, const chb_debug_planes *dbg);

/* SimpleProcessor::process (src/simple.rs:26-168). */
int chb_simple(chb_stack *stack, const chb_simple_params *params, const int32_t *indices, int n_indices, uint8_t *out_image);

/* Device-resident variants used for kernel-only timing: results stay in the library's device buffers, nothing is
 * copied to the host. kernel_ms (may be NULL) receives the device time of the launches, measured with CUDA events
 * on the launching stream, max over the context's devices. chb_fetch_last copies the results of the stack's last
 * *_device call to the host (threads that share a stack must use chb_outlier / chb_simple, which launch and fetch in one
 * critical section). */
int chb_outlier_device(chb_stack *stack, const chb_outlier_params *params, const int32_t *indices, int n_indices,
                       int want_mask, float *kernel_ms);
int chb_simple_device(chb_stack *stack, const chb_simple_params *params, const int32_t *indices, int n_indices, float *kernel_ms);
int chb_fetch_last(chb_stack *stack, uint8_t *out_image, uint8_t *out_mask, uint64_t *n_warnings);
/* The band of device slot dev_slot of the stack's last compositing call, copied into caller-owned DEVICE buffers on that
 * GPU (rows * W * C bytes each; d_mask nullable), asynchronously on the context's compute stream: the input of the final
 * band gather when one rank per GPU holds a band and the image is assembled over NVLink. */
int chb_fetch_last_device(chb_stack *stack, int dev_slot, void *d_image, void *d_mask);
/* Back-to-back launches without a host round trip per call: chb_outlier_enqueue only launches (it waits by itself when
 * the window / sample / fade tables differ from the previous call's), chb_stack_wait waits for everything enqueued and
 * returns the device time of the last launch and its warning count (an enqueued call leaves its counters on the device;
 * chb_stack_wait and chb_fetch_last fetch them). */
int chb_outlier_enqueue(chb_stack *stack, const chb_outlier_params *params, const int32_t *indices, int n_indices, int want_mask);
int chb_stack_wait(chb_stack *stack, float *last_kernel_ms, uint64_t *n_warnings);

/* ---- chrono-video: a run of sliding windows in one call --------------------------------------------------------
 * Replaces the per-output-frame loop of create_video (src/main.rs:254-331: one OutlierProcessor::process per frame of the
 * video, each re-reading its window) for a run of n_windows windows of window_len consecutive frames, window i =
 * frames [first_start + i, first_start + i + window_len) -- what `--video-in a/b/1` produces once the window has its
 * full length (src/main.rs:262-286); frame_offset of window i is first_start + i (src/chrono.rs:102-103). Every frame
 * group is loaded once per 16 windows and the window slides inside registers. out_images / out_masks (nullable) are
 * [n_windows][H*W*C] planes; n_warnings (nullable) receives n_windows per-window counts. Limits: window_len <= 64 and no
 * --sample below the window length (CHB_ERR_UNSUPPORTED: composite such windows one by one with chb_outlier). Results
 * are identical to n_windows chb_outlier calls. */
int chb_outlier_video(chb_stack *stack, const chb_outlier_params *params, int first_start, int window_len, int n_windows,
                      uint8_t *out_images, uint8_t *out_masks, uint64_t *n_warnings);
/* Same, results stay on the device(s); kernel_ms = device time of all launches of the run (max over devices). */
int chb_outlier_video_device(chb_stack *stack, const chb_outlier_params *params, int first_start, int window_len,
                             int n_windows, int want_mask, float *kernel_ms);

/* Counters for bench.py: kernels launched by this library since the last reset (process-wide),
 * and, for the last outlier call on this thread, how many pixels left the certified fast path. */
uint64_t chb_launch_count(void);
void chb_launch_count_reset(void);
uint64_t chb_last_slow_pixels(void);
uint64_t chb_last_hard_pixels(void);
/* Device time (ms, max over devices) of the streaming kernel alone in the calling thread's last outlier call; the call's
 * kernel_ms also covers the two tier kernels that follow it. */
float chb_last_main_kernel_ms(void); /* pixels whose medians needed the iterative solver */

/* ---- JPEG at either end of the path, on the GPU (nvJPEG) -------------------------------------------------------------
 * chb_stack_upload_jpeg replaces ImageStream::next -> image::open (src/streams.rs:63-69) + the time-slice write for one frame:
 * the COMPRESSED bytes cross PCIe, the frame is decoded on the device into interleaved RGB8 and re-laid-out like
 * chb_stack_upload (crop origin = the frame's Crop, src/shake.rs:136-176). Callable from several decode threads (nvJPEG's
 * Huffman stage runs in the caller). RGB stacks only. chb_encode_jpeg is save_image's JPEG branch (src/main.rs:520-571,
 * quality 1..100): *out_size receives the stream length; with out == NULL or a too small buffer it fails with
 * CHB_ERR_INVALID after setting *out_size to the needed size. Decoders differ in the last bit: parity is defined on decoded
 * frames. */
int chb_stack_upload_jpeg(chb_stack *stack, int frame_idx, const uint8_t *jpeg, size_t n_bytes, int crop_x, int crop_y);
int chb_encode_jpeg(chb_ctx *ctx, const uint8_t *rgb, int width, int height, size_t row_pitch, int quality, uint8_t *out,
                    size_t out_cap, size_t *out_size);
/* JPEG stream -> interleaved RGB8 HOST image (decoded on the device, copied back): the host copy the callers around the path
 * need, e.g. the shake analysis that reads every frame before ingest (image::open in src/shake.rs:248-283). out == NULL:
 * only *out_width / *out_height are set. */
int chb_decode_jpeg(chb_ctx *ctx, const uint8_t *jpeg, size_t n_bytes, uint8_t *out, size_t out_cap, size_t row_pitch,
                    int *out_width, int *out_height);

/* Tuning / test knobs (not needed for normal use; initial values come from the environment variables CHB_<KEY> read once at
 * load time): "force_variant", "hist", "pdl", "video_queue_cap", "inline_min", "hard_inline_min"; value -1 = automatic. */
int chb_set_tuning(const char *key, int value);

/* The --sample subset the library draws for (seed, window length n, cnt): cnt ascending positions in [0, n).
 * Deterministic replacement of rand::seq::sample_indices (src/chrono.rs:157), exported so a checker can use the same set. */
int chb_sample_positions(uint64_t seed, int n, int cnt, int32_t *out_positions);

/* Host arithmetic on the path, exported so the bindings need not re-implement it:
 * Threshold::new (src/options.rs:197-213). */
void chb_threshold_new(int absolute, float min, float max, float *out_min, float *out_max, float *out_scale);
/* Fade::new (src/options.rs:69-94): builds the LUT from (frame, value) pairs; returns the number of values or <0. */
int chb_fade_build(const int32_t *frames, const float *values, int n_pairs, float *out_values, int out_cap, int32_t *out_offset);
/* ---- camera-shake analysis (ShakeAnalyzer::analyze, src/shake.rs:190-305) ------------------------------------
 * chb_shake_create takes the FIRST frame and keeps the (2r+1)^2 x C window around every anchor (fill_windows,
 * src/shake.rs:307-336). chb_shake_offset replaces one iteration of the par_iter over the remaining files
 * (src/shake.rs:248-283): sums of squared differences over the (2s+1)^2 search square for all anchors (calc_diffs,
 * :338-386, i32 with wrap-around like the release build) and the FIRST minimum -> (dx, dy). Only the patches around the
 * anchors cross PCIe. diffs (nullable) receives the table, row-major (oy, ox). Where the reference panics with "Image
 * coordinate out of range" the call fails with CHB_ERR_INVALID and that message. Calls on one analyzer are serialised
 * (callable from the rayon pool). The offsets feed chb_crop_create; frame 0's offset is (0, 0) by definition. */
typedef struct chb_shake chb_shake;
int chb_shake_create(chb_ctx *ctx, int width, int height, int channels, const int32_t *anchors_xy, int n_anchors,
                     int anchor_radius, int search_radius, const uint8_t *first_frame, size_t row_pitch, chb_shake **out);
int chb_shake_offset(chb_shake *analyzer, const uint8_t *frame, size_t row_pitch, int32_t *out_dx, int32_t *out_dy,
                     int32_t *diffs);
int chb_shake_destroy(chb_shake *analyzer);

/* Crop::create (src/shake.rs:136-176). Returns 1 and fills out_xy (n x (x,y)), out_w, out_h; 0 if all offsets are zero. */
int chb_crop_create(const int32_t *offsets_xy, int n, int width, int height, int32_t *out_xy, int32_t *out_w, int32_t *out_h);
/* Window index math of create_video / create_video_simple (src/main.rs:230-286, :349-404). Returns the number of
 * output frames; writes at most cap windows [start, end) (step = in_step; empty = skipped frame) and their numbers. */
int chb_video_windows(int image_count, int in_has_start, int in_start, int in_has_end, int in_end, int in_step,
                      int out_has_start, int out_start, int out_has_end, int out_end, int out_step,
                      int32_t *win_start, int32_t *win_end, int32_t *out_number, int cap);

#ifdef __cplusplus
}
#endif
#endif
