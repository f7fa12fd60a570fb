/*
 * chrono_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, no SIMD, no FMA contraction) of the frame-stack compositing path of
 * mlange-42/chrono-photo v0.6.5. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library; the product (chrono_photo_b200/, libchrono_b200.so)
 * never links, imports or calls it.
 *
 * Parity pin: the reference cannot be compiled here (no cargo/rustc), and its own tests hold exactly
 * one known-answer vector for this path -- quartiles([0..6]) == (1,3,5) (src/chrono.rs:598-604) -- plus
 * the commented-out blend expectations 0 -> 0, 0.5 -> 128, 1 -> 255 (src/color.rs:55-63). Both are checked
 * in tests/test_oracle.py. Everything beyond those vectors is "parity unpinned": the pin is this
 * line-by-line restatement, cross-checked by an independent numpy restatement (tests/np_restatement.py).
 *
 * The RNG of the reference (`rand::thread_rng`, OS seeded: src/chrono.rs:68,157,357,542,553) is not
 * reproducible even by the reference itself; this oracle and the CUDA path share a counter-based RNG
 * (orc_rng_range) instead, so `--background random` is bit-comparable between the two but only
 * distribution-comparable with the reference.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 */
#ifndef CHRONO_ORACLE_H
#define CHRONO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/options.rs:188-194 (Threshold, already in internal units: abs thresholds are scaled by 255). */
typedef struct {
    int32_t absolute;
    float min, max, scale;
} orc_threshold;

/* src/options.rs:59-66 (Fade). mode: 0 = clamp, 1 = repeat. */
typedef struct {
    int32_t is_none;
    int32_t mode;
    int32_t absolute;
    int32_t offset;
    int32_t n_values;
    const float *values;
} orc_fade;

enum { ORC_BG_FIRST = 0, ORC_BG_RANDOM = 1, ORC_BG_AVERAGE = 2, ORC_BG_MEDIAN = 3 };
enum { ORC_OUT_FIRST = 0, ORC_OUT_LAST = 1, ORC_OUT_EXTREME = 2, ORC_OUT_AVERAGE = 3, ORC_OUT_FORWARD = 4, ORC_OUT_BACKWARD = 5 };

/* Arguments of OutlierProcessor::new (src/chrono.rs:46-54) that reach the arithmetic. */
typedef struct {
    orc_threshold threshold;
    int32_t background;
    int32_t outlier;
    float weights[4];
    orc_fade fade;
    uint64_t seed;         /* counter-based RNG seed (replaces thread_rng) */
    uint64_t pixel_offset; /* global index of pixel 0 of this image/band, so row shards draw the same numbers */
} orc_outlier_params;

/* Optional per-pixel sub-results for bit-exact checks of medians, quartiles, outlier counts and indices. */
typedef struct {
    float *median;       /* [P][4] */
    float *q1;           /* [P][4] (rel thresholds only, else 0) */
    float *q3;           /* [P][4] */
    int32_t *n_outliers; /* [P] */
    int32_t *sel_index;  /* [P] window position used as the outlier (-1: none / blended list / average) */
    int32_t *bg_index;   /* [P] window position used as background (-1 for average/median) */
} orc_debug;

/* ---- option arithmetic ---- */
void orc_threshold_new(int absolute, float min, float max, orc_threshold *out);   /* options.rs:197-213 */
float orc_threshold_blend_value(const orc_threshold *t, float dist);              /* options.rs:223-231 */
/* options.rs:69-94. frames/values: n_pairs (frame, value) pairs ordered by frame. Writes last-first+1 values. Returns count. */
int orc_fade_build(const int32_t *frames, const float *vals, int n_pairs, float *out_values, int out_cap, int32_t *out_offset);
float orc_fade_get(const orc_fade *f, int32_t frame);                              /* options.rs:113-139 */

/* ---- order statistics on a sorted sample (src/chrono.rs:559-591) ---- */
float orc_median(const uint8_t *sorted, size_t len);
float orc_quantile(const uint8_t *sorted, size_t len, float q);
void orc_quartiles(const uint8_t *sorted, size_t len, float *q1, float *med, float *q3);

/* ---- colour blending (src/color.rs:4-44) ---- */
void orc_blend_into_u8(uint8_t *a, const uint8_t *b, int n, float blend);
void orc_blend_into_f32_u8(float *a, const uint8_t *b, int n, float blend);

/* ---- shared counter-based RNG ---- */
uint32_t orc_rng_u32(uint64_t seed, uint64_t pixel, uint32_t draw);
uint32_t orc_rng_range(uint64_t seed, uint64_t pixel, uint32_t draw, uint32_t n); /* uniform in [0, n) */

/*
 * OutlierProcessor::process + calc_pixel (src/chrono.rs:73-494) on an in-memory stack.
 *   stack        frame-major [n_frames][height][width][channels] u8, tightly packed
 *   indices      window: ascending frame indices (image_indices, chrono.rs:102-139), or NULL for all frames
 *   sample_pos   positions inside the window used for median/IQR (sample_indices, chrono.rs:151-163), or NULL for all
 *   out_image / out_mask   [height][width][channels] u8 (mask may be NULL)
 * Returns 0, or a negative value for arguments on which the reference would panic.
 */
int orc_outlier(const uint8_t *stack, int n_frames, int height, int width, int channels,
                const orc_outlier_params *prm,
                const int32_t *indices, int n_indices,
                const int32_t *sample_pos, int n_sample,
                uint8_t *out_image, uint8_t *out_mask, uint64_t *n_warnings, const orc_debug *dbg,
                int n_threads);

/* SimpleProcessor::process (src/simple.rs:26-176), same stack/window conventions. */
int orc_simple(const uint8_t *stack, int n_frames, int height, int width, int channels,
               int darker, const float weights[4], const orc_fade *fade,
               const int32_t *indices, int n_indices,
               uint8_t *out_image, int n_threads);

/* Crop::create (src/shake.rs:136-176). offsets: n x (dx,dy). out_xy: n x (x,y). Returns 0 if no crop (all zero), else 1. */
int orc_crop_create(const int32_t *offsets, int n, int width, int height, int32_t *out_xy, int32_t *out_w, int32_t *out_h);

/*
 * Window index math of create_video / create_video_simple (src/main.rs:230-286, :349-404).
 * *_has flags say whether start/end are Some. For output frame i writes start/end (exclusive) of its window
 * (step = in_step); an empty window (start >= end) means the reference skips the frame.
 * Returns the number of output frames, writes at most cap entries.
 */
/* ShakeAnalyzer (src/shake.rs:190-386): anchor windows of the first frame, then per frame the sums of squared differences
 * over the search square and the first minimum. Both return -1 when a coordinate leaves the image (the reference panics). */
int orc_shake_fill_windows(const uint8_t *image, int width, int height, int channels, const int32_t *anchors_xy, int n_anchors,
                           int anchor_radius, uint8_t *windows);
int orc_shake_offset(const uint8_t *image, int width, int height, int channels, const int32_t *anchors_xy, int n_anchors,
                     int anchor_radius, int search_radius, const uint8_t *windows, int32_t *diffs, int32_t *out_dx, int32_t *out_dy);

int orc_video_windows(int image_count,
                      int in_has_start, int in_start, int in_has_end, int in_end, int in_step,
                      int out_has_start, int out_start, int out_has_end, int out_end, int out_step,
                      int32_t *win_start, int32_t *win_end, int32_t *out_number, int cap);

/* Synthetic series of the bench / tests (twin of the device generator, not part of the reference): frames [f0, f0 + n_out)
 * of an n_frames series, rows [row0, row0 + rows) of a width x full_height image, out = [n_out][rows][width][channels]. */
int orc_synth_frames(int kind, uint64_t seed, int f0, int n_out, int n_frames, int width, int full_height, int channels,
                     int row0, int rows, uint8_t *out, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
