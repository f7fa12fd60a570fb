/*
 * chrono_oracle.c -- TEST INFRASTRUCTURE ONLY (see chrono_oracle.h for the scope statement and the parity pin).
 *
 * Plain-C restatement of chrono-photo v0.6.5's compositing arithmetic. Build with
 *   gcc -O2 -fPIC -shared -ffp-contract=off -fno-fast-math -pthread
 * (-ffp-contract=off: Rust never contracts a*b+c into an FMA; GCC on x86-64 would).
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 */
#include "chrono_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* Rust `f32 as u8`: saturating, NaN -> 0. */
static inline uint8_t sat_u8(float v) {
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v; /* trunc */
}

/* Rust `f32::signum`: 1.0 for +x/+0.0, -1.0 for -x/-0.0, NaN for NaN. */
static inline float signum_f32(float w) {
    if (!(w == w)) return w;
    return signbit(w) ? -1.0f : 1.0f;
}

/* ---------------------------------------------------------------- options.rs */

/* src/options.rs:197-213 */
void orc_threshold_new(int absolute, float min, float max, orc_threshold *out) {
    out->absolute = absolute ? 1 : 0;
    if (absolute) {
        out->min = min * 255.0f;
        out->max = max * 255.0f;
        out->scale = 1.0f / ((max - min) * 255.0f);
    } else {
        out->min = min;
        out->max = max;
        out->scale = 1.0f / (max - min);
    }
}

/* src/options.rs:223-231 */
float orc_threshold_blend_value(const orc_threshold *t, float dist) {
    if (dist <= t->min) return 0.0f;
    if (dist >= t->max) return 1.0f;
    return (dist - t->min) * t->scale;
}

/* src/options.rs:69-94 */
int orc_fade_build(const int32_t *frames, const float *vals, int n_pairs, float *out_values, int out_cap, int32_t *out_offset) {
    if (n_pairs < 2) return -1; /* the reference indexes frames[idx + 1]: panics with fewer than two pairs */
    int32_t offset = frames[0];
    int32_t len = frames[n_pairs - 1] - offset;
    if (len < 0 || len + 1 > out_cap) return -1;
    int idx = 0;
    for (int32_t i = 0; i < len + 1; i++) {
        if (idx + 1 >= n_pairs) return -1; /* index panic in the reference */
        int32_t f1 = frames[idx], f2 = frames[idx + 1];
        float v1 = vals[idx], v2 = vals[idx + 1];
        int32_t frame = i + offset;
        out_values[i] = v1 + (v2 - v1) * (float)(frame - f1) / (float)(f2 - f1);
        if (frame == f2 && idx + 2 < n_pairs) idx += 1;
        /* note: the reference increments idx unconditionally at frame == f2; at the very last frame that is
           harmless because the loop ends. Guarded here only to keep the index check above meaningful. */
    }
    *out_offset = offset;
    return len + 1;
}

/* src/options.rs:113-139 */
float orc_fade_get(const orc_fade *f, int32_t frame) {
    if (f->is_none) return 1.0f;
    int32_t i = frame - f->offset;
    int32_t len = f->n_values;
    if (i >= 0 && i < len) return f->values[i];
    if (f->mode == 0) { /* clamp */
        return i < 0 ? f->values[0] : f->values[len - 1];
    }
    while (i < 0) i += len; /* repeat */
    i = i % len;
    return f->values[i];
}

/* src/chrono.rs:496-502 and src/simple.rs:170-176 */
static inline float fade_for(const orc_fade *f, int32_t frame, int32_t total, int32_t offset) {
    if (f->absolute) return orc_fade_get(f, offset + frame);
    return orc_fade_get(f, total - frame - 1);
}

/* ---------------------------------------------------------------- chrono.rs order statistics */

/* src/chrono.rs:582-591 */
float orc_median(const uint8_t *d, size_t len) {
    if ((len + 1) % 2 == 0) return (float)d[(len + 1) / 2 - 1];
    size_t idx = (len + 1) / 2;
    return 0.5f * ((float)d[idx - 1] + (float)d[idx]);
}

/* src/chrono.rs:568-579 */
float orc_quantile(const uint8_t *d, size_t len, float q) {
    float pos = (float)(len + 1) * q;
    size_t p1 = (size_t)pos - 1; /* `pos as usize - 1`; len >= 3 is checked by the callers */
    float frac = pos - truncf(pos);
    if (frac < 0.001f) return (float)d[p1];
    if (frac > 0.999f) return (float)d[p1 + 1];
    return (1.0f - frac) * (float)d[p1] + frac * (float)d[p1 + 1];
}

/* src/chrono.rs:559-565 */
void orc_quartiles(const uint8_t *d, size_t len, float *q1, float *med, float *q3) {
    *q1 = orc_quantile(d, len, 0.25f);
    *med = orc_median(d, len);
    *q3 = orc_quantile(d, len, 0.75f);
}

/* ---------------------------------------------------------------- color.rs */

/* src/color.rs:4-16 */
void orc_blend_into_u8(uint8_t *a, const uint8_t *b, int n, float blend) {
    if (blend <= 0.0f) {
    } else if (blend >= 1.0f) {
        for (int i = 0; i < n; i++) a[i] = b[i];
    } else {
        for (int i = 0; i < n; i++) {
            float aa = (float)a[i];
            float t = ((float)b[i] - aa) * blend;
            a[i] = sat_u8(roundf(aa + t));
        }
    }
}

/* src/color.rs:32-44 */
void orc_blend_into_f32_u8(float *a, const uint8_t *b, int n, float blend) {
    if (blend <= 0.0f) {
    } else if (blend >= 1.0f) {
        for (int i = 0; i < n; i++) a[i] = (float)b[i];
    } else {
        for (int i = 0; i < n; i++) {
            float aa = a[i];
            float t = ((float)b[i] - aa) * blend;
            a[i] = aa + t;
        }
    }
}

/* ---------------------------------------------------------------- shared RNG (replaces rand::thread_rng) */

uint32_t orc_rng_u32(uint64_t seed, uint64_t pixel, uint32_t draw) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (pixel + 1) + 0xD1B54A32D192ED03ULL * (uint64_t)draw;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}

uint32_t orc_rng_range(uint64_t seed, uint64_t pixel, uint32_t draw, uint32_t n) {
    return (uint32_t)(((uint64_t)orc_rng_u32(seed, pixel, draw) * (uint64_t)n) >> 32);
}

/* ---------------------------------------------------------------- chrono.rs calc_pixel */

/* Stand-in for slice::sort_unstable (src/chrono.rs:242). Any correct sort gives the same sorted bytes; a counting sort
 * (insertion sort for short slices) is used so that the CPU baseline is not handicapped by libc qsort's indirect
 * compare calls -- pdqsort on u8 keys is at best on par with this. */
static void sort_u8(uint8_t *v, int n) {
    if (n <= 24) {
        for (int i = 1; i < n; i++) {
            uint8_t k = v[i];
            int j = i - 1;
            while (j >= 0 && v[j] > k) { v[j + 1] = v[j]; j--; }
            v[j + 1] = k;
        }
        return;
    }
    uint32_t hist[256];
    memset(hist, 0, sizeof hist);
    for (int i = 0; i < n; i++) hist[v[i]]++;
    int o = 0;
    for (int b = 0; b < 256; b++) {
        uint32_t c = hist[b];
        if (c) { memset(v + o, b, c); o += (int)c; }
    }
}

typedef struct {
    /* ThreadData, src/chrono.rs:23-28 */
    int32_t *out_idx;
    float *out_dist;
    int32_t *non_outlier;
    uint8_t *values;
} scratch_t;

typedef struct {
    const uint8_t *stack;
    int n_frames, height, width, channels;
    const orc_outlier_params *prm;
    const int32_t *indices; /* window position -> frame index */
    int n;                  /* window length ("samples") */
    const int32_t *sample_pos;
    int n_sample;
    int32_t frame_offset;
    uint8_t *out_image, *out_mask;
    const orc_debug *dbg;
} job_t;

/* src/chrono.rs:505-530 */
static int first_excluded(const scratch_t *s, int samples, int num_outliers, int *warning) {
    if (num_outliers == samples) {
        *warning = 1;
        return 0;
    }
    int excl = 0;
    for (int i = 0; i < samples; i++) {
        if (excl < num_outliers && i == s->out_idx[excl]) excl++;
        else {
            *warning = 0;
            return i;
        }
    }
    *warning = 0;
    return -1; /* unreachable: Err() in the reference */
}

/* src/chrono.rs:532-555, including the position-swap quirk described in SURVEY.md a7 */
static int sample_excluded(scratch_t *s, int samples, int num_outliers, uint64_t seed, uint64_t pixel, int *warning) {
    if (num_outliers == samples) {
        *warning = 1;
        return (int)orc_rng_range(seed, pixel, 0, (uint32_t)samples);
    }
    for (int i = 0; i < samples; i++) s->non_outlier[i] = i;
    int candidates = samples;
    for (int k = 0; k < num_outliers; k++) {
        int a = s->out_idx[k], b = candidates - 1;
        int32_t t = s->non_outlier[a];
        s->non_outlier[a] = s->non_outlier[b];
        s->non_outlier[b] = t;
        candidates--;
    }
    *warning = 0;
    return s->non_outlier[orc_rng_range(seed, pixel, 0, (uint32_t)candidates)];
}

/* src/chrono.rs:208-494. `pix` addresses the pixel inside a frame (byte offset). Returns the mask byte. */
static uint8_t calc_pixel(const job_t *J, scratch_t *S, size_t pix, uint64_t pixel_id, uint8_t *pixel, int *has_warning, size_t dbg_p) {
    const orc_outlier_params *P = J->prm;
    const int channels = J->channels;
    const int samples = J->n;
    const int sub = J->n_sample;
    const size_t fstride = (size_t)J->height * J->width * channels;
    const uint8_t *data = J->stack;
#define SAMPLE_PTR(pos) (data + (size_t)J->indices[(pos)] * fstride + pix)

    const float threshold_sq = P->threshold.min * P->threshold.min; /* :220 */
    float median[4] = {0, 0, 0, 0}, iqr_inv[4] = {0, 0, 0, 0};
    float dq1[4] = {0, 0, 0, 0}, dq3[4] = {0, 0, 0, 0};

    /* :227-235 gather */
    for (int si = 0; si < sub; si++) {
        const uint8_t *p = SAMPLE_PTR(J->sample_pos[si]);
        for (int ch = 0; ch < channels; ch++)
            if (P->weights[ch] != 0.0f) S->values[ch * sub + si] = p[ch];
    }
    /* :238-255 medians / inverse IQR */
    for (int i = 0; i < channels; i++) {
        if (P->weights[i] != 0.0f) {
            uint8_t *sl = S->values + (size_t)i * sub;
            sort_u8(sl, sub);
            if (P->threshold.absolute) {
                median[i] = orc_median(sl, (size_t)sub);
            } else {
                float q1, med, q3;
                orc_quartiles(sl, (size_t)sub, &q1, &med, &q3);
                median[i] = med;
                dq1[i] = q1;
                dq3[i] = q3;
                iqr_inv[i] = q3 - q1;
                if (iqr_inv[i] == 0.0f) iqr_inv[i] = 1.0f;
                iqr_inv[i] = 1.0f / iqr_inv[i];
            }
        }
    }

    /* :257-288 distance + classification */
    int num_outliers = 0;
    float max_dist_sq = 0.0f;
    int max_index = 0;
    for (int s = 0; s < samples; s++) {
        const uint8_t *p = SAMPLE_PTR(s);
        float dist_sq = 0.0f;
        for (int i = 0; i < channels; i++) {
            float w = P->weights[i];
            if (w != 0.0f) {
                float diff = median[i] - (float)p[i];
                float term;
                if (diff == 0.0f) {
                    term = 0.0f;
                } else if (P->threshold.absolute) {
                    float t = w * diff;
                    term = signum_f32(w) * (t * t);
                } else {
                    float t = w * iqr_inv[i];
                    t = t * diff;
                    term = signum_f32(w) * (t * t);
                }
                dist_sq += term;
            }
        }
        if (dist_sq >= threshold_sq) {
            S->out_idx[num_outliers] = s;
            S->out_dist[num_outliers] = dist_sq;
            num_outliers++;
            if (dist_sq > max_dist_sq) {
                max_dist_sq = dist_sq;
                max_index = s;
            }
        }
    }

    const int has_outliers = num_outliers > 0;
    *has_warning = 0;
    int bg_index = -1, sel_index = -1;

    /* :294-375 background */
    switch (P->background) {
    case ORC_BG_AVERAGE: {
        float mean[4] = {0, 0, 0, 0};
        for (int s = 0; s < samples; s++) {
            const uint8_t *p = SAMPLE_PTR(s);
            for (int i = 0; i < channels; i++) mean[i] += (float)p[i];
        }
        for (int i = 0; i < channels; i++) mean[i] /= (float)samples;
        if (has_outliers) {
            if (num_outliers == 1) {
                const uint8_t *sample = SAMPLE_PTR(S->out_idx[0]);
                for (int ch = 0; ch < channels; ch++) {
                    float a = (float)samples / (float)(samples - 1);
                    float b = mean[ch] * a;
                    float c = (float)sample[ch] / (float)samples;
                    pixel[ch] = sat_u8(roundf(b - c));
                }
            } else {
                float outlier_sum[4] = {0, 0, 0, 0};
                for (int k = 0; k < num_outliers; k++) {
                    const uint8_t *p = SAMPLE_PTR(S->out_idx[k]);
                    for (int ch = 0; ch < channels; ch++) outlier_sum[ch] += (float)p[ch];
                }
                int num_non = samples - num_outliers;
                for (int ch = 0; ch < channels; ch++) {
                    float a = (float)samples / (float)num_non;
                    float b = mean[ch] * a;
                    float c = outlier_sum[ch] / (float)samples;
                    pixel[ch] = sat_u8(roundf(b - c));
                }
            }
        } else {
            for (int ch = 0; ch < channels; ch++) pixel[ch] = sat_u8(roundf(mean[ch]));
        }
        break;
    }
    case ORC_BG_MEDIAN:
        for (int ch = 0; ch < channels; ch++) pixel[ch] = sat_u8(roundf(median[ch]));
        break;
    default: {
        int warning = 0, idx;
        if (P->background == ORC_BG_FIRST) {
            idx = has_outliers ? first_excluded(S, samples, num_outliers, &warning) : 0;
        } else {
            idx = has_outliers ? sample_excluded(S, samples, num_outliers, P->seed, pixel_id, &warning)
                               : (int)orc_rng_range(P->seed, pixel_id, 0, (uint32_t)samples);
        }
        const uint8_t *sample = SAMPLE_PTR(idx);
        for (int ch = 0; ch < channels; ch++) pixel[ch] = sample[ch];
        if (warning) *has_warning = 1;
        bg_index = idx;
    }
    }

    uint8_t mask = 0;
    if (has_outliers) {
        if (num_outliers == 1) { /* :379-388 */
            int sidx = S->out_idx[0];
            float dist_sq = S->out_dist[0];
            const uint8_t *sample = SAMPLE_PTR(sidx);
            float fade = fade_for(&P->fade, sidx, samples, J->frame_offset);
            float blend = fade * orc_threshold_blend_value(&P->threshold, sqrtf(dist_sq));
            orc_blend_into_u8(pixel, sample, channels, blend);
            mask = sat_u8(roundf(blend * 255.0f));
            sel_index = sidx;
        } else if (P->outlier == ORC_OUT_FORWARD || P->outlier == ORC_OUT_BACKWARD) { /* :391-427 */
            float pix_new[4] = {0, 0, 0, 0};
            float blend_inv = 1.0f;
            for (int ch = 0; ch < channels; ch++) pix_new[ch] = (float)pixel[ch];
            for (int kk = 0; kk < num_outliers; kk++) {
                int k = (P->outlier == ORC_OUT_FORWARD) ? kk : (num_outliers - 1 - kk);
                int sidx = S->out_idx[k];
                const uint8_t *sample = SAMPLE_PTR(sidx);
                float fade = fade_for(&P->fade, sidx, samples, J->frame_offset);
                float blend = fade * orc_threshold_blend_value(&P->threshold, sqrtf(S->out_dist[k]));
                orc_blend_into_f32_u8(pix_new, sample, channels, blend);
                blend_inv *= 1.0f - blend;
            }
            for (int ch = 0; ch < channels; ch++) pixel[ch] = sat_u8(roundf(pix_new[ch]));
            mask = sat_u8(roundf((1.0f - blend_inv) * 255.0f));
        } else {
            uint8_t temp_sample[4] = {0, 0, 0, 0};
            const uint8_t *sample;
            int sidx;
            float dist;
            if (P->outlier == ORC_OUT_AVERAGE) { /* :430-468 */
                float mean[4] = {0, 0, 0, 0};
                float mean_dist = 0.0f;
                for (int k = 0; k < num_outliers; k++) {
                    const uint8_t *p = SAMPLE_PTR(S->out_idx[k]);
                    for (int ch = 0; ch < channels; ch++) mean[ch] += (float)p[ch];
                    mean_dist += sqrtf(S->out_dist[k]);
                }
                for (int ch = 0; ch < channels; ch++) temp_sample[ch] = sat_u8(roundf(mean[ch] / (float)num_outliers));
                sidx = 0;
                sample = temp_sample;
                dist = mean_dist / (float)num_outliers;
            } else { /* :470-483 */
                float dist_sq;
                switch (P->outlier) {
                case ORC_OUT_FIRST: sidx = S->out_idx[0]; dist_sq = S->out_dist[0]; break;
                case ORC_OUT_LAST: sidx = S->out_idx[num_outliers - 1]; dist_sq = S->out_dist[num_outliers - 1]; break;
                case ORC_OUT_EXTREME: sidx = max_index; dist_sq = max_dist_sq; break;
                default: sidx = 0; dist_sq = 0.0f; break;
                }
                sample = SAMPLE_PTR(sidx);
                dist = sqrtf(dist_sq);
                sel_index = sidx;
            }
            float fade = fade_for(&P->fade, sidx, samples, J->frame_offset); /* :485-488 */
            float blend = fade * orc_threshold_blend_value(&P->threshold, dist);
            orc_blend_into_u8(pixel, sample, channels, blend);
            mask = sat_u8(roundf(blend * 255.0f));
        }
    }

    if (J->dbg) {
        const orc_debug *D = J->dbg;
        for (int ch = 0; ch < 4; ch++) {
            if (D->median) D->median[dbg_p * 4 + ch] = median[ch];
            if (D->q1) D->q1[dbg_p * 4 + ch] = dq1[ch];
            if (D->q3) D->q3[dbg_p * 4 + ch] = dq3[ch];
        }
        if (D->n_outliers) D->n_outliers[dbg_p] = num_outliers;
        if (D->sel_index) D->sel_index[dbg_p] = sel_index;
        if (D->bg_index) D->bg_index[dbg_p] = bg_index;
    }
#undef SAMPLE_PTR
    return mask;
}

typedef struct {
    const job_t *J;
    int row0, row1;
    uint64_t warnings;
    int err;
} worker_t;

/* src/chrono.rs:169-192: the per-pixel loop and the output / mask packing */
static void *outlier_worker(void *arg) {
    worker_t *W = (worker_t *)arg;
    const job_t *J = W->J;
    scratch_t S;
    S.out_idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)J->n);
    S.out_dist = (float *)malloc(sizeof(float) * (size_t)J->n);
    S.non_outlier = (int32_t *)malloc(sizeof(int32_t) * (size_t)J->n);
    S.values = (uint8_t *)malloc((size_t)J->n_sample * 4);
    if (!S.out_idx || !S.out_dist || !S.non_outlier || !S.values) {
        W->err = -2;
        return NULL;
    }
    const int C = J->channels;
    uint8_t pixel[4];
    for (int y = W->row0; y < W->row1; y++) {
        for (int x = 0; x < J->width; x++) {
            size_t p = (size_t)y * J->width + x;
            int warn = 0;
            uint8_t blend = calc_pixel(J, &S, p * C, J->prm->pixel_offset + p, pixel, &warn, p);
            if (warn) W->warnings++;
            for (int ch = 0; ch < C; ch++) {
                J->out_image[p * C + ch] = pixel[ch];
                if (J->out_mask) J->out_mask[p * C + ch] = (ch < 3) ? blend : 255;
            }
        }
    }
    free(S.out_idx);
    free(S.out_dist);
    free(S.non_outlier);
    free(S.values);
    return NULL;
}

int orc_outlier(const uint8_t *stack, int n_frames, int height, int width, int channels,
                const orc_outlier_params *prm, const int32_t *indices, int n_indices,
                const int32_t *sample_pos, int n_sample, uint8_t *out_image, uint8_t *out_mask,
                uint64_t *n_warnings, const orc_debug *dbg, int n_threads) {
    if (channels < 1 || channels > 4 || n_frames < 1) return -1;
    int n = indices ? n_indices : n_frames;
    if (n < 1) return -1;
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < n; i++) {
        idx[i] = indices ? indices[i] : i;
        if (idx[i] < 0 || idx[i] >= n_frames || (i > 0 && idx[i] <= idx[i - 1])) {
            free(idx);
            return -1;
        }
    }
    int ns = sample_pos ? n_sample : n;
    int32_t *sp = (int32_t *)malloc(sizeof(int32_t) * (size_t)ns);
    for (int i = 0; i < ns; i++) {
        sp[i] = sample_pos ? sample_pos[i] : i;
        if (sp[i] < 0 || sp[i] >= n) {
            free(idx);
            free(sp);
            return -1;
        }
    }
    /* quantile() underflows `pos as usize - 1` for fewer than 3 samples (chrono.rs:569-570): index panic */
    if (!prm->threshold.absolute && ns < 3) {
        free(idx);
        free(sp);
        return -3;
    }
    job_t J;
    J.stack = stack;
    J.n_frames = n_frames;
    J.height = height;
    J.width = width;
    J.channels = channels;
    J.prm = prm;
    J.indices = idx;
    J.n = n;
    J.sample_pos = sp;
    J.n_sample = ns;
    J.frame_offset = indices ? indices[0] : 0; /* chrono.rs:102-103 */
    J.out_image = out_image;
    J.out_mask = out_mask;
    J.dbg = dbg;

    if (n_threads < 1) n_threads = 1;
    if (n_threads > height) n_threads = height;
    worker_t *W = (worker_t *)calloc((size_t)n_threads, sizeof(worker_t));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; t++) {
        W[t].J = &J;
        W[t].row0 = (int)((long long)height * t / n_threads);
        W[t].row1 = (int)((long long)height * (t + 1) / n_threads);
    }
    if (n_threads == 1) outlier_worker(&W[0]);
    else {
        for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, outlier_worker, &W[t]);
        for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    }
    uint64_t warn = 0;
    int err = 0;
    for (int t = 0; t < n_threads; t++) {
        warn += W[t].warnings;
        if (W[t].err) err = W[t].err;
    }
    if (n_warnings) *n_warnings = warn;
    free(W);
    free(th);
    free(idx);
    free(sp);
    return err;
}

/* ---------------------------------------------------------------- simple.rs */

typedef struct {
    const uint8_t *stack;
    int height, width, channels, darker;
    const float *weights;
    const orc_fade *fade;
    const int32_t *indices;
    int n;
    int32_t frame_offset;
    uint8_t *out;
    float *extreme;
    int row0, row1;
} simple_worker_t;

/* src/simple.rs:43-136: frames strictly in order, per-pixel strict compare, running blend */
static void *simple_worker(void *arg) {
    simple_worker_t *W = (simple_worker_t *)arg;
    const int C = W->channels;
    const size_t fstride = (size_t)W->height * W->width * C;
    for (int s = 0; s < W->n; s++) {
        const uint8_t *frame = W->stack + (size_t)W->indices[s] * fstride;
        float fade = fade_for(W->fade, s, W->n, W->frame_offset); /* :122 (same for every pixel of the frame) */
        for (size_t p = (size_t)W->row0 * W->width; p < (size_t)W->row1 * W->width; p++) {
            const uint8_t *in_pix = frame + p * C;
            uint8_t *out_pix = W->out + p * C;
            float value = 0.0f;
            for (int ch = 0; ch < C; ch++) {
                float t = (float)in_pix[ch] * W->weights[ch]; /* :105 */
                value += t;
            }
            int is_extreme = W->darker ? (value < W->extreme[p]) : (value > W->extreme[p]); /* :108-118 */
            if (is_extreme) {
                W->extreme[p] = value;
                if (fade > 0.0f) {
                    if (fade >= 1.0f) {
                        for (int ch = 0; ch < C; ch++) out_pix[ch] = in_pix[ch];
                    } else {
                        orc_blend_into_u8(out_pix, in_pix, C, fade);
                    }
                }
            }
        }
    }
    return NULL;
}

int orc_simple(const uint8_t *stack, int n_frames, int height, int width, int channels, int darker,
               const float weights[4], const orc_fade *fade, const int32_t *indices, int n_indices,
               uint8_t *out_image, int n_threads) {
    if (channels < 1 || channels > 4 || n_frames < 1) return -1;
    int n = indices ? n_indices : n_frames;
    if (n < 1) return -1;
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < n; i++) {
        idx[i] = indices ? indices[i] : i;
        if (idx[i] < 0 || idx[i] >= n_frames) {
            free(idx);
            return -1;
        }
    }
    size_t P = (size_t)height * width;
    float *extreme = (float *)malloc(sizeof(float) * P);
    for (size_t p = 0; p < P; p++) extreme[p] = darker ? 3.40282347e+38f : -3.40282347e+38f; /* :75-83 f32::MAX / f32::MIN */
    memset(out_image, 0, P * channels);                                                      /* :71-74 */
    if (n_threads < 1) n_threads = 1;
    if (n_threads > height) n_threads = height;
    simple_worker_t *W = (simple_worker_t *)calloc((size_t)n_threads, sizeof(simple_worker_t));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; t++) {
        W[t].stack = stack;
        W[t].height = height;
        W[t].width = width;
        W[t].channels = channels;
        W[t].darker = darker;
        W[t].weights = weights;
        W[t].fade = fade;
        W[t].indices = idx;
        W[t].n = n;
        W[t].frame_offset = indices ? indices[0] : 0; /* :54-57 */
        W[t].out = out_image;
        W[t].extreme = extreme;
        W[t].row0 = (int)((long long)height * t / n_threads);
        W[t].row1 = (int)((long long)height * (t + 1) / n_threads);
    }
    if (n_threads == 1) simple_worker(&W[0]);
    else {
        for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, simple_worker, &W[t]);
        for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    }
    free(W);
    free(th);
    free(extreme);
    free(idx);
    return 0;
}

/* ---------------------------------------------------------------- shake.rs Crop::create */

/* src/shake.rs:136-176 */
int orc_crop_create(const int32_t *offsets, int n, int width, int height, int32_t *out_xy, int32_t *out_w, int32_t *out_h) {
    int32_t xmin = 0, ymin = 0, xmax = 0, ymax = 0;
    for (int i = 0; i < n; i++) {
        int32_t x = offsets[2 * i], y = offsets[2 * i + 1];
        if (x < xmin) xmin = x;
        if (y < ymin) ymin = y;
        if (x > xmax) xmax = x;
        if (y > ymax) ymax = y;
    }
    if (xmin == 0 && ymin == 0 && xmax == 0 && ymax == 0) return 0;
    *out_w = width + xmin - xmax;
    *out_h = height + ymin - ymax;
    for (int i = 0; i < n; i++) {
        out_xy[2 * i] = -xmin + offsets[2 * i];
        out_xy[2 * i + 1] = -ymin + offsets[2 * i + 1];
    }
    return 1;
}

/* ---------------------------------------------------------------- shake.rs ShakeAnalyzer */

/* fill_windows (src/shake.rs:307-336): the (2r+1)^2 x C bytes around every anchor of the FIRST frame.
 * Returns 0, or -1 when a coordinate leaves the image (the reference panics: "Image coordinate out of range"). */
int orc_shake_fill_windows(const uint8_t *image, int width, int height, int channels, const int32_t *anchors_xy, int n_anchors,
                           int anchor_radius, uint8_t *windows) {
    const int size = 2 * anchor_radius + 1;
    const size_t win_len = (size_t)size * size * channels;
    for (int i = 0; i < n_anchors; i++) {
        uint8_t *win = windows + (size_t)i * win_len;
        const int cx = anchors_xy[2 * i], cy = anchors_xy[2 * i + 1];
        for (int dy = 0; dy < size; dy++) {
            const int yy = cy + dy - anchor_radius;
            for (int dx = 0; dx < size; dx++) {
                const int xx = cx + dx - anchor_radius;
                if (xx < 0 || yy < 0 || xx >= width || yy >= height) return -1;
                for (int ch = 0; ch < channels; ch++)
                    win[((size_t)dy * size + dx) * channels + ch] = image[((size_t)yy * width + xx) * channels + ch];
            }
        }
    }
    return 0;
}

/* calc_diffs (src/shake.rs:338-386) + the first minimum (:275-279, Iterator::min_by_key keeps the first of equal minima).
 * diffs: (2s+1)^2 sums of squared differences over all anchors, i32 with wrap-around (release build, overflow-checks off,
 * Cargo.toml:13). Returns 0, or -1 when a coordinate leaves the image. */
int orc_shake_offset(const uint8_t *image, int width, int height, int channels, const int32_t *anchors_xy, int n_anchors,
                     int anchor_radius, int search_radius, const uint8_t *windows, int32_t *diffs, int32_t *out_dx, int32_t *out_dy) {
    const int size = 2 * anchor_radius + 1, search_size = 2 * search_radius + 1;
    const size_t win_len = (size_t)size * size * channels;
    for (int i = 0; i < search_size * search_size; i++) diffs[i] = 0;
    for (int i = 0; i < n_anchors; i++) {
        const uint8_t *win = windows + (size_t)i * win_len;
        const int cx = anchors_xy[2 * i], cy = anchors_xy[2 * i + 1];
        for (int oy = 0; oy < search_size; oy++) {
            for (int ox = 0; ox < search_size; ox++) {
                uint32_t acc = (uint32_t)diffs[oy * search_size + ox];
                for (int dy = 0; dy < size; dy++) {
                    const int yy = cy + (oy - search_radius) + dy - anchor_radius;
                    for (int dx = 0; dx < size; dx++) {
                        const int xx = cx + (ox - search_radius) + dx - anchor_radius;
                        if (xx < 0 || yy < 0 || xx >= width || yy >= height) return -1;
                        for (int ch = 0; ch < channels; ch++) {
                            const int d = (int)win[((size_t)dy * size + dx) * channels + ch] - (int)image[((size_t)yy * width + xx) * channels + ch];
                            acc += (uint32_t)(d * d);
                        }
                    }
                }
                diffs[oy * search_size + ox] = (int32_t)acc;
            }
        }
    }
    int min_idx = 0;
    for (int i = 1; i < search_size * search_size; i++)
        if (diffs[i] < diffs[min_idx]) min_idx = i;
    *out_dx = (min_idx % search_size) - search_radius;
    *out_dy = (min_idx / search_size) - search_radius;
    return 0;
}

/* ---------------------------------------------------------------- main.rs video windows */

/* src/main.rs:230-286 (identical in :349-404). Rust `%` is a remainder with the sign of the dividend, like C. */
int orc_video_windows(int image_count, int in_has_start, int in_start, int in_has_end, int in_end, int in_step,
                      int out_has_start, int out_start, int out_has_end, int out_end, int out_step,
                      int32_t *win_start, int32_t *win_end, int32_t *out_number, int cap) {
    int v_lower;
    if (out_has_start) v_lower = out_start;
    else if (in_has_start && in_has_end) v_lower = -(in_end - in_start) + 1; /* frames.range() */
    else v_lower = 0;
    int v_upper = out_has_end ? out_end : image_count;
    int count = (v_upper - v_lower) / out_step;
    int written = 0;
    for (int i = 0; i < count; i++) {
        int frame = i * out_step + v_lower;
        int start, end;
        if (in_has_start) {
            int st = frame + in_start;
            while (st < 0) st += in_step;
            int a = st % in_step, b = frame + in_start;
            start = a > b ? a : b;
        } else {
            start = 0;
        }
        if (in_has_end) {
            int a = image_count + (frame + in_end) % in_step - in_step, b = frame + in_end;
            end = a < b ? a : b;
        } else {
            end = image_count;
        }
        if (written < cap) {
            win_start[written] = start;
            win_end[written] = end;
            out_number[written] = frame - v_lower;
            written++;
        }
    }
    return count < 0 ? 0 : count;
}

/* ------------------------------------------------------------------------------------------------ synthetic series
 * Twin of the device generator (chrono_photo_b200/csrc/chb_common.cuh synth_byte; recipes of SURVEY.md 8d / DESIGN.md):
 * the bench's reference arm and the post-run verification build their input without touching the product library.
 * kind 1 = S1 (src/util/create_example_data.rs:10-49 without the JPEG round trip), 2 = S2 (gradient + uniform noise
 * +-5 + 8 discs), 3 = iid uniform bytes, 4 = gradient + Gaussian-like noise + discs. Not part of the reference. */
static inline uint32_t synth_hash32(uint64_t seed, uint32_t f, uint32_t y, uint32_t x, uint32_t ch) {
    uint64_t z = seed ^ (0x9E3779B97F4A7C15ULL * ((uint64_t)f + 1));
    z ^= ((uint64_t)y << 32) | (uint64_t)x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z += (uint64_t)ch * 0xD1B54A32D192ED03ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}

static uint8_t synth_sample(int kind, uint64_t seed, int f, int n_frames, int y, int x, int ch, int W, int H) {
    const uint32_t h = synth_hash32(seed, (uint32_t)f, (uint32_t)y, (uint32_t)x, (uint32_t)ch);
    if (kind == 3) return (uint8_t)(h >> 24);
    if (kind == 1) {
        int v = (ch == 2 ? 140 : 240) + (int)(((uint64_t)h * 10) >> 32);
        if (ch == 0) {
            int dx = x - (100 + f * 10), dy = y - (H / 3 + f * 5);
            if (dx >= -8 && dx <= 8 && dy >= -8 && dy <= 8) v = 0;
            dx = x - (W - 24);
            dy = y - (H - 68);
            if (dx >= -8 && dx <= 8 && dy >= -8 && dy <= 8) v = 0;
        }
        return (uint8_t)v;
    }
    int v = 48 + (int)(((long long)x * 96) / W) + (int)(((long long)y * 64) / H) + 16 * ch;
    if (kind == 2) v += (int)(((uint64_t)h * 11) >> 32) - 5;
    else v += (int)(h & 7) + (int)((h >> 8) & 7) + (int)((h >> 16) & 7) + (int)((h >> 24) & 7) - 14;
    for (int k = 0; k < 8; k++) { /* disc k on its linear track, wrapping around the image */
        const long long x0 = ((long long)W * (2 * k + 1)) / 16, y0 = ((long long)H * ((5 * k + 3) % 16)) / 16;
        const long long lx = 200LL * (1 + (k % 3)), ly = (k & 1) ? 200LL : -200LL;
        const long long nf = n_frames > 0 ? n_frames : 1;
        const long long cx = (x0 + (lx * f) / nf) % W, cy = ((y0 + (ly * f) / nf) % H + H) % H;
        const long long dx = x - cx, dy = y - cy;
        if (dx * dx + dy * dy <= 1600) v = (k & 1) ? 232 - 8 * ch : 24 + 8 * ch;
    }
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

typedef struct {
    int kind, f0, n_out, n_frames, W, H, C, row0, rows, t, n_threads;
    uint64_t seed;
    uint8_t *out;
} synth_job_t;

static void *synth_worker(void *arg) {
    const synth_job_t *J = (const synth_job_t *)arg;
    const long long tasks = (long long)J->n_out * J->rows;
    for (long long task = J->t; task < tasks; task += J->n_threads) {
        const int fi = (int)(task / J->rows), r = (int)(task % J->rows);
        uint8_t *o = J->out + ((size_t)fi * J->rows + r) * (size_t)J->W * J->C;
        for (int x = 0; x < J->W; x++)
            for (int c = 0; c < J->C; c++)
                o[(size_t)x * J->C + c] = synth_sample(J->kind, J->seed, J->f0 + fi, J->n_frames, J->row0 + r, x, c, J->W, J->H);
    }
    return NULL;
}

/* Frames [f0, f0 + n_out) of an n_frames series, rows [row0, row0 + rows) of a W x H image: out is [n_out][rows][W][C]. */
int orc_synth_frames(int kind, uint64_t seed, int f0, int n_out, int n_frames, int width, int full_height, int channels,
                     int row0, int rows, uint8_t *out, int n_threads) {
    if (!out || kind < 1 || kind > 4 || width < 1 || rows < 0 || n_out < 0 || channels < 1 || channels > 4) return -1;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    synth_job_t J[256];
    for (int t = 0; t < n_threads; t++) {
        J[t] = (synth_job_t){kind, f0, n_out, n_frames, width, full_height, channels, row0, rows, t, n_threads, seed, out};
        if (n_threads > 1) pthread_create(&th[t], NULL, synth_worker, &J[t]);
    }
    if (n_threads == 1) synth_worker(&J[0]);
    else
        for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    return 0;
}
