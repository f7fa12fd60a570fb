// chb_kernels.cuh -- sm_100a kernels of the compositing path.
//
//   K1 outlier_kernel   fused OutlierProcessor::calc_pixel (src/chrono.rs:208-494): exact per-band order statistics
//                       by an in-register SAD search, a certified "no outlier" filter, background / outlier policies.
//   K2 simple_kernel    SimpleProcessor::process (src/simple.rs:96-133): streaming per-pixel arg-extreme.
//   pack_frame_kernel   ingest: interleaved u8 frame -> time-sliced stack layout (replaces src/slicer.rs:185-204).
//   synth_fill_kernel   synthetic stacks generated in place.
//
// Everything is integer / f32 scalar and packed-byte work (VABSDIFF4, IDP.4A, LOP3); there is no dense contraction,
// hence no tensor-core use. Compiled with -fmad=false: Rust never contracts a*b+c, and blended bytes must round alike.
#pragma once
#include "chb_common.cuh"

namespace chb {

// ------------------------------------------------------------------------------------------------ small device utils
__device__ __forceinline__ uint32_t sad4_acc(uint32_t a, uint32_t b, uint32_t acc) {
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));  // VABSDIFF4.U8.ACC
    return d;
}
__device__ __forceinline__ uint32_t absdiff4(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(0u));  // VABSDIFF4.U8 (per-byte |a-b|)
    return d;
}
__device__ __forceinline__ uint32_t rep4(int v) { return (uint32_t)v * 0x01010101u; }
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
template <int G>
__device__ __forceinline__ uint32_t group_sum(uint32_t v) {
#pragma unroll
    for (int m = 1; m < G; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
template <int G>
__device__ __forceinline__ uint32_t group_or(uint32_t v) {
#pragma unroll
    for (int m = 1; m < G; m <<= 1) v |= __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// ------------------------------------------------------------------------------------------------ kernel arguments
struct OutlierArgs {
    const uint8_t* stack;
    long long n_pixels, n_tiles;
    int NG, C;
    int g0, n_groups;        // frame groups spanned by the window: [g0, g0 + n_groups)
    unsigned patch_slots;    // bit i: register slot i may hold bytes that are not window frames (patched for the certificate)
    int window_masked;       // 1: the span holds frames outside the window; they are masked to zero after the load
    const uint32_t* wmask;   // [capacity_groups * 4] byte masks of window frames (0xFF = in window)
    const uint32_t* smask;   // [capacity_groups * 4] byte masks of the --sample subset (SUB kernels only)
    const int32_t* win_frames;  // [n] window position -> frame index
    int n, n_sub;            // window length ("samples"), subsample size
    int first_frame;         // (frame of window position 0) & 15: its byte inside the first group
    int rk[6];               // 0-based ranks inside the subsample: q1 lo/hi, median lo/hi, q3 lo/hi
    float q1_frac, q3_frac;  // interpolation weights of quantile() (src/chrono.rs:568-579)
    float inv_n_sub;
    int absolute;
    float thr_min, thr_max, thr_scale, thr_sq;
    float w[4];
    int bg, om;
    FadeDev fade;
    int frame_offset;
    unsigned long long seed, pixel_offset;
    uint8_t* out_image;
    uint8_t* out_mask;  // may be null
    unsigned long long* counters;  // [0] warnings (all-outlier pixels), [1] pixels that left the certified fast path
    float* dbg_median; float* dbg_q1; float* dbg_q3; int* dbg_nout;
};

// ------------------------------------------------------------------------------------------------ exact order statistics
// A pixel-band's samples sit in registers as packed bytes (4 frames per word), split over G lanes. For a candidate
// value c, F(c) = sum |x - c| costs one VABSDIFF4.ACC per word; F(c+1) - F(c) = 2*#{x <= c} - CAP gives an exact count,
// and the k-th smallest value is min{c : #{x <= c} >= k+1}. Bytes that are not part of the sample are zero, which
// shifts every rank by the (known) number of such bytes.
//
// The search keeps a bracket [lo, hi] per pixel-band plus the last pair of adjacent F values (kc, F(kc), F(kc+1)):
// a probe next to that pair needs ONE new F evaluation, so walking outward from a good first guess costs one
// evaluation per step; far targets gallop, then bisect (two evaluations per probe).
template <int W4, int G>
struct Sel {
    const int cap;  // bytes held by the G lanes of one pixel-band (real + zero padding)
    // cache of adjacent F values
    int kc;
    uint32_t fk0, fk1;
    bool k0, k1;
    __device__ __forceinline__ explicit Sel(int cap_) : cap(cap_), kc(-4), fk0(0), fk1(0), k0(false), k1(false) {}

    __device__ __forceinline__ uint32_t F(const uint32_t (&x)[W4], int c) const {
        const uint32_t cc = rep4(c);
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
        for (int q = 0; q < W4; q += 4) {
            a0 = sad4_acc(x[q], cc, a0);
            a1 = sad4_acc(x[q + 1], cc, a1);
            a2 = sad4_acc(x[q + 2], cc, a2);
            a3 = sad4_acc(x[q + 3], cc, a3);
        }
        return group_sum<G>((a0 + a1) + (a2 + a3));
    }

    // Evaluates F at g-1, g, g+1 (g clamped to [1, 254]) and narrows the bracket [lo, hi] of padded rank kp.
    // mode: 0 bisect, 1 walk/gallop up, 2 walk/gallop down.
    __device__ __forceinline__ void window3(const uint32_t (&x)[W4], int kp, int& g, int& lo, int& hi, int& cnt_hi, int& mode,
                                            uint32_t& f_at_g) {
        g = g < 1 ? 1 : (g > 254 ? 254 : g);
        const uint32_t c0 = rep4(g - 1), c1 = rep4(g), c2 = rep4(g + 1);
        uint32_t f0 = 0, f1 = 0, f2 = 0, h0 = 0, h1 = 0, h2 = 0;
#pragma unroll
        for (int q = 0; q < W4; q += 2) {
            f0 = sad4_acc(x[q], c0, f0);
            f1 = sad4_acc(x[q], c1, f1);
            f2 = sad4_acc(x[q], c2, f2);
            h0 = sad4_acc(x[q + 1], c0, h0);
            h1 = sad4_acc(x[q + 1], c1, h1);
            h2 = sad4_acc(x[q + 1], c2, h2);
        }
        f0 = group_sum<G>(f0 + h0);
        f1 = group_sum<G>(f1 + h1);
        f2 = group_sum<G>(f2 + h2);
        f_at_g = f1;
        const int A = ((int)f1 - (int)f0 + cap) >> 1;  // #{x <= g-1}
        const int B = ((int)f2 - (int)f1 + cap) >> 1;  // #{x <= g}
        mode = 0;
        if (kp < A) {
            if (g - 1 < hi) { hi = g - 1; cnt_hi = A; }
            mode = 2;
            kc = g - 1; fk0 = f0; fk1 = f1; k0 = k1 = true;
        } else if (kp < B) {
            if (g >= lo && g <= hi) { lo = hi = g; cnt_hi = B; }
            kc = g; fk0 = f1; fk1 = f2; k0 = k1 = true;
        } else {
            if (g + 1 > lo) lo = g + 1;
            mode = 1;
            kc = g; fk0 = f1; fk1 = f2; k0 = k1 = true;
        }
    }

    // Narrows [lo, hi] to one value. Requires #{x <= hi} >= kp+1 (cnt_hi = that count; cap for hi = 255) and
    // #{x <= lo-1} <= kp. Warp-synchronous: every lane takes part in every evaluation (shuffles inside F).
    __device__ __forceinline__ void narrow(const uint32_t (&x)[W4], int kp, int& lo, int& hi, int& cnt_hi, int mode) {
        int steps = 0;
        while (__any_sync(0xffffffffu, lo < hi)) {
            const bool act = lo < hi;
            int c;
            const int gal = steps < 2 ? 1 : (1 << (steps - 1));
            if (mode == 1) { c = lo + gal - 1; c = c < hi - 1 ? c : hi - 1; }
            else if (mode == 2) { c = hi - gal; c = c > lo ? c : lo; }
            else c = (lo + hi) >> 1;
            if (!act) c = kc;  // idle lanes re-evaluate a cached point (result unused)
            // re-anchor the cached pair at c
            if (c == kc) {
            } else if (c == kc + 1 && k1) { kc = c; fk0 = fk1; k0 = true; k1 = false; }
            else if (c == kc - 1 && k0) { kc = c; fk1 = fk0; k1 = true; k0 = false; }
            else { kc = c; k0 = k1 = false; }
            const bool second = k0;  // F(c) known -> evaluate c+1, else evaluate c
            const uint32_t val = F(x, second ? kc + 1 : kc);
            if (second) { fk1 = val; k1 = true; } else { fk0 = val; k0 = true; }
            if (act && k0 && k1) {
                const int cnt = ((int)fk1 - (int)fk0 + cap) >> 1;  // #{x <= c}
                if (cnt >= kp + 1) {
                    hi = c; cnt_hi = cnt;
                    if (mode == 1) mode = 0; else steps++;
                } else {
                    lo = c + 1;
                    if (mode == 2) mode = 0; else steps++;
                }
            }
        }
    }

    // Two adjacent padded ranks kp1 <= kp2 <= kp1+1 with a first guess.
    __device__ __forceinline__ void pair(const uint32_t (&x)[W4], int kp1, int kp2, int guess, int& v1, int& v2, uint32_t& f_at_g, int& g_used) {
        int lo = 0, hi = 255, cnt_hi = cap, mode;
        g_used = guess;
        window3(x, kp1, g_used, lo, hi, cnt_hi, mode, f_at_g);
        narrow(x, kp1, lo, hi, cnt_hi, mode);
        v1 = lo;
        v2 = v1;
        const bool need = (kp2 != kp1) && (cnt_hi < kp2 + 1);  // the next order statistic is a larger value
        if (__any_sync(0xffffffffu, need)) {
            int lo2 = need ? v1 + 1 : 0, hi2 = need ? 255 : 0, cnt2 = cap;
            narrow(x, kp2, lo2, hi2, cnt2, 1);
            if (need) v2 = lo2;
        }
    }
};

// ------------------------------------------------------------------------------------------------ exact pixel path
// Line-for-line semantics of calc_pixel once medians / inverse IQRs are known, walking the window's frames from
// global memory (they were just streamed, so they mostly sit in L2). Taken only by pixels the certificate cannot clear;
// those are queued per warp and processed 32 at a time, one pixel per lane, so the walk runs at full SIMT width.
struct PixelSrc {
    const uint8_t* tile;  // tile base
    int NG, C, p;
    __device__ __forceinline__ uint8_t at(int frame, int c) const {
        return __ldg(tile + ((((long long)c * NG + (frame >> 4)) * kTilePixels) + p) * kUnitBytes + (frame & 15));
    }
};

// Reads a pixel's samples frame by frame, keeping the current 16-frame unit of every band in registers.
struct ColumnReader {
    const uint8_t* base;  // address of unit (c = 0, g = 0) of this pixel
    long long band_stride;  // NG * 512
    int C, cur_g;
    uint4 u[4];
    __device__ __forceinline__ ColumnReader(const PixelSrc& s)
        : base(s.tile + (long long)s.p * kUnitBytes), band_stride((long long)s.NG * kTilePixels * kUnitBytes), C(s.C), cur_g(-1) {}
    __device__ __forceinline__ void fetch(int frame, uint8_t (&px)[4]) {
        const int g = frame >> 4;
        if (g != cur_g) {  // uniform across the lanes of a batch: every lane walks the same window
            cur_g = g;
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (c < C) u[c] = __ldg(reinterpret_cast<const uint4*>(base + c * band_stride + (long long)g * (kTilePixels * kUnitBytes)));
        }
        const int wsel = (frame >> 2) & 3, sh = (frame & 3) * 8;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (c < C) {
                const uint32_t w = wsel == 0 ? u[c].x : (wsel == 1 ? u[c].y : (wsel == 2 ? u[c].z : u[c].w));
                px[c] = (uint8_t)((w >> sh) & 0xffu);
            }
        }
    }
};

struct DistCtx {  // per-band constants of the distance (src/chrono.rs:265-278)
    float med[4], fac[4], sgn[4];
    bool use[4];
};
__device__ __forceinline__ void make_dist_ctx(const OutlierArgs& a, const float (&median)[4], const float (&iqr_inv)[4], DistCtx& d) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float w = a.w[i];
        d.use[i] = (i < a.C) && (w != 0.0f);
        d.med[i] = median[i];
        d.fac[i] = a.absolute ? w : w * iqr_inv[i];  // abs: (w*diff)^2; rel: ((w*iqr_inv)*diff)^2
        d.sgn[i] = (w != w) ? w : (signbit(w) ? -1.0f : 1.0f);
    }
}
__device__ __forceinline__ float dist_sq_px(const DistCtx& d, const uint8_t (&px)[4]) {
    float dist_sq = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (d.use[i]) {
            const float diff = d.med[i] - (float)px[i];
            float term = 0.0f;
            if (diff != 0.0f) {
                const float t = d.fac[i] * diff;
                term = d.sgn[i] * (t * t);
            }
            dist_sq += term;
        }
    }
    return dist_sq;
}

__device__ __forceinline__ float blend_value(const OutlierArgs& a, float dist) {  // src/options.rs:223-231
    if (dist <= a.thr_min) return 0.0f;
    if (dist >= a.thr_max) return 1.0f;
    return (dist - a.thr_min) * a.thr_scale;
}
__device__ __forceinline__ void blend_into_u8(uint8_t (&pa)[4], const uint8_t (&pb)[4], int C, float blend) {  // src/color.rs:4-16
    if (blend <= 0.0f) return;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (i < C) {
            if (blend >= 1.0f) pa[i] = pb[i];
            else {
                float aa = (float)pa[i];
                float t = ((float)pb[i] - aa) * blend;
                pa[i] = sat_u8(roundf(aa + t));
            }
        }
    }
}
__device__ __forceinline__ void blend_into_f32_u8(float (&pa)[4], const uint8_t (&pb)[4], int C, float blend) {  // src/color.rs:32-44
    if (blend <= 0.0f) return;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (i < C) {
            if (blend >= 1.0f) pa[i] = (float)pb[i];
            else {
                float aa = pa[i];
                float t = ((float)pb[i] - aa) * blend;
                pa[i] = aa + t;
            }
        }
    }
}

// Returns the mask byte; writes the composite pixel; n_out = number of outliers; warn = all-outlier warning.
// `active` lanes hold a pixel; inactive lanes run along (uniform loops) and their results are discarded.
__device__ __forceinline__ uint8_t exact_pixel(const OutlierArgs& a, const PixelSrc& src, unsigned long long pixel_id,
                                               const float (&median)[4], const float (&iqr_inv)[4], uint8_t (&pixel)[4],
                                               int& n_out, int& warn) {
    const int n = a.n, C = a.C;
    const float thr_sq = a.thr_sq;
    DistCtx dc;
    make_dist_ctx(a, median, iqr_inv, dc);
    ColumnReader rd(src);
    // pass 1 (src/chrono.rs:261-288 plus the sums the policies need)
    int k = 0, first_idx = 0, last_idx = 0, max_index = 0, first_non = -1;
    float first_d = 0.0f, last_d = 0.0f, max_dist_sq = 0.0f, mean_dist = 0.0f;
    float out_sum[4] = {0, 0, 0, 0}, all_sum[4] = {0, 0, 0, 0};
    uint8_t px[4] = {0, 0, 0, 0};
    const bool need_all = a.bg == 2, need_avg = a.om == 3 || a.bg == 2;
    for (int s = 0; s < n; s++) {
        rd.fetch(__ldg(a.win_frames + s), px);
        const float d = dist_sq_px(dc, px);
        if (need_all) {
#pragma unroll
            for (int i = 0; i < 4; i++) all_sum[i] += (float)px[i];
        }
        if (d >= thr_sq) {
            if (k == 0) { first_idx = s; first_d = d; }
            last_idx = s; last_d = d;
            k++;
            if (d > max_dist_sq) { max_dist_sq = d; max_index = s; }
            if (need_avg) {
#pragma unroll
                for (int i = 0; i < 4; i++) out_sum[i] += (float)px[i];
                mean_dist += sqrtf(d);
            }
        } else if (first_non < 0) {
            first_non = s;
        }
    }
    n_out = k;
    warn = 0;
    const bool has_outliers = k > 0;

    // background (src/chrono.rs:294-375)
    if (a.bg == 2) {  // Average
        float mean[4];
#pragma unroll
        for (int i = 0; i < 4; i++) mean[i] = all_sum[i] / (float)n;
        if (has_outliers) {
            const float ratio = (float)n / (float)(n - k);  // k == 1: samples/(samples-1); k > 1: samples/num_non_outliers
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (i < C) pixel[i] = sat_u8(roundf(mean[i] * ratio - out_sum[i] / (float)n));
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (i < C) pixel[i] = sat_u8(roundf(mean[i]));
        }
    } else if (a.bg == 3) {  // Median
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < C) pixel[i] = sat_u8(roundf(median[i]));
    } else {
        int idx;
        if (a.bg == 0) {  // First: first_excluded (src/chrono.rs:505-530)
            if (!has_outliers) idx = 0;
            else if (k == n) { idx = 0; warn = 1; }
            else idx = first_non;
        } else {  // Random: sample_excluded (src/chrono.rs:532-555)
            if (!has_outliers) idx = (int)rng_range(a.seed, pixel_id, 0, (uint32_t)n);
            else if (k == n) { idx = (int)rng_range(a.seed, pixel_id, 0, (uint32_t)n); warn = 1; }
            else {
                // The reference swaps position idx_t with position n-1-t for the t-th outlier (t ascending), then draws
                // r < n-k and returns perm[r]. Position r is only ever written when r is itself the t-th outlier, and
                // then receives the content of position n-1-t, which no earlier swap can have touched (earlier outlier
                // positions are < r, earlier partner positions are > n-1-t). So perm[r] = n-1-t for an outlier r, else r.
                const int r = (int)rng_range(a.seed, pixel_id, 0, (uint32_t)(n - k));
                uint8_t tmp[4] = {0, 0, 0, 0};
                idx = r;
                rd.fetch(__ldg(a.win_frames + r), tmp);
                if (dist_sq_px(dc, tmp) >= thr_sq) {
                    int order = 0;  // number of outliers before r
                    for (int s = 0; s < r; s++) {
                        rd.fetch(__ldg(a.win_frames + s), tmp);
                        if (dist_sq_px(dc, tmp) >= thr_sq) order++;
                    }
                    idx = n - 1 - order;
                }
            }
        }
        const int f = __ldg(a.win_frames + idx);
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < C) pixel[i] = src.at(f, i);
    }

    if (!has_outliers) return 0;

    uint8_t sample[4] = {0, 0, 0, 0};
    if (k == 1) {  // src/chrono.rs:379-388
        const int f = __ldg(a.win_frames + first_idx);
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < C) sample[i] = src.at(f, i);
        const float fade = fade_for(a.fade, first_idx, n, a.frame_offset);
        const float blend = fade * blend_value(a, sqrtf(first_d));
        blend_into_u8(pixel, sample, C, blend);
        return sat_u8(roundf(blend * 255.0f));
    }
    if (a.om == 4 || a.om == 5) {  // forward / backward (src/chrono.rs:391-427): second walk in list order
        float pix_new[4], blend_inv = 1.0f;
#pragma unroll
        for (int i = 0; i < 4; i++) pix_new[i] = (float)pixel[i];
        // only the span [first_idx, last_idx] holds outliers
        for (int ss = first_idx; ss <= last_idx; ss++) {
            const int s = (a.om == 4) ? ss : last_idx - (ss - first_idx);
            rd.fetch(__ldg(a.win_frames + s), sample);
            const float d = dist_sq_px(dc, sample);
            if (d >= thr_sq) {
                const float fade = fade_for(a.fade, s, n, a.frame_offset);
                const float blend = fade * blend_value(a, sqrtf(d));
                blend_into_f32_u8(pix_new, sample, C, blend);
                blend_inv *= 1.0f - blend;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < C) pixel[i] = sat_u8(roundf(pix_new[i]));
        return sat_u8(roundf((1.0f - blend_inv) * 255.0f));
    }
    int sidx;
    float dist;
    if (a.om == 3) {  // average (src/chrono.rs:430-468)
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < C) sample[i] = sat_u8(roundf(out_sum[i] / (float)k));
        sidx = 0;
        dist = mean_dist / (float)k;
    } else {  // first / last / extreme (src/chrono.rs:470-483)
        float dsq;
        if (a.om == 0) { sidx = first_idx; dsq = first_d; }
        else if (a.om == 1) { sidx = last_idx; dsq = last_d; }
        else { sidx = max_index; dsq = max_dist_sq; }
        const int f = __ldg(a.win_frames + sidx);
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < C) sample[i] = src.at(f, i);
        dist = sqrtf(dsq);
    }
    const float fade = fade_for(a.fade, sidx, n, a.frame_offset);  // src/chrono.rs:485-488
    const float blend = fade * blend_value(a, dist);
    blend_into_u8(pixel, sample, C, blend);
    return sat_u8(roundf(blend * 255.0f));
}

// Median (and, for relative thresholds, quartiles and the inverse IQR) of one pixel-band (src/chrono.rs:238-255).
template <int W4, int G>
__device__ __forceinline__ void band_stats(int cap, const uint32_t (&xs)[W4], uint32_t ssum, const OutlierArgs& a,
                                           int pad, float& median, float& q1o, float& q3o, float& iqr_inv, int& center, float& halfw) {
    Sel<W4, G> sel(cap);
    int g = __float2int_rn((float)ssum * a.inv_n_sub);  // mean as the first guess
    int mlo, mhi, gc;
    uint32_t f_at_g;
    sel.pair(xs, a.rk[2] + pad, a.rk[3] + pad, g, mlo, mhi, f_at_g, gc);
    median = (mlo == mhi) ? (float)mlo : 0.5f * ((float)mlo + (float)mhi);  // src/chrono.rs:582-591
    center = (mlo + mhi) >> 1;
    halfw = median - (float)center;
    if (!a.absolute) {  // quartiles (src/chrono.rs:559-579) and inverse IQR (:246-252)
        // spread estimate for the quartile guesses: mean absolute deviation around the (clamped) guess
        const float mad = ((float)f_at_g - (float)pad * (float)gc) * a.inv_n_sub;
        const int dq = __float2int_rn(0.95f * mad);
        int alo, ahi, blo, bhi, gdummy;
        uint32_t dummy;
        sel.pair(xs, a.rk[0] + pad, a.rk[1] + pad, mlo - dq, alo, ahi, dummy, gdummy);
        sel.pair(xs, a.rk[4] + pad, a.rk[5] + pad, mhi + dq, blo, bhi, dummy, gdummy);
        const float q1 = (a.rk[0] == a.rk[1]) ? (float)alo : (1.0f - a.q1_frac) * (float)alo + a.q1_frac * (float)ahi;
        const float q3 = (a.rk[4] == a.rk[5]) ? (float)blo : (1.0f - a.q3_frac) * (float)blo + a.q3_frac * (float)bhi;
        q1o = q1;
        q3o = q3;
        float iq = q3 - q1;
        if (iq == 0.0f) iq = 1.0f;
        iqr_inv = 1.0f / iq;
    }
}

// ------------------------------------------------------------------------------------------------ K1
// One warp = one tile slice: 32/G pixels x G lanes per pixel. Each lane keeps WPL 16-frame units per band in
// registers (slot i of lane j holds frame group g0 + i*G + j), so the whole time series of the warp's pixels is
// read from HBM exactly once with 128-bit loads that are contiguous per (band, group) row.
// Pixels whose "no outlier" certificate fails are queued in shared memory (per warp) with their medians and handled
// by exact_pixel 32 at a time.
constexpr int kWarpsPerCta = 8;
constexpr int kQueueCap = 64;  // per warp
struct QueueEntry {
    long long pix;
    float median[4];
    float iqr_inv[4];
};

template <int C>
__device__ __forceinline__ void store_pixel(const OutlierArgs& a, long long pix, const uint8_t (&pixel)[4], uint8_t mask) {
    // src/chrono.rs:183-191
#pragma unroll
    for (int c = 0; c < C; c++) {
        a.out_image[pix * C + c] = pixel[c];
        if (a.out_mask) a.out_mask[pix * C + c] = (c < 3) ? mask : 255;
    }
}

template <int C>
__device__ __noinline__ void drain_queue(const OutlierArgs& a, const QueueEntry* q, int count, int lane) {
    const bool active = lane < count;
    const QueueEntry e = q[active ? lane : 0];
    const long long tile = e.pix >> 5;
    const PixelSrc src{a.stack + tile * tile_bytes(C, a.NG), a.NG, C, (int)(e.pix & 31)};
    uint8_t pixel[4] = {0, 0, 0, 0};
    int n_out = 0, warn = 0;
    const uint8_t mask = exact_pixel(a, src, a.pixel_offset + (unsigned long long)e.pix, e.median, e.iqr_inv, pixel, n_out, warn);
    if (active) {
        store_pixel<C>(a, e.pix, pixel, mask);
        if (a.dbg_nout) a.dbg_nout[e.pix] = n_out;
    }
    const unsigned wb = __ballot_sync(0xffffffffu, active && warn);
    if (lane == 0) {
        if (wb) atomicAdd(a.counters, (unsigned long long)__popc(wb));
        atomicAdd(a.counters + 1, (unsigned long long)count);
    }
}

template <int C, int WPL, int G, bool SUB>
__global__ void __launch_bounds__(kWarpsPerCta * 32, (C * WPL <= 24) ? 2 : 1) outlier_kernel(const __grid_constant__ OutlierArgs a) {
    constexpr int W4 = 4 * WPL;
    constexpr int PPW = 32 / G;
    __shared__ QueueEntry s_queue[kWarpsPerCta][kQueueCap];
    const int lane = threadIdx.x & 31, warp_in_cta = threadIdx.x >> 5;
    const int j = lane % G, pl = lane / G;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long n_tasks = a.n_tiles * G;
    const int cap = W4 * 4 * G;     // bytes per pixel-band across the G lanes
    const int pad = cap - a.n_sub;  // zero bytes that take part in the selection
    const long long tbytes = tile_bytes(C, a.NG);
    const long long band_stride = (long long)a.NG * (kTilePixels * kUnitBytes);
    QueueEntry* queue = s_queue[warp_in_cta];
    int qcount = 0;

    for (long long task = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); task < n_tasks; task += n_warps) {
        const long long tile = task / G;
        const int p = (int)(task % G) * PPW + pl;
        const long long pix = tile * kTilePixels + p;
        const bool valid = pix < a.n_pixels;
        const uint8_t* tb = a.stack + tile * tbytes;
        const uint8_t* lane_base = tb + ((long long)(a.g0 + j) * kTilePixels + p) * kUnitBytes;  // unit (c=0, slot 0) of this lane

        // ---- load the time series (only HBM read of the kernel)
        uint32_t x[C][W4];
#pragma unroll
        for (int c = 0; c < C; c++) {
#pragma unroll
            for (int i = 0; i < WPL; i++) {
                uint4 v = make_uint4(0, 0, 0, 0);
                if (i * G + j < a.n_groups) v = ldg_stream(lane_base + c * band_stride + (long long)i * (G * kTilePixels * kUnitBytes));
                x[c][4 * i + 0] = v.x; x[c][4 * i + 1] = v.y; x[c][4 * i + 2] = v.z; x[c][4 * i + 3] = v.w;
            }
        }
        if (a.window_masked) {  // frames outside the window must read as zero (whole-stack launches skip this)
#pragma unroll
            for (int i = 0; i < WPL; i++) {
                const uint4 m = __ldg(reinterpret_cast<const uint4*>(a.wmask) + (i * G + j));
#pragma unroll
                for (int c = 0; c < C; c++) { x[c][4 * i] &= m.x; x[c][4 * i + 1] &= m.y; x[c][4 * i + 2] &= m.z; x[c][4 * i + 3] &= m.w; }
            }
        }

        // ---- per band: sum, order statistics
        float median[4] = {0, 0, 0, 0}, iqr_inv[4] = {0, 0, 0, 0}, q1v[4] = {0, 0, 0, 0}, q3v[4] = {0, 0, 0, 0};
        int center[C];
        float halfw[C];
        uint32_t sum[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            center[c] = 0; halfw[c] = 0.0f; sum[c] = 0;
            if (a.bg == 2 || a.w[c] != 0.0f) {  // window sum (IDP.4A: FMA pipe)
                uint32_t s0 = 0, s1 = 0;
#pragma unroll
                for (int q = 0; q < W4; q += 2) { s0 = __dp4a(x[c][q], 0x01010101u, s0); s1 = __dp4a(x[c][q + 1], 0x01010101u, s1); }
                sum[c] = group_sum<G>(s0 + s1);
            }
            if (a.w[c] != 0.0f) {
                if (SUB) {
                    uint32_t xs[W4];
                    uint32_t s = 0;
#pragma unroll
                    for (int i = 0; i < WPL; i++) {
                        const uint4 m = __ldg(reinterpret_cast<const uint4*>(a.smask) + (i * G + j));
                        xs[4 * i] = x[c][4 * i] & m.x; xs[4 * i + 1] = x[c][4 * i + 1] & m.y;
                        xs[4 * i + 2] = x[c][4 * i + 2] & m.z; xs[4 * i + 3] = x[c][4 * i + 3] & m.w;
                    }
#pragma unroll
                    for (int q = 0; q < W4; q++) s = __dp4a(xs[q], 0x01010101u, s);
                    band_stats<W4, G>(cap, xs, group_sum<G>(s), a, pad, median[c], q1v[c], q3v[c], iqr_inv[c], center[c], halfw[c]);
                } else {
                    band_stats<W4, G>(cap, x[c], sum[c], a, pad, median[c], q1v[c], q3v[c], iqr_inv[c], center[c], halfw[c]);
                }
            }
        }

        uint4 x0[C];  // slot 0 as loaded (window position 0 for lane j == 0), before the certificate patch
#pragma unroll
        for (int c = 0; c < C; c++) x0[c] = make_uint4(x[c][0], x[c][1], x[c][2], x[c][3]);
        // ---- certificate: an upper bound of every frame's distance to the median.
        // Bytes that are not window frames (zero in the registers) are replaced by the band's centre value so that they
        // contribute |c - c| = 0. Whole-stack launches only have such bytes in the last slot(s).
        if (a.patch_slots) {
#pragma unroll
            for (int i = 0; i < WPL; i++) {
                if ((a.patch_slots >> i) & 1u) {
                    const uint4 m = __ldg(reinterpret_cast<const uint4*>(a.wmask) + (i * G + j));
#pragma unroll
                    for (int c = 0; c < C; c++) {
                        const uint32_t cc = rep4(center[c]);
                        x[c][4 * i] |= cc & ~m.x; x[c][4 * i + 1] |= cc & ~m.y; x[c][4 * i + 2] |= cc & ~m.z; x[c][4 * i + 3] |= cc & ~m.w;
                    }
                }
            }
        }
        float bound = 0.0f;
#pragma unroll
        for (int c = 0; c < C; c++) {
            const float w = a.w[c];
            if (w != 0.0f && !(w < 0.0f)) {  // negative weights only lower dist_sq; NaN poisons the bound (-> exact path)
                const uint32_t cc = rep4(center[c]);
                uint32_t o0 = 0, o1 = 0;
#pragma unroll
                for (int q = 0; q < W4; q += 2) {
                    o0 |= absdiff4(x[c][q], cc);
                    o1 |= absdiff4(x[c][q + 1], cc);
                }
                uint32_t o = o0 | o1;
                o |= o >> 16;
                o |= o >> 8;
                o = group_or<G>(o & 0xffu);  // >= max over frames of |x - centre| (OR dominates max)
                const float aw = a.absolute ? w : w * iqr_inv[c];
                const float t = aw * ((float)o + halfw[c]);
                bound += t * t;
            }
        }
        const bool clean = bound * 1.0001f < a.thr_sq;  // margin covers the f32 roundings of the reference's sum

        // ---- output of certified pixels; the others are queued
        const bool owner = valid && j == 0;
        if (owner && clean) {
            uint8_t pixel[4] = {0, 0, 0, 0};
            if (a.bg == 2) {
#pragma unroll
                for (int c = 0; c < C; c++) pixel[c] = sat_u8(roundf((float)sum[c] / (float)a.n));  // src/chrono.rs:297-306,335-337
            } else if (a.bg == 3) {
#pragma unroll
                for (int c = 0; c < C; c++) pixel[c] = sat_u8(roundf(median[c]));  // :340-345
            } else if (a.bg == 0) {
                // frame of window position 0 lives in slot 0 of lane j == 0 (this lane); its byte index is uniform
                const int wsel = (a.first_frame >> 2) & 3, sh = (a.first_frame & 3) * 8;
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const uint32_t wv = wsel == 0 ? x0[c].x : (wsel == 1 ? x0[c].y : (wsel == 2 ? x0[c].z : x0[c].w));
                    pixel[c] = (uint8_t)((wv >> sh) & 0xffu);
                }
            } else {
                const int pos = (int)rng_range(a.seed, a.pixel_offset + (unsigned long long)pix, 0, (uint32_t)a.n);  // :357
                const int f = __ldg(a.win_frames + pos);
                const PixelSrc src{tb, a.NG, C, p};
#pragma unroll
                for (int c = 0; c < C; c++) pixel[c] = src.at(f, c);
            }
            store_pixel<C>(a, pix, pixel, 0);
            if (a.dbg_nout) a.dbg_nout[pix] = 0;
        }
        if (owner && a.dbg_median) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                a.dbg_median[pix * 4 + c] = median[c];
                if (a.dbg_q1) a.dbg_q1[pix * 4 + c] = q1v[c];
                if (a.dbg_q3) a.dbg_q3[pix * 4 + c] = q3v[c];
            }
        }
        const bool dirty = owner && !clean;
        const unsigned db = __ballot_sync(0xffffffffu, dirty);
        if (db) {
            const int nd = __popc(db);
            if (qcount + nd > kQueueCap) {  // make room: drain full batches
                __syncwarp();
                while (qcount >= 32) { drain_queue<C>(a, queue + (qcount - 32), 32, lane); qcount -= 32; }
                __syncwarp();
            }
            if (dirty) {
                QueueEntry& e = queue[qcount + __popc(db & ((1u << lane) - 1u))];
                e.pix = pix;
#pragma unroll
                for (int c = 0; c < 4; c++) { e.median[c] = median[c]; e.iqr_inv[c] = iqr_inv[c]; }
            }
            qcount += nd;
            __syncwarp();
            while (qcount >= 32) { drain_queue<C>(a, queue + (qcount - 32), 32, lane); qcount -= 32; }
            __syncwarp();
        }
    }
    __syncwarp();
    if (qcount > 0) drain_queue<C>(a, queue, qcount, lane);
}

// ------------------------------------------------------------------------------------------------ K2
struct SimpleArgs {
    const uint8_t* stack;
    long long n_pixels, n_tiles;
    int NG, C;
    int g0, n_groups;
    const uint32_t* wmask;      // [n_groups * 4] byte masks of window frames, or null when every byte of the span is in the window
    const int32_t* pos_of_group;  // [n_groups] window position of the first window frame at or after the group start
    int n;                      // window length
    int darker;
    float w[4];
    unsigned use_mask;          // integer kernel: bit c set = band c has weight 1 (others 0)
    FadeDev fade;
    int frame_offset;
    uint8_t* out_image;
};

__device__ __forceinline__ float byte_to_float(uint32_t word, int k) {
    // (float)byte without I2F: splice the byte into the mantissa of 2^23 and subtract 2^23 (exact)
    uint32_t bits = __byte_perm(word, 0x4B000000u, 0x7540 + k);
    return __uint_as_float(bits) - 8388608.0f;
}

// General kernel: any weights, any fade. One thread per pixel streams the pixel's groups in frame order
// (src/simple.rs:138-165 processes frames strictly in order; :102-133 is the per-pixel body). FADE = false: Fade::none(),
// so the result is the pixel of the first strict extreme. FADE = true keeps the running, order-dependent blend.
template <int C, bool FADE>
__global__ void __launch_bounds__(256) simple_kernel(const __grid_constant__ SimpleArgs a) {
    const long long n_threads = (long long)gridDim.x * blockDim.x;
    const long long tbytes = tile_bytes(C, a.NG);
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < a.n_tiles * kTilePixels; pix += n_threads) {
        const long long tile = pix >> 5;
        const int p = (int)(pix & 31);
        const uint8_t* tb = a.stack + tile * tbytes;
        float extreme = a.darker ? 3.40282347e+38f : -3.40282347e+38f;  // src/simple.rs:75-83
        int best_frame = -1;
        uint8_t outp[4] = {0, 0, 0, 0};  // src/simple.rs:71-74: the buffer starts at 0
        uint4 cur[C], nxt[C];
#pragma unroll
        for (int c = 0; c < C; c++) cur[c] = ldg_stream(tb + ((((long long)c * a.NG + a.g0) * kTilePixels) + p) * kUnitBytes);
        for (int gi = 0; gi < a.n_groups; gi++) {
            if (gi + 1 < a.n_groups) {
#pragma unroll
                for (int c = 0; c < C; c++) nxt[c] = ldg_stream(tb + ((((long long)c * a.NG + (a.g0 + gi + 1)) * kTilePixels) + p) * kUnitBytes);
            }
            uint4 m = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            if (a.wmask) m = __ldg(reinterpret_cast<const uint4*>(a.wmask) + gi);
            int pos = FADE ? a.pos_of_group[gi] : 0;
#pragma unroll
            for (int wq = 0; wq < 4; wq++) {
                const uint32_t mw = wq == 0 ? m.x : (wq == 1 ? m.y : (wq == 2 ? m.z : m.w));
                uint32_t xw[C];
#pragma unroll
                for (int c = 0; c < C; c++) xw[c] = wq == 0 ? cur[c].x : (wq == 1 ? cur[c].y : (wq == 2 ? cur[c].z : cur[c].w));
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const bool in_win = (mw >> (8 * k)) & 1u;
                    float value = 0.0f;
#pragma unroll
                    for (int c = 0; c < C; c++) value += byte_to_float(xw[c], k) * a.w[c];  // src/simple.rs:103-106 (no FMA: -fmad=false)
                    const bool is_ext = in_win && (a.darker ? (value < extreme) : (value > extreme));  // :108-118
                    if (is_ext) {
                        extreme = value;
                        if (!FADE) best_frame = (a.g0 + gi) * kGroupFrames + wq * 4 + k;
                        else {
                            float fade = fade_for(a.fade, pos, a.n, a.frame_offset);  // :122
                            if (fade > 0.0f) {
                                uint8_t in_pix[4] = {0, 0, 0, 0};
#pragma unroll
                                for (int c = 0; c < C; c++) in_pix[c] = (uint8_t)((xw[c] >> (8 * k)) & 0xffu);
                                blend_into_u8(outp, in_pix, C, fade);  // fade >= 1 copies (:124-127), else blends (:128-130)
                            }
                        }
                    }
                    if (FADE) pos += in_win ? 1 : 0;
                }
            }
#pragma unroll
            for (int c = 0; c < C; c++) cur[c] = nxt[c];
        }
        if (pix < a.n_pixels) {
            if (!FADE && best_frame >= 0) {
                const PixelSrc src{tb, a.NG, C, p};
#pragma unroll
                for (int c = 0; c < C; c++) outp[c] = src.at(best_frame, c);
            }
#pragma unroll
            for (int c = 0; c < C; c++) a.out_image[pix * C + c] = outp[c];
        }
    }
}

// Integer kernel for the default case: every weight is 0 or 1 and there is no fade. The weighted sum of a frame is then
// an exact integer <= 1020 (10 bits), so 16-bit lanes hold (sum << 6 | word index) keys for two frames per register and
// the running first-extreme is one packed min/max (VIMNMX.U16x2) per register: for equal sums the smaller index wins in
// `darker` (min of key); `lighter` stores 63 - index and takes the max. A pixel is streamed in chunks of 4 frame groups
// (12-16 independent 128-bit loads in flight per thread); the four runs (frame mod 4) of a chunk are merged by
// (sum, frame) and compared strictly with the best so far, so the first extreme wins exactly like the f32 compare of
// src/simple.rs:103-118 (all values are exact integers). The winner's bytes are fetched while its chunk is still in L2.
constexpr int kChunkGroups = 4;
template <int C>
__global__ void __launch_bounds__(256) simple_int_kernel(const __grid_constant__ SimpleArgs a) {
    const long long n_threads = (long long)gridDim.x * blockDim.x;
    const long long tbytes = tile_bytes(C, a.NG);
    const bool darker = a.darker != 0;
    const long long band_stride = (long long)a.NG * (kTilePixels * kUnitBytes);
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < a.n_tiles * kTilePixels; pix += n_threads) {
        const long long tile = pix >> 5;
        const int p = (int)(pix & 31);
        const uint8_t* pbase = a.stack + tile * tbytes + ((long long)a.g0 * kTilePixels + p) * kUnitBytes;
        int best_sum = darker ? 0x7fffffff : -1;
        uint8_t outp[4] = {0, 0, 0, 0};  // src/simple.rs:71-74: the buffer starts at 0
        for (int g_base = 0; g_base < a.n_groups; g_base += kChunkGroups) {
            uint4 u[kChunkGroups][C];
#pragma unroll
            for (int gg = 0; gg < kChunkGroups; gg++) {
#pragma unroll
                for (int c = 0; c < C; c++) {
                    u[gg][c] = make_uint4(0, 0, 0, 0);
                    if (g_base + gg < a.n_groups && ((a.use_mask >> c) & 1u))
                        u[gg][c] = ldg_stream(pbase + c * band_stride + (long long)(g_base + gg) * (kTilePixels * kUnitBytes));
                }
            }
            uint32_t best_e = darker ? 0xffffffffu : 0u, best_o = best_e;  // even frames (0,2) / odd frames (1,3) of each word
#pragma unroll
            for (int gg = 0; gg < kChunkGroups; gg++) {
                uint4 m = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                if (g_base + gg >= a.n_groups) m = make_uint4(0, 0, 0, 0);
                else if (a.wmask) m = __ldg(reinterpret_cast<const uint4*>(a.wmask) + (g_base + gg));
#pragma unroll
                for (int wq = 0; wq < 4; wq++) {
                    const uint32_t mw = wq == 0 ? m.x : (wq == 1 ? m.y : (wq == 2 ? m.z : m.w));
                    uint32_t se = 0, so = 0;  // 16-bit lanes: sums of frames (0,2) and (1,3) of this word
#pragma unroll
                    for (int c = 0; c < C; c++) {
                        const uint32_t xw = wq == 0 ? u[gg][c].x : (wq == 1 ? u[gg][c].y : (wq == 2 ? u[gg][c].z : u[gg][c].w));
                        se += xw & 0x00ff00ffu;               // bands with weight 0 were not loaded (zero)
                        so += __byte_perm(xw, 0u, 0x4341);    // bytes 1 and 3 into the low bytes of the two 16-bit lanes
                    }
                    const uint32_t idx = (uint32_t)(gg * 4 + wq);
                    const uint32_t irep = (darker ? idx : 63u - idx) * 0x00010001u;
                    uint32_t ke = se * 64u + irep, ko = so * 64u + irep;  // (sum << 6) | index in both lanes
                    if (mw != 0xffffffffu) {  // frames outside the window must never win: force their lanes to 0xffff / 0
                        uint32_t me = mw & 0x00ff00ffu, mo = __byte_perm(mw, 0u, 0x4341);
                        me |= me << 8;
                        mo |= mo << 8;
                        if (darker) { ke |= ~me; ko |= ~mo; }
                        else { ke &= me; ko &= mo; }
                    }
                    best_e = darker ? __vminu2(best_e, ke) : __vmaxu2(best_e, ke);
                    best_o = darker ? __vminu2(best_o, ko) : __vmaxu2(best_o, ko);
                }
            }
            // merge the four runs of the chunk: lane k of (best_e: k = 0, 2; best_o: k = 1, 3)
            int c_sum = 0, c_frame = -1;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t key = (((k & 1) ? best_o : best_e) >> ((k >> 1) * 16)) & 0xffffu;
                const bool never = darker ? (key == 0xffffu) : (key == 0u);  // no window frame in this run
                const int sum = (int)(key >> 6);
                int wd = (int)(key & 63u);
                if (!darker) wd = 63 - wd;
                const int frame = (wd << 2) + k;  // relative to the chunk
                const bool better = !never && (c_frame < 0 || (darker ? (sum < c_sum) : (sum > c_sum)) || (sum == c_sum && frame < c_frame));
                if (better) { c_sum = sum; c_frame = frame; }
            }
            if (c_frame >= 0 && (darker ? (c_sum < best_sum) : (c_sum > best_sum))) {  // strict: earlier chunks win ties
                best_sum = c_sum;
                const int f = (a.g0 + g_base) * kGroupFrames + c_frame;
                const PixelSrc src{a.stack + tile * tbytes, a.NG, C, p};
#pragma unroll
                for (int c = 0; c < C; c++) outp[c] = src.at(f, c);
            }
        }
        if (pix < a.n_pixels) {
#pragma unroll
            for (int c = 0; c < C; c++) a.out_image[pix * C + c] = outp[c];
        }
    }
}

// ------------------------------------------------------------------------------------------------ ingest / generator
// src: rows x width x C interleaved bytes (tightly packed band of one frame) -> byte (frame & 15) of each pixel-band unit.
__global__ void __launch_bounds__(256) pack_frame_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ stack,
                                                         long long n_pixels, int C, int NG, int frame) {
    const long long n_threads = (long long)gridDim.x * blockDim.x;
    const int g = frame >> 4, b = frame & 15;
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < n_pixels; pix += n_threads) {
        const long long tile = pix >> 5;
        const int p = (int)(pix & 31);
        for (int c = 0; c < C; c++) stack[unit_offset(tile, C, NG, c, g, p) + b] = src[pix * C + c];
    }
}

// Inverse of pack_frame_kernel (debug / bench helper): frame `frame` of the stack -> interleaved bytes.
__global__ void __launch_bounds__(256) unpack_frame_kernel(const uint8_t* __restrict__ stack, uint8_t* __restrict__ dst,
                                                           long long n_pixels, int C, int NG, int frame) {
    const long long n_threads = (long long)gridDim.x * blockDim.x;
    const int g = frame >> 4, b = frame & 15;
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < n_pixels; pix += n_threads) {
        const long long tile = pix >> 5;
        const int p = (int)(pix & 31);
        for (int c = 0; c < C; c++) dst[pix * C + c] = stack[unit_offset(tile, C, NG, c, g, p) + b];
    }
}

// Same for 16 frames at once: src[k] is frame 16g+k (null = frame absent, byte left zero); writes whole 16-byte units.
struct PackGroupArgs { const uint8_t* src[16]; };
__global__ void __launch_bounds__(256) pack_group_kernel(const PackGroupArgs srcs, uint8_t* __restrict__ stack,
                                                         long long n_pixels, int C, int NG, int g) {
    const long long n_threads = (long long)gridDim.x * blockDim.x;
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < n_pixels; pix += n_threads) {
        const long long tile = pix >> 5;
        const int p = (int)(pix & 31);
        for (int c = 0; c < C; c++) {
            uint32_t wd[4] = {0, 0, 0, 0};
#pragma unroll
            for (int k = 0; k < 16; k++) {
                uint32_t v = srcs.src[k] ? (uint32_t)srcs.src[k][pix * C + c] : 0u;
                wd[k >> 2] |= v << (8 * (k & 3));
            }
            *reinterpret_cast<uint4*>(stack + unit_offset(tile, C, NG, c, g, p)) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        }
    }
}

// One thread per 16-byte unit.
__global__ void __launch_bounds__(256) synth_fill_kernel(uint8_t* __restrict__ stack, long long n_pixels, long long n_tiles, int C,
                                                         int NG, int n_frames, int kind, unsigned long long seed, int width,
                                                         int row0_global, int full_height) {
    const long long n_units = n_tiles * C * NG * kTilePixels;
    const long long n_threads = (long long)gridDim.x * blockDim.x;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += n_threads) {
        const int p = (int)(u & 31);
        long long r = u >> 5;
        const int g = (int)(r % NG); r /= NG;
        const int c = (int)(r % C);
        const long long tile = r / C;
        const long long pix = tile * kTilePixels + p;
        uint32_t wd[4] = {0, 0, 0, 0};
        if (pix < n_pixels) {
            const int y = row0_global + (int)(pix / width), xx = (int)(pix % width);
#pragma unroll 1
            for (int k = 0; k < 16; k++) {
                const int f = g * 16 + k;
                if (f < n_frames) wd[k >> 2] |= (uint32_t)synth_byte(kind, seed, f, n_frames, y, xx, c, width, full_height) << (8 * (k & 3));
            }
        }
        *reinterpret_cast<uint4*>(stack + u * kUnitBytes) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
}

}  // namespace chb
