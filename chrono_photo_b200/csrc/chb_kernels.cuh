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
#include <climits>
#include <type_traits>

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
// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may become resident
// while its predecessor in the stream drains; it must not touch the predecessor's results before this wait (a no-op without
// the attribute).
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// The G lanes that share a pixel are lanes j * (32/G) + pl, j = 0..G-1 (pixel index in the low bits, so a quarter warp
// reads 128 contiguous bytes of a staged row): reductions over j use xor masks 32/G, 2*32/G, .., 16.
template <int G>
__device__ __forceinline__ uint32_t group_sum(uint32_t v) {
#pragma unroll
    for (int m = 32 / G; m < 32; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
template <int G>
__device__ __forceinline__ uint32_t group_or(uint32_t v) {
#pragma unroll
    for (int m = 32 / G; m < 32; m <<= 1) v |= __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}


// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier, used to stage the next pixel-band in shared memory
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr));
    return r;
}
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(saddr));
    return r;
}
__device__ __forceinline__ void sts32(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// per-lane asynchronous 16-byte copies (cp.async, SASS LDGSTS): the G > 1 variants prefetch every lane's own units
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    for (int spin = 0; !ok; spin++) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
        if (spin > (1 << 22)) asm volatile("trap;");  // a lost copy must abort the launch, never hang the GPU
    }
}

// ------------------------------------------------------------------------------------------------ kernel arguments
struct OutlierArgs {
    const uint8_t* stack;
    long long n_pixels, n_tiles;
    int NG, C;
    int g0, n_groups;        // frame groups spanned by the window: [g0, g0 + n_groups)
    unsigned patch_slots;    // != 0: a register slot other than the last holds bytes that are not window frames
    int lead_slots_full;     // 1: slots 0..WPL-2 are real frame groups for every lane (plain loads)
    int window_masked;       // 1: the span holds frames outside the window; they are masked to zero after the load
    const uint32_t* wmask;   // [capacity_groups * 4] byte masks of window frames (0xFF = in window)
    const uint32_t* smask;   // [capacity_groups * 4] byte masks of the --sample subset, or null (GENERIC kernels only)
    const int32_t* win_frames;  // [n] window position -> frame index
    int n, n_sub;            // window length ("samples"), subsample size
    int first_frame;         // (frame of window position 0) & 15: its byte inside the first group
    int rk[6];               // 0-based ranks inside the subsample: q1 lo/hi, median lo/hi, q3 lo/hi
    float q1_frac, q3_frac;  // interpolation weights of quantile() (src/chrono.rs:568-579)
    float inv_n_sub;
    int absolute;
    int int_dist;            // 1: absolute thresholds and every weight 0 or 1: 4 * dist_sq is an exact integer (IntDist)
    int thr4;                // ceil(4 * thr_sq): dist_sq >= thr_sq <=> 4 * dist_sq >= thr4
    float thr_min, thr_max, thr_scale, thr_sq;
    float w[4];
    int bg, om;
    FadeDev fade;
    int frame_offset;
    int exact_quartiles;     // 1: the fast tier computes exact quartiles (debug planes requested) instead of an IQR bound
    int contig_f0;           // >= 0: the window is the contiguous frame range starting here (position s = frame contig_f0 + s)
    int mask_path;           // 1: integer distance, contiguous window of at most kMaskGroups groups: dense_masks_int + finish_masks
    int inline_min;          // > 0 (G == 1 kernels): a tile with at least this many uncertified pixels is finished inside the streaming kernel
    int hard_inline_min;     // > 0 (G == 1 kernels): same for a warp-full of the iterative tier, finished inside outlier_hard_kernel
    int hist_all;            // 1: outlier_hist_kernel takes every pixel of the band (series beyond the register-resident variants), not a queue
    int hard_window;         // 1 (G == 1 kernels, absolute thresholds): the iterative tier tries the two straight-line windows before the solver
    int hard_drains_all;     // 1: outlier_hard_kernel also finishes the pixels the streaming kernel queued itself (dense pass): no outlier_exact_kernel launch
    unsigned long long seed, pixel_offset;
    unsigned long long block_pixels, block_skip;  // interleaved row-block shards (see chrono_b200.h); 0 / 0: one contiguous band
    uint8_t* out_image;
    uint8_t* out_mask;  // may be null
    unsigned long long* counters;  // [0] warnings (all-outlier pixels), [1] pixels on the exact path, [2] pixels on the iterative (hard) path
    // exact-path queue of the launch, one slot per pixel of the band (outlier_exact_kernel drains it). Slot i < ghq_count mirrors
    // entry i of the iterative tier's queue (pix = -1: that pixel was certified after all), so the exact path sees neighbouring
    // pixels -- whose objects dwell on them during the same frames -- side by side; pixels the streaming kernel queues
    // directly fill the array from its far end (gq_count of them).
    struct QueueEntry* gq;
    unsigned int* gq_count;
    long long* ghq;                // pixels for the iterative tier in tile order (compact_hard_kernel builds it from hflags)
    unsigned int* ghq_count;
    uint32_t* hflags;              // [n_tiles] bit p: pixel p of the tile goes to the iterative tier (zeroed by the host)
    float* dbg_median; float* dbg_q1; float* dbg_q3; int* dbg_nout;
};

// Global index of local pixel `pix`: what keys the per-pixel random draws, so that a shard draws what the whole image would.
__device__ __forceinline__ unsigned long long pixel_gid(const OutlierArgs& a, long long pix) {
    unsigned long long g = a.pixel_offset + (unsigned long long)pix;
    if (a.block_pixels) g += ((unsigned long long)pix / a.block_pixels) * a.block_skip;
    return g;
}

// ------------------------------------------------------------------------------------------------ exact order statistics
// A pixel-band's samples sit in registers as packed bytes (4 frames per word), split over G lanes. For a candidate
// value c, F(c) = sum |x - c| costs one VABSDIFF4.ACC per word; F(c+1) - F(c) = 2*#{x <= c} - CAP gives an exact count,
// and the k-th smallest value is min{c : #{x <= c} >= k+1}. Bytes that are not part of the sample are zero, which
// shifts every rank by the (known) number of such bytes.
//
// Every iteration evaluates F at two points per lane (four accumulator chains over the same registers). A rank starts
// with a "jump" at a guess g (F(g), F(g+1) -> #{x <= g}), then walks away from g two values per iteration, re-using the
// F value at the edge of the known range so that two new evaluations give two new counts; after three walking steps it
// bisects. All ranks a band needs (median pair; quartile pairs for relative thresholds) run through ONE warp-synchronous
// loop as a per-lane state machine, so lanes that finish a rank early move on to their next rank instead of idling.
template <int W4, int G>
struct Sel {
    static __device__ __forceinline__ void eval2(const uint32_t (&x)[W4], int e1, int e2, uint32_t& f1, uint32_t& f2) {
        const uint32_t c1 = rep4(e1), c2 = rep4(e2);
        uint32_t a1 = 0, a2 = 0, b1 = 0, b2 = 0;
#pragma unroll
        for (int q = 0; q < W4; q += 2) {
            a1 = sad4_acc(x[q], c1, a1);
            a2 = sad4_acc(x[q], c2, a2);
            b1 = sad4_acc(x[q + 1], c1, b1);
            b2 = sad4_acc(x[q + 1], c2, b2);
        }
        f1 = group_sum<G>(a1 + b1);
        f2 = group_sum<G>(a2 + b2);
    }
};

struct BandRanks {  // results of one pixel-band
    int mlo, mhi, q1a, q1b, q3a, q3b;
};

// Solves the median pair (and the two quartile pairs when rel) of one pixel-band; padded ranks are a.rk[] + pad.
// The per-iteration bookkeeping is written with selects instead of branches: every lane runs the same short
// instruction sequence whatever phase it is in; only "a rank has been resolved" takes a (divergent) branch.
template <int W4, int G>
__device__ __forceinline__ void band_solve(const uint32_t (&x)[W4], uint32_t ssum, float inv_cnt, const OutlierArgs& a, int pad, int cap, BandRanks& r) {
    const bool rel = !a.absolute;
    const int nt = rel ? 6 : 2;  // targets in processing order: m1, m2, q1a, q1b, q3a, q3b
    // per-lane state. mode 0: jump at g; 1: walking in direction d from edge b with fc = F(b); 2: bisecting [b, hi].
    // mode 2 keeps the counts at both ends of its bracket: cntl = #{x <= b - 1}, cnth = #{x <= hi}
    int t = 0, kp = a.rk[2] + pad, mode = 0, steps = 0, d = 1, b = 0, hi = 255, cnth = cap, cntl = 0;
    uint32_t fc = 0;
    int g = __float2int_rn((float)ssum * inv_cnt);  // mean as the first guess for the median
    g = g < 0 ? 0 : (g > 254 ? 254 : g);
    int cv_mlo = cap, dq = 0;
    bool rejumped = false;  // the first rank's guess has been moved to the mean of the samples on the answer's side
    r.mlo = r.mhi = r.q1a = r.q1b = r.q3a = r.q3b = 0;
    while (__any_sync(0xffffffffu, t < nt)) {
        const bool act = t < nt;
        const bool m0 = mode == 0, m1 = mode == 1;
        const bool u = d > 0;
        // ---- the two evaluation points of this lane
        int e2w = b + 2 * d;
        const bool e2_out = (e2w < 0) | (e2w > 255);
        e2w = e2w < 0 ? 0 : (e2w > 255 ? 255 : e2w);
        // mode 2: the bracket is cut where a uniform spread of its cnth - cntl samples over its values puts the rank (iid bytes:
        // one or two steps instead of log2(width)); every third cut is a plain bisection, which bounds the worst case
        int mid = (b + hi) >> 1;
        if (mode == 2 && steps % 3 != 2) {
            const float est = __fdividef((float)(kp + 1 - cntl) * (float)(hi - b + 1), (float)(cnth - cntl));
            const int ei = b - 1 + __float2int_rn(est);
            mid = ei < b ? b : (ei > hi - 1 ? hi - 1 : ei);
        }
        int e1 = m0 ? g : (m1 ? b + d : mid);
        int e2 = m0 ? g + 1 : (m1 ? e2w : mid + 1);
        e1 = act ? e1 : 0;
        e2 = act ? e2 : 0;
        uint32_t f1, f2;
        Sel<W4, G>::eval2(x, e1, e2, f1, f2);
        // ---- counts
        const int nj = ((int)f2 - (int)f1 + cap) >> 1;                         // jump / bisect: #{x <= e1}
        const int n1 = ((u ? (int)f1 - (int)fc : (int)fc - (int)f1) + cap) >> 1;  // walk: #{x <= b} (up) / #{x <= b-1} (down)
        int n2 = ((u ? (int)f2 - (int)f1 : (int)f1 - (int)f2) + cap) >> 1;       // walk: #{x <= b+1} (up) / #{x <= b-2} (down)
        n2 = e2_out ? (u ? cap : 0) : n2;
        if (rel && act && m0 && t == 0) {  // spread estimate for the quartile guesses: mean absolute deviation around g
            const float mad = ((float)f1 - (float)pad * (float)g) * a.inv_n_sub;
            dq = __float2int_rn(0.95f * mad);
        }
        // ---- mode 0: jump.  up0: the answer is above g
        const bool up0 = kp >= nj;
        // ---- mode 1: walk.  rb: resolved at b, rb1: resolved at b + d
        const bool s1 = n1 >= kp + 1, s2 = n2 >= kp + 1;
        const bool rb = (s1 == u), rb1 = !rb && (s2 == u);
        // ---- mode 2: bisect
        const bool ge = nj >= kp + 1;
        // ---- new state
        int nb, nhi = hi, ncnth = cnth, ncntl = cntl, nmode = mode, nsteps = steps, nd = d;
        uint32_t nfc = fc;
        bool res;
        int v, cv;
        uint32_t fnext = 0;
        bool has_next = false;
        if (m0) {
            nd = up0 ? 1 : -1;
            nb = up0 ? g + 1 : g;
            nfc = up0 ? f2 : f1;
            ncnth = up0 ? cap : nj;
            nmode = 1; nsteps = 0;
            res = up0 ? (nb == 255) : (nb == 0);
            v = nb; cv = ncnth;
            // second jump of the first rank (marked by steps == 7; hi / cntl, unused outside mode 2, carry the first jump's value
            // and count): when the two jumps straddle the answer they ARE a bracket with both counts known -- iid bytes, whose
            // side mean lies far beyond the answer, go on by interpolation inside it instead of walking back and bisecting
            if (!res && t == 0 && steps == 7) {
                const int g0 = hi, n0 = cntl;
                if (g > g0 && !up0) { nmode = 2; nb = g0 + 1; nhi = g; ncntl = n0; ncnth = nj; }
                else if (g < g0 && up0) { nmode = 2; nb = g + 1; nhi = g0; ncntl = nj; ncnth = n0; }
                if (nmode == 2) { res = nb >= nhi; v = nhi; cv = ncnth; }
            }
            if (!res && t == 0 && !rejumped) {
                // A pixel reaches this solver because its mean is no guess for its median: an object rests on it for a good part
                // of the series. The samples on the median's side of g are mostly background, and F(g), F(g+1) and the sum give
                // their mean for free (sum of (x - g - 1) over x > g is (F(g+1) + sum - cap (g+1)) / 2; likewise below g, without
                // the pad zeros): jump there once before walking.
                rejumped = true;
                int g2;
                if (up0) {
                    const int above = cap - nj;  // > 0: the rank lies above g
                    const int A = ((int)f2 + (int)ssum - cap * (g + 1)) >> 1;
                    g2 = g + 1 + __float2int_rn(__fdividef((float)A, (float)above));
                } else {
                    const int below = nj - pad;  // > 0: real samples <= g
                    const int B = (((int)f1 - (int)ssum + cap * g) >> 1) - pad * g;
                    g2 = g - __float2int_rn(__fdividef((float)B, (float)max(below, 1)));
                }
                // ... unless the samples on that side spread far beyond the answer (iid bytes, wide unimodal noise: the side mean is
                // tens of values away while only a few samples separate g from the rank). Then the number of samples still to be
                // passed over the density the mean absolute deviation implies -- n / (3.5 MAD) per value, for uniform and Gaussian
                // spreads alike -- is the better estimate; it is taken when it is less than a quarter of the way to the side mean
                // (two clusters give a ratio of 7 (1 - p)(1/2 - p) for an object share p: below 1/4 only for p > 0.43).
                {
                    const int far = up0 ? kp + 1 - nj : nj - kp;
                    const float madn = ((float)(up0 ? f2 : f1) - (float)pad * (float)(up0 ? g + 1 : g)) * a.inv_n_sub;
                    const int step = __float2int_rn((float)far * 3.5f * madn * a.inv_n_sub);
                    const int dside = up0 ? g2 - g : g - g2;
                    if (4 * step < dside) g2 = up0 ? g + step : g - step;
                }
                g2 = g2 < 0 ? 0 : (g2 > 254 ? 254 : g2);
                if (g2 > g + 3 || g2 < g - 3) { nhi = g; ncntl = nj; g = g2; nmode = 0; nsteps = 7; }  // close guesses just walk
            }
        } else if (m1) {
            res = rb | rb1;
            v = rb ? b : b + d;
            cv = rb ? (u ? n1 : cnth) : (u ? n2 : n1);
            fnext = rb ? f1 : (u ? f2 : fc);
            has_next = rb ? u : (u ? (b + 2 <= 255) : true);
            nb = b + 2 * d;
            nfc = f2;
            ncnth = u ? cnth : n2;
            nsteps = steps + 1;
            if (!res) {
                if (u ? (nb >= 255) : (nb <= 0)) {  // walked into the end of the byte range
                    res = true; v = u ? 255 : 0; cv = ncnth; has_next = false;
                } else if (nsteps >= 3) {  // far from the guess: bracket what is left
                    nmode = 2;
                    nsteps = 0;
                    nhi = u ? 255 : nb;
                    ncntl = u ? n2 : pad;  // going down the bracket starts at 0, where the pad zeros sit
                    nb = u ? nb : 0;
                    ncnth = u ? cap : ncnth;
                }
            }
        } else {
            nhi = ge ? e1 : hi;
            ncnth = ge ? nj : cnth;
            ncntl = ge ? cntl : nj;
            nb = ge ? b : e1 + 1;
            nsteps = steps + 1;
            res = nb >= nhi;
            v = nhi; cv = ncnth;
        }
        if (act) {
            b = nb; hi = nhi; cnth = ncnth; cntl = ncntl; mode = nmode; steps = nsteps; d = nd; fc = nfc;
            // ---- resolved: store, then set up the next rank(s) of this lane
            if (res) {
#pragma unroll 1
                for (;;) {
                    if (t == 0) { r.mlo = v; cv_mlo = cv; }
                    else if (t == 1) r.mhi = v;
                    else if (t == 2) r.q1a = v;
                    else if (t == 3) r.q1b = v;
                    else if (t == 4) r.q3a = v;
                    else r.q3b = v;
                    t++;
                    if (t >= nt) break;
                    const int kprev = kp;
                    kp = a.rk[t == 1 ? 3 : (t == 2 ? 0 : (t == 3 ? 1 : (t == 4 ? 4 : 5)))] + pad;
                    if (t & 1) {  // second rank of a pair: the same value unless fewer than kp+1 samples are <= v
                        if (kp == kprev || cv >= kp + 1) continue;
                        // answer >= v+1: walk up from there (v < 255 because cv < cap)
                        if (has_next) {
                            if (v + 1 == 255) { v = 255; cv = cap; has_next = false; continue; }
                            mode = 1; d = 1; b = v + 1; fc = fnext; steps = 0; cnth = cap;
                        } else {
                            mode = 0; g = v + 1 > 254 ? 254 : v + 1;
                        }
                    } else {  // first rank of a quartile pair: jump at median -/+ spread
                        g = (t == 2) ? r.mlo - dq : r.mhi + dq;
                        g = g < 0 ? 0 : (g > 254 ? 254 : g);
                        mode = 0;
                    }
                    break;
                }
            }
        }
    }
}

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
    // makes frame group g the current one (its units in u[])
    __device__ __forceinline__ void enter_group(int g) {
        if (g != cur_g) {
            cur_g = g;
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (c < C) u[c] = __ldg(reinterpret_cast<const uint4*>(base + c * band_stride + (long long)g * (kTilePixels * kUnitBytes)));
        }
    }
    // word q (a compile-time constant) of the current group, every band
    template <int Q>
    __device__ __forceinline__ void group_word(uint32_t (&xw)[4]) const {
#pragma unroll
        for (int c = 0; c < 4; c++) xw[c] = (c < C) ? (Q == 0 ? u[c].x : (Q == 1 ? u[c].y : (Q == 2 ? u[c].z : u[c].w))) : 0u;
    }
    // the 4-frame word (frames frame .. frame+3, frame a multiple of 4) of every band
    __device__ __forceinline__ void fetch_word(int frame, uint32_t (&xw)[4]) {
        const int g = frame >> 4;
        if (g != cur_g) {
            cur_g = g;
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (c < C) u[c] = __ldg(reinterpret_cast<const uint4*>(base + c * band_stride + (long long)g * (kTilePixels * kUnitBytes)));
        }
        const int wsel = (frame >> 2) & 3;
#pragma unroll
        for (int c = 0; c < 4; c++) xw[c] = (c < C) ? (wsel == 0 ? u[c].x : (wsel == 1 ? u[c].y : (wsel == 2 ? u[c].z : u[c].w))) : 0u;
    }
};

// Conservative per-word pre-test of the exact path: a frame can only be an outlier if some band deviates from the
// band's centre by more than cap_c, where the caps split thr_min^2 evenly over the positive-weight bands
// ((fac_c * (cap_c + halfwidth_c))^2 <= thr^2 / bands, with margin). One VABSDIFF4 + three logic ops per band test four
// frames at once; words without a flagged frame are skipped. Flags: bit 7 of byte k <=> frame k may be an outlier.
struct WordScan {
    uint32_t cc[4], kadd[4];
    // "certainly an outlier": a deviation from the band's centre from which this band ALONE reaches the threshold -- every
    // other band's term is non-negative when no weight is negative, and f32 rounding is monotone, so the reference's sum
    // does too (checked below with the reference's own f32 operations). kadd_hi: 128 - that deviation; m_hi: 0 = test off.
    uint32_t kadd_hi[4], m_hi[4];
    bool use[4];
    bool all;  // caps are useless (tiny threshold, NaN): every frame is evaluated
};
__device__ __forceinline__ void make_word_scan(const OutlierArgs& a, const float (&median)[4], const float (&iqr_inv)[4], WordScan& ws) {
    int nb = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float w = a.w[i];
        ws.use[i] = (i < a.C) && (w != 0.0f) && !(w < 0.0f);  // negative weights only lower dist_sq
        nb += ws.use[i] ? 1 : 0;
    }
    ws.all = false;
    bool no_negative = a.thr_sq > 0.0f;  // (false for NaN as well)
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (i < a.C) no_negative = no_negative && (a.w[i] == 0.0f || a.w[i] > 0.0f);
    const float share = nb > 0 ? sqrtf(a.thr_sq * 0.9999f / (float)nb) : 0.0f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        ws.cc[i] = 0; ws.kadd[i] = 0; ws.kadd_hi[i] = 0; ws.m_hi[i] = 0;
        if (ws.use[i]) {
            const float facs = a.absolute ? a.w[i] : a.w[i] * iqr_inv[i];  // the factor of src/chrono.rs:265-278, as make_dist_ctx forms it
            if (no_negative) {
                const float need = sqrtf(a.thr_sq) / facs + 0.5f;  // |x - centre| >= need => |median - x| >= need - 1/2 => (fac * diff)^2 >= thr^2
                const int ch = (need == need && need > 0.0f && need < 125.0f) ? (int)ceilf(need) + 1 : 128;
                if (ch <= 127) {
                    const float t = facs * ((float)ch - 0.5f);
                    if (t * t >= a.thr_sq) { ws.kadd_hi[i] = rep4(128 - ch); ws.m_hi[i] = 0x80808080u; }
                }
            }
            const float fac = fabsf(facs);
            const int center = (int)median[i];  // floor: medians are >= 0
            const float halfw = median[i] - (float)center;
            const float capf = share / fac - halfw - 1e-3f;
            int cap = (capf == capf && capf > -1.0f) ? (int)floorf(capf) : -1;  // NaN / negative -> no cap
            if (cap < 0) ws.all = true;
            cap = cap < 0 ? 0 : (cap > 127 ? 127 : cap);
            ws.cc[i] = rep4(center);
            ws.kadd[i] = rep4(127 - cap);  // (d & 0x7f) + (127 - cap) sets bit 7 iff (d & 0x7f) > cap
        }
    }
    if (!(a.thr_sq > 0.0f)) ws.all = true;  // threshold 0 (or NaN): everything is an outlier
}
__device__ __forceinline__ uint32_t may_exceed(const WordScan& ws, const uint32_t (&xw)[4]) {
    if (ws.all) return 0x80808080u;
    uint32_t ex = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (ws.use[i]) {
            const uint32_t d = absdiff4(xw[i], ws.cc[i]);
            ex |= d | ((d & 0x7f7f7f7fu) + ws.kadd[i]);  // bit 7: d >= 128, or low 7 bits above the cap
        }
    }
    return ex & 0x80808080u;
}
// Same, plus the frames that certainly are outliers (flags in `sure`, a subset of the returned flags)
__device__ __forceinline__ uint32_t may_exceed_sure(const WordScan& ws, const uint32_t (&xw)[4], uint32_t& sure) {
    sure = 0;
    if (ws.all) return 0x80808080u;
    uint32_t ex = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (ws.use[i]) {
            const uint32_t d = absdiff4(xw[i], ws.cc[i]), lo7 = d & 0x7f7f7f7fu;
            ex |= d | (lo7 + ws.kadd[i]);
            sure |= (d | (lo7 + ws.kadd_hi[i])) & ws.m_hi[i];  // bit 7: d >= 128, or low 7 bits at or above the band's own reach
        }
    }
    return (ex | sure) & 0x80808080u;
}

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

// Integer form of the distance for absolute thresholds with weights in {0, 1} (the CLI default): median = (lo + hi) / 2 with
// lo = floor(median), hi = ceil(median) (no sample lies strictly between them), so |2 * median - 2 * x| = |x - lo| + |x - hi|
// and 4 * dist_sq = sum over bands of (L + H)^2 = L.L + 2 L.H + H.H -- three IDP.4A on the per-frame byte vectors L, H
// (VABSDIFF4 per band and word, transposed with PRMT). Every f32 operation of the reference (src/chrono.rs:265-278) is exact
// on these values (multiples of 0.25 below 2^18), so the integers reproduce it bit for bit.
struct IntDist {
    uint32_t lo[4], hi[4];
    bool use[4];
};
__device__ __forceinline__ void make_int_dist(const OutlierArgs& a, const float (&median)[4], IntDist& d) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        d.use[i] = (i < a.C) && (a.w[i] != 0.0f);
        const int lo = (int)median[i];  // medians are >= 0: truncation is floor
        d.lo[i] = rep4(lo);
        d.hi[i] = rep4(lo + (median[i] != (float)lo ? 1 : 0));
    }
}
// 4 * dist_sq of the four frames of a word (frame k in out[k])
__device__ __forceinline__ void int_dist4(const IntDist& d, const uint32_t (&xw)[4], uint32_t (&out)[4]) {
    uint32_t L[4], H[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        L[i] = d.use[i] ? absdiff4(xw[i], d.lo[i]) : 0u;
        H[i] = d.use[i] ? absdiff4(xw[i], d.hi[i]) : 0u;
    }
    // 4 x 4 byte transposes: band-major words -> one word per frame
    const uint32_t l01a = __byte_perm(L[0], L[1], 0x5140), l01b = __byte_perm(L[0], L[1], 0x7362);
    const uint32_t l23a = __byte_perm(L[2], L[3], 0x5140), l23b = __byte_perm(L[2], L[3], 0x7362);
    const uint32_t h01a = __byte_perm(H[0], H[1], 0x5140), h01b = __byte_perm(H[0], H[1], 0x7362);
    const uint32_t h23a = __byte_perm(H[2], H[3], 0x5140), h23b = __byte_perm(H[2], H[3], 0x7362);
    const uint32_t Lf[4] = {__byte_perm(l01a, l23a, 0x5410), __byte_perm(l01a, l23a, 0x7632), __byte_perm(l01b, l23b, 0x5410), __byte_perm(l01b, l23b, 0x7632)};
    const uint32_t Hf[4] = {__byte_perm(h01a, h23a, 0x5410), __byte_perm(h01a, h23a, 0x7632), __byte_perm(h01b, h23b, 0x5410), __byte_perm(h01b, h23b, 0x7632)};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t sq = __dp4a(Hf[k], Hf[k], __dp4a(Lf[k], Lf[k], 0u));
        out[k] = sq + 2u * __dp4a(Lf[k], Hf[k], 0u);
    }
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

// ---- dense per-frame pass (src/chrono.rs:261-288) for the integer distance: EVERY frame of a contiguous window is
// classified, four frames per instruction group and without a data-dependent branch, at about a dozen instructions per
// pixel-frame -- the path for pixels (noise at the threshold, iid bytes) where most 4-frame words hold a candidate, so
// that a pre-test only adds work. A lane walks its pixel's 16-frame units group by group (three or four 128-bit loads,
// the next group's in flight while the current one is evaluated). Per 4-frame word: the band-major words are transposed
// to one word per frame (8 PRMT), L = |x - lo|, H = |x - hi| bytewise (2 VABSDIFF4 per frame) and
// 4 dist_sq = L.L + H.H + 2 L.H (3 IDP.4A). Per frame: one compare sets the frame's bit in the group's 16-bit outlier mask
// and a key 16 * (4 dist_sq) + (15 - j) tracks the group's first maximum. Per group: the masks give the outlier count, the
// first non-outlier and the first / last outlier; the group's key is folded into (4 dist_sq << 12 | 4095 - position), whose
// maximum over the window is the reference's first strict maximum. Outlier sums (--outlier average, --background average)
// are taken under a branch in frame order (AVG).
struct DensePass {
    int k, first_idx, last_idx, first_non;
    uint32_t maxkey;
    float out_sum[4], mean_dist;
};
template <int C, bool AVG>
__device__ __forceinline__ void dense_pass_int(const OutlierArgs& a, const uint8_t* colbase, long long band_stride, int f0, int n,
                                               const float (&median)[4], DensePass& r) {
    uint32_t lof = 0, hif = 0;  // frame-major constants: byte c = floor / ceil of band c's median (0 for a band of weight 0)
    bool use[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        use[c] = (c < C) && (a.w[c] != 0.0f);
        const int lo = (int)median[c];
        const int hi = lo + (median[c] != (float)lo ? 1 : 0);
        if (use[c]) { lof |= (uint32_t)lo << (8 * c); hif |= (uint32_t)hi << (8 * c); }
    }
    const int thr4 = a.thr4;
    const uint32_t negthr = (uint32_t)(-thr4);
    r.k = 0; r.first_idx = 0x7fffffff; r.last_idx = -1; r.first_non = 0x7fffffff; r.maxkey = 0; r.mean_dist = 0.0f;
#pragma unroll
    for (int c = 0; c < 4; c++) r.out_sum[c] = 0.0f;
    const int gA = f0 >> 4, gB = (f0 + n - 1) >> 4;
    uint4 cur[4], nxt[4];
    auto load = [&](uint4 (&u)[4], int g) {
#pragma unroll
        for (int c = 0; c < 4; c++)
            u[c] = (use[c] || (AVG && c < C)) ? __ldg(reinterpret_cast<const uint4*>(colbase + c * band_stride + (long long)g * (kTilePixels * kUnitBytes))) : make_uint4(0, 0, 0, 0);
    };
    // AVG sums the samples of every band, weight 0 or not: such a band is loaded and masked out of the distance
    const uint32_t usemask = (use[0] ? 0xffu : 0u) | (use[1] ? 0xff00u : 0u) | (use[2] ? 0xff0000u : 0u) | (use[3] ? 0xff000000u : 0u);
    load(cur, gA);
#pragma unroll 1
    for (int g = gA; g <= gB; g++) {
        if (g < gB) load(nxt, g + 1);
        const int sg = 16 * g - f0;  // window position of the group's frame 0
        const bool full = (sg >= 0) && (sg + 16 <= n);
        uint32_t om16 = 0;
        int gkey = INT_MIN;  // max over the group's frames of 16 * (4 dist_sq - thr4) + (15 - j)
        // FULL: every frame of the group lies in the window (all groups but the first and the last of a window); the bits of
        // om16 are then collected by a funnel shift of the sign of (thr4 - 1 - d4), frame j at bit 15 - j, and reversed below
        auto group_body = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t xw[4];
#pragma unroll
                for (int c = 0; c < 4; c++) xw[c] = q == 0 ? cur[c].x : (q == 1 ? cur[c].y : (q == 2 ? cur[c].z : cur[c].w));
                const uint32_t t0 = __byte_perm(xw[0], xw[1], 0x5140), t1 = __byte_perm(xw[0], xw[1], 0x7362);
                const uint32_t t2 = __byte_perm(xw[2], xw[3], 0x5140), t3 = __byte_perm(xw[2], xw[3], 0x7362);
                const uint32_t pf[4] = {__byte_perm(t0, t2, 0x5410), __byte_perm(t0, t2, 0x7632), __byte_perm(t1, t3, 0x5410), __byte_perm(t1, t3, 0x7632)};
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    const int j = 4 * q + kk;
                    const uint32_t pfm = AVG ? (pf[kk] & usemask) : pf[kk];
                    const uint32_t L = absdiff4(pfm, lof), H = absdiff4(pfm, hif);
                    const uint32_t lh = __dp4a(L, H, 0u);
                    // dm = 4 dist_sq - thr4 (the subtraction rides in the first accumulator): negative <=> not an outlier
                    const int dm = (int)(__dp4a(H, H, __dp4a(L, L, negthr)) + lh + lh);
                    if (FULL || (sg + j >= 0 && sg + j < n)) {  // uniform
                        if (FULL && !AVG) {
                            om16 = __funnelshift_l((uint32_t)dm, om16, 1);  // collects the NON-outlier bits, frame j at bit 15 - j
                        } else {
                            const bool isout = dm >= 0;
                            om16 |= isout ? (1u << j) : 0u;
                            if (AVG && isout) {  // frame order: the reference's f32 sum of sqrt(dist_sq) is order-dependent
#pragma unroll
                                for (int c = 0; c < 4; c++) r.out_sum[c] += (float)((pf[kk] >> (8 * c)) & 0xffu);
                                r.mean_dist += sqrtf(0.25f * (float)(dm + thr4));
                            }
                        }
                        gkey = max(gkey, dm * 16 + (15 - j));
                    }
                }
            }
            if (FULL && !AVG) om16 = (~__brev(om16)) >> 16;
        };
        if (full) group_body(std::true_type{});
        else group_body(std::false_type{});
        // ---- fold the group
        uint32_t valid = 0xffffu;
        if (!full) {
            const int lo_j = sg < 0 ? -sg : 0, hi_j = (n - sg < 16) ? n - sg : 16;  // valid frames j in [lo_j, hi_j)
            valid = ((1u << hi_j) - 1u) & ~((1u << lo_j) - 1u);
        }
        const uint32_t non = ~om16 & valid;
        r.k += __popc(om16);
        if (om16) {
            r.first_idx = min(r.first_idx, sg + __ffs(om16) - 1);
            r.last_idx = sg + 31 - __clz(om16);
        }
        if (non) r.first_non = min(r.first_non, sg + __ffs(non) - 1);
        {
            const int s = sg + 15 - (gkey & 15);
            const uint32_t cand = ((uint32_t)((gkey >> 4) + thr4) << 12) | (uint32_t)(4095 - s);
            // (a partial group without a valid frame cannot occur: gA and gB both hold window frames)
            r.maxkey = max(r.maxkey, cand);
        }
#pragma unroll
        for (int c = 0; c < 4; c++) cur[c] = nxt[c];
    }
}

// ---- dense pass with per-group outlier masks (windows of at most kMaskGroups frame groups) -------------------------
// Same classification as dense_pass_int, but the 16-bit outlier mask of every group is parked in a shared-memory slot of
// the thread: the masks ARE the reference's outlier list (src/chrono.rs:280-287), so every background / outlier policy can
// be finished by walking set bits (finish_masks) -- no second pass over the frames, whatever the policy. Two groups are
// kept in registers, the reload of a buffer is issued as soon as the buffer has been evaluated.
constexpr int kMaskGroups = 16;  // 256 frames: the G == 1 variants of K1 hold at most 13 groups
__host__ __device__ constexpr int dense_ring_depth(int C) { return C == 3 ? 3 : 2; }  // frame groups of a pixel in flight (one being evaluated)
constexpr uint32_t kRingStride = 256 * 16;  // every kernel that runs the dense pass has 256 threads: one 16-byte unit each per row
struct MaskSlots {
    uint32_t base;    // shared-memory address of this thread's first halfword
    uint32_t stride;  // bytes between the halfwords of consecutive groups
    // the dense pass's ring of frame groups: row (slot * C + band) holds one 16-byte unit per thread (filled by cp.async)
    uint32_t ring;         // shared-memory address of this thread's unit in row 0; rows are kRingStride bytes apart
    __device__ __forceinline__ void put(int gi, uint32_t m) const { asm volatile("st.shared.u16 [%0], %1;" ::"r"(base + gi * stride), "h"((unsigned short)m) : "memory"); }
    __device__ __forceinline__ uint32_t get(int gi) const {
        unsigned short v;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(base + gi * stride));
        return (uint32_t)v;
    }
};
struct IntMedians {  // frame-major constants of the integer distance: byte c = floor / ceil of band c's median (0: weight 0)
    uint32_t lof, hif;
    // one-sided form used by the dense pass. With lo = floor(median), hi = lo + d (d = 0 or 1), L = |x - lo|:
    //   |x - lo| + |x - hi| = 2 L + s d with s = +1 for x <= lo, -1 for x >= hi, and L s = lo - x on both sides, hence
    //   4 dist_sq = sum_c (2 L + s d)^2 = 4 [ sum L^2 + sum d lo - sum d x ] + sum d
    // -- one VABSDIFF4 and two IDP.4A per frame instead of two and three. negd: byte c = -d_c (signed operand of the
    // second IDP.4A); k2 = sum d; t = ceil((thr4 - k2) / 4), so that 4 dist_sq >= thr4 <=> q >= t for the bracket q;
    // c0 = sum d lo - t rides in the first accumulator: the pass works on q - t (negative <=> not an outlier).
    uint32_t negd;
    int c0, t, k2;
};
template <int C>
__device__ __forceinline__ IntMedians make_int_medians(const OutlierArgs& a, const float (&median)[4]) {
    IntMedians m{0u, 0u, 0u, 0, 0, 0};
    int k1 = 0;
#pragma unroll
    for (int c = 0; c < C; c++) {
        if (a.w[c] != 0.0f) {
            const int lo = (int)median[c];  // medians are >= 0: truncation is floor
            const int d = median[c] != (float)lo ? 1 : 0;
            m.lof |= (uint32_t)lo << (8 * c);
            m.hif |= (uint32_t)(lo + d) << (8 * c);
            m.negd |= d ? (0xffu << (8 * c)) : 0u;
            k1 += d * lo;
            m.k2 += d;
        }
    }
    m.t = (a.thr4 - m.k2 + 3) >> 2;  // ceil((thr4 - k2) / 4), also for negative numerators
    m.c0 = k1 - m.t;
    return m;
}
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b_signed, int c) {  // sum of a.u8[i] * b.s8[i] + c
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b_signed), "r"(c));
    return d;
}
// 4 * dist_sq of one frame given as a frame-major word (byte c = band c; bands of weight 0 must be zero in x)
__device__ __forceinline__ uint32_t int_dist_frame(const IntMedians& m, uint32_t x) {
    const uint32_t L = absdiff4(x, m.lof), H = absdiff4(x, m.hif);
    const uint32_t lh = __dp4a(L, H, 0u);
    return __dp4a(H, H, __dp4a(L, L, 0u)) + lh + lh;
}
template <int C>
__device__ __forceinline__ void dense_masks_int(const OutlierArgs& a, const uint8_t* colbase, long long band_stride, int f0, int n,
                                                const IntMedians& im, const MaskSlots ms, int& k_out, uint32_t& maxkey_out) {
    const uint32_t lof = im.lof, negd = im.negd;
    const int c0 = im.c0;
    bool all_use = true;
#pragma unroll
    for (int c = 0; c < C; c++) all_use = all_use && (a.w[c] != 0.0f);
    const int gA = f0 >> 4, gB = (f0 + n - 1) >> 4;
    int k = 0;
    uint32_t maxkey = 0;
    // one pointer per band and one running offset (in 16-byte units: the next frame group of the tile is kTilePixels units on)
    const uint4* pb[C];
#pragma unroll
    for (int c = 0; c < C; c++) pb[c] = reinterpret_cast<const uint4*>(colbase + c * band_stride) + (long long)gA * kTilePixels;
    // The pixel's units travel global -> shared memory by cp.async (LDGSTS: no register is tied up while a load is in flight)
    // into a ring of D frame groups per thread, D - 1 of them ahead of the group being evaluated: the L2 latency of a group is
    // covered by the evaluation of D - 1 others, where a register double buffer covered one and paid twelve moves per group.
    // Every thread reads back only the 16 bytes it copied itself, so cp.async.wait_group is all the synchronisation needed.
    constexpr int D = dense_ring_depth(C);
    constexpr uint32_t kSlotBytes = C * kRingStride;
    const int last_off = (gB - gA) * kTilePixels;
    auto issue = [&](uint32_t saddr, int off) {  // (a group beyond the window's last one re-reads the last one: no branch, never used)
        const int o = min(off, last_off);
#pragma unroll
        for (int c = 0; c < C; c++) cp_async16(saddr + c * kRingStride, pb[c] + o);
    };
    // ONE copy of the 16-frame body in the instruction stream (the streaming kernel's warps interleave this loop with the band
    // code: the footprint decides whether both stay in the instruction cache -- a second copy of the body, tried as a register
    // ping-pong, cost 20 % through instruction-fetch stalls). Groups that are only partly inside the window (the first and the
    // last one) run the same body: their bytes outside the window are replaced by floor(median), which gives such a frame the
    // smallest value dm can take (-t < 0 since the host only takes this path for thr4 >= 5): never an outlier, and it can only
    // hold the group's maximum when no frame of the pixel is an outlier, in which case the maximum is not used.
#pragma unroll
    for (int sl = 0; sl < D - 1; sl++) {
        issue(ms.ring + sl * kSlotBytes, sl * kTilePixels);
        cp_async_commit();
    }
    const uint32_t ring_end = ms.ring + D * kSlotBytes;
    uint32_t rd = ms.ring, wr = ms.ring + (D - 1) * kSlotBytes;
    int off = (D - 1) * kTilePixels;
#pragma unroll 1
    for (int g = gA; g <= gB; g++, off += kTilePixels) {
        issue(wr, off);
        cp_async_commit();
        asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
        uint4 cur[C];
#pragma unroll
        for (int c = 0; c < C; c++) cur[c] = lds128(rd + c * kRingStride);
        rd += kSlotBytes; rd = rd == ring_end ? ms.ring : rd;
        wr += kSlotBytes; wr = wr == ring_end ? ms.ring : wr;
        const int sg = 16 * g - f0;  // window position of the group's frame 0
        if (!all_use) {  // (uniform, rare) a band of weight 0 must read as zero: its bytes of lof / negd are zero as well
#pragma unroll
            for (int c = 0; c < C; c++)
                if (a.w[c] == 0.0f) cur[c] = make_uint4(0, 0, 0, 0);
        }
        if (sg < 0 || sg + 16 > n) {  // (uniform) partial group
            const int lo_j = sg < 0 ? -sg : 0, hi_j = (n - sg < 16) ? n - sg : 16;  // valid frames j in [lo_j, hi_j)
            const uint32_t valid = ((1u << hi_j) - 1u) & ~((1u << lo_j) - 1u);
            uint32_t vm[4];
#pragma unroll
            for (int q = 0; q < 4; q++) vm[q] = ((((valid >> (4 * q)) & 0xfu) * 0x00204081u) & 0x01010101u) * 0xffu;  // bit kk -> byte kk
#pragma unroll
            for (int c = 0; c < C; c++) {
                const uint32_t lo4 = rep4((int)((lof >> (8 * c)) & 0xffu));
                cur[c].x = (cur[c].x & vm[0]) | (lo4 & ~vm[0]);
                cur[c].y = (cur[c].y & vm[1]) | (lo4 & ~vm[1]);
                cur[c].z = (cur[c].z & vm[2]) | (lo4 & ~vm[2]);
                cur[c].w = (cur[c].w & vm[3]) | (lo4 & ~vm[3]);
            }
        }
        uint32_t om16 = 0;
        // max over the group's frames of 19 * (q - t) + (15 - j): the first maximum of the distance wins. (19, not 16: a multiplier
        // that is no power of two plus or minus one keeps the per-frame key ONE IMAD on the FMA pipe; the ALU pipe -- PRMT,
        // VABSDIFF4, SHF at half rate -- is the pass's bottleneck.)
        int gkey = INT_MIN;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t xw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int c = 0; c < C; c++) xw[c] = q == 0 ? cur[c].x : (q == 1 ? cur[c].y : (q == 2 ? cur[c].z : cur[c].w));
            const uint32_t t0 = __byte_perm(xw[0], xw[1], 0x5140), t1 = __byte_perm(xw[0], xw[1], 0x7362);
            const uint32_t t2 = __byte_perm(xw[2], xw[3], 0x5140), t3 = __byte_perm(xw[2], xw[3], 0x7362);
            const uint32_t pfw[4] = {__byte_perm(t0, t2, 0x5410), __byte_perm(t0, t2, 0x7632), __byte_perm(t1, t3, 0x5410), __byte_perm(t1, t3, 0x7632)};
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const int j = 4 * q + kk;
                const uint32_t L = absdiff4(pfw[kk], lof);
                // dm = q - t with 4 dist_sq = 4 q + k2 (see IntMedians): negative <=> not an outlier
                const int dm = dp4a_us(pfw[kk], negd, (int)__dp4a(L, L, (uint32_t)c0));
                om16 = __funnelshift_l((uint32_t)dm, om16, 1);  // collects the NON-outlier bits, frame j at bit 15 - j
                gkey = max(gkey, dm * 19 + (15 - j));
            }
        }
        om16 = (~__brev(om16)) >> 16;
        ms.put(g - gA, om16);
        k += __popc(om16);
        const uint32_t gq = (uint32_t)(gkey + 19 * im.t);  // 19 q + (15 - j) >= 0
        const uint32_t qv = gq / 19u;
        const int s = sg + 15 - (int)(gq - 19u * qv);
        maxkey = max(maxkey, ((4u * qv + (uint32_t)im.k2) << 12) | (uint32_t)(4095 - s));
    }
    // the D - 1 copies issued past the window's last group are still in flight: they must have landed before this thread's next
    // pass starts writing the same ring slots (compute-sanitizer racecheck: write-after-write between two passes that follow
    // each other closely, e.g. the iterative tier's last warp-full and the first batch of the streaming kernel's queue)
    cp_async_wait_all();
    k_out = k;
    maxkey_out = maxkey;
}

// Everything calc_pixel does after the classification loop (src/chrono.rs:290-494), from the per-group outlier masks:
// set bits are the outlier list in frame order. Returns the mask byte; `pixel` receives the composite.
template <int C>
__device__ __forceinline__ uint8_t finish_masks(const OutlierArgs& a, const uint8_t* colbase, long long band_stride, unsigned long long pixel_id,
                                                const float (&median)[4], const uint32_t (&band_sum)[4], const IntMedians& im, const MaskSlots ms,
                                                int f0, int n, int frame_offset, int k, uint32_t maxkey, uint8_t (&pixel)[4], int& warn) {
    const int gA = f0 >> 4, n_g = ((f0 + n - 1) >> 4) - gA + 1;
    const int bit0 = f0 & 15;  // window position s sits at bit (s + bit0) & 15 of group (s + bit0) >> 4
    const uint32_t usemask = (a.w[0] != 0.0f ? 0xffu : 0u) | (a.w[1] != 0.0f ? 0xff00u : 0u) | (a.w[2] != 0.0f ? 0xff0000u : 0u) |
                             ((C > 3 && a.w[3] != 0.0f) ? 0xff000000u : 0u);
    auto sample = [&](int s) -> uint32_t {  // the pixel of window position s, band c in byte c
        const int f = f0 + s;
        const uint8_t* p = colbase + (long long)(f >> 4) * (kTilePixels * kUnitBytes) + (f & 15);
        uint32_t v = 0;
#pragma unroll
        for (int c = 0; c < C; c++) v |= (uint32_t)__ldg(p + c * band_stride) << (8 * c);
        return v;
    };
    auto unpack = [&](uint32_t v, uint8_t (&px)[4]) {
#pragma unroll
        for (int c = 0; c < 4; c++) px[c] = (uint8_t)((v >> (8 * c)) & 0xffu);
    };
    auto dist_of = [&](uint32_t v) { return 0.25f * (float)int_dist_frame(im, v & usemask); };  // dist_sq, exact in f32
    warn = 0;
    const bool has_outliers = k > 0;
    float out_sum[4] = {0.0f, 0.0f, 0.0f, 0.0f}, mean_dist = 0.0f;
    const bool need_avg = has_outliers && (a.om == 3 || a.bg == 2);
    if (need_avg) {  // outlier sums in frame order (the f32 sum of sqrt(dist_sq) is order-dependent)
        for (int gi = 0; gi < n_g; gi++) {
            uint32_t m = ms.get(gi);
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t v = sample(16 * gi + b - bit0);
#pragma unroll
                for (int c = 0; c < 4; c++) out_sum[c] += (float)((v >> (8 * c)) & 0xffu);
                mean_dist += sqrtf(dist_of(v));
            }
        }
    }
    // background (src/chrono.rs:294-375)
    if (a.bg == 2) {
        const float ratio = (float)n / (float)(n - k);
#pragma unroll
        for (int c = 0; c < C; c++) {
            const float mean = (float)band_sum[c] / (float)n;
            pixel[c] = has_outliers ? sat_u8(roundf(mean * ratio - out_sum[c] / (float)n)) : sat_u8(roundf(mean));
        }
    } else if (a.bg == 3) {
#pragma unroll
        for (int c = 0; c < C; c++) pixel[c] = sat_u8(roundf(median[c]));
    } else {
        int idx = 0;
        if (a.bg == 0) {  // first_excluded (src/chrono.rs:505-530)
            if (has_outliers) {
                if (k == n) warn = 1;
                else {
                    for (int gi = 0; gi < n_g; gi++) {
                        // valid bits of group gi: positions 16 gi + b - bit0 in [0, n)
                        const int lo_b = gi == 0 ? bit0 : 0, hi_b = min(16, n + bit0 - 16 * gi);
                        const uint32_t valid = ((1u << hi_b) - 1u) & ~((1u << lo_b) - 1u);
                        const uint32_t non = ~ms.get(gi) & valid;
                        if (non) { idx = 16 * gi + __ffs(non) - 1 - bit0; break; }
                    }
                }
            }
        } else {  // sample_excluded (src/chrono.rs:532-555), closed form of the position swaps (see exact_pixel)
            if (!has_outliers) idx = (int)rng_range(a.seed, pixel_id, 0, (uint32_t)n);
            else if (k == n) { idx = (int)rng_range(a.seed, pixel_id, 0, (uint32_t)n); warn = 1; }
            else {
                const int r = (int)rng_range(a.seed, pixel_id, 0, (uint32_t)(n - k));
                const int rb = r + bit0, rg = rb >> 4;
                idx = r;
                if ((ms.get(rg) >> (rb & 15)) & 1u) {
                    int order = __popc(ms.get(rg) & ((1u << (rb & 15)) - 1u));  // outliers before r
                    for (int gi = 0; gi < rg; gi++) order += __popc(ms.get(gi));
                    idx = n - 1 - order;
                }
            }
        }
        unpack(sample(idx), pixel);
    }
    if (!has_outliers) return 0;

    auto first_bit = [&]() {
        for (int gi = 0; gi < n_g; gi++) {
            const uint32_t m = ms.get(gi);
            if (m) return 16 * gi + __ffs(m) - 1 - bit0;
        }
        return 0;
    };
    auto last_bit = [&]() {
        for (int gi = n_g - 1; gi >= 0; gi--) {
            const uint32_t m = ms.get(gi);
            if (m) return 16 * gi + 31 - __clz(m) - bit0;
        }
        return 0;
    };
    uint8_t smp[4] = {0, 0, 0, 0};
    if (k > 1 && (a.om == 4 || a.om == 5)) {  // forward / backward (src/chrono.rs:391-427): the list in (reverse) frame order
        float pix_new[4], blend_inv = 1.0f;
#pragma unroll
        for (int c = 0; c < 4; c++) pix_new[c] = (float)pixel[c];
        for (int t = 0; t < n_g; t++) {
            const int gi = a.om == 4 ? t : n_g - 1 - t;
            uint32_t m = ms.get(gi);
            while (m) {
                const int b = a.om == 4 ? __ffs(m) - 1 : 31 - __clz(m);
                m &= ~(1u << b);
                const int s = 16 * gi + b - bit0;
                const uint32_t v = sample(s);
                unpack(v, smp);
                const float blend = fade_for(a.fade, s, n, frame_offset) * blend_value(a, sqrtf(dist_of(v)));
                blend_into_f32_u8(pix_new, smp, C, blend);
                blend_inv *= 1.0f - blend;
            }
        }
#pragma unroll
        for (int c = 0; c < C; c++) pixel[c] = sat_u8(roundf(pix_new[c]));
        return sat_u8(roundf((1.0f - blend_inv) * 255.0f));
    }
    int sidx;
    float dist;
    if (k > 1 && a.om == 3) {  // average (src/chrono.rs:430-468)
#pragma unroll
        for (int c = 0; c < C; c++) smp[c] = sat_u8(roundf(out_sum[c] / (float)k));
        sidx = 0;
        dist = mean_dist / (float)k;
    } else {  // a single outlier (:379-388), or first / last / extreme (:470-483)
        float dsq;
        if (k == 1 || a.om == 2 || a.om >= 3) {  // the only outlier is the maximum
            sidx = 4095 - (int)(maxkey & 4095u);
            dsq = 0.25f * (float)(maxkey >> 12);
            unpack(sample(sidx), smp);
        } else {
            sidx = a.om == 0 ? first_bit() : last_bit();
            const uint32_t v = sample(sidx);
            unpack(v, smp);
            dsq = dist_of(v);
        }
        dist = sqrtf(dsq);
    }
    const float blend = fade_for(a.fade, sidx, n, frame_offset) * blend_value(a, dist);  // src/chrono.rs:485-488
    blend_into_u8(pixel, smp, C, blend);
    return sat_u8(roundf(blend * 255.0f));
}

// Returns the mask byte; writes the composite pixel; n_out = number of outliers; warn = all-outlier warning.
// `active` lanes hold a pixel; inactive lanes run along (uniform loops) and their results are discarded.
// DENSE (single-window launches of K1): contiguous windows with the integer distance take dense_pass_int for pass 1.
template <int CT = 0, bool DENSE = false>
__device__ __forceinline__ uint8_t exact_pixel(const OutlierArgs& a, const PixelSrc& src, unsigned long long pixel_id,
                                               const float (&median)[4], const float (&iqr_inv)[4], const uint32_t (&band_sum)[4],
                                               uint8_t (&pixel)[4], int& n_out, int& warn, int contig_f0, int frame_offset) {
    // contig_f0 / frame_offset: OutlierArgs' values for a single-window launch; per lane in the chrono-video kernel,
    // where the pixels of one batch belong to different windows
    const int n = a.n, C = a.C;
    auto frame_at = [&](int s) { return contig_f0 >= 0 ? contig_f0 + s : __ldg(a.win_frames + s); };
    const float thr_sq = a.thr_sq;
    DistCtx dc;
    make_dist_ctx(a, median, iqr_inv, dc);
    WordScan ws;
    make_word_scan(a, median, iqr_inv, ws);
    IntDist idc;
    make_int_dist(a, median, idc);
    ColumnReader rd(src);
    // pass 1 (src/chrono.rs:261-288 plus the sums the policies need)
    int k = 0, first_idx = 0, last_idx = 0, max_index = 0, first_non = -1;
    float first_d = 0.0f, last_d = 0.0f, max_dist_sq = 0.0f, mean_dist = 0.0f;
    float out_sum[4] = {0, 0, 0, 0};
    uint8_t px[4] = {0, 0, 0, 0};
    const bool need_avg = a.om == 3 || a.bg == 2;
    const bool class_only = (a.om == 4 || a.om == 5) && !need_avg && !a.int_dist;  // (uniform) see the word loop of pass 1
    auto visit = [&](int s, float d) {  // one frame of the window, in order
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
    };
    // forward / backward chains without an average: pass 1 only has to CLASSIFY the frames (the chain walk computes the distances
    // it needs itself). Frames whose deviation in one band alone reaches the threshold are outliers without a distance
    // evaluation -- the frames an object dwells on --, the other candidates get the reference's f32 distance as before; counts
    // and first / last positions come off the flag word of the four frames.
    auto classify_word = [&](const uint32_t (&xw)[4], int s, uint32_t ex, uint32_t sure) {
        uint32_t out4 = sure, unc = ex & ~sure;
        while (unc) {
            const int bit = __ffs(unc) - 1;  // 8 kk + 7
            unc &= unc - 1u;
#pragma unroll
            for (int i = 0; i < 4; i++) px[i] = (uint8_t)((xw[i] >> (bit - 7)) & 0xffu);
            if (dist_sq_px(dc, px) >= thr_sq) out4 |= 1u << bit;
        }
        if (out4) {
            if (k == 0) first_idx = s + ((__ffs(out4) - 1) >> 3);
            last_idx = s + ((31 - __clz(out4)) >> 3);
            k += __popc(out4);
        }
        const uint32_t non4 = ~out4 & 0x80808080u;
        if (first_non < 0 && non4) first_non = s + ((__ffs(non4) - 1) >> 3);
    };
    if (DENSE && a.int_dist && contig_f0 >= 0 && n <= 4096) {  // (the pass packs window positions into 12 bits)
        // every frame classified by the branch-free integer pass; the per-policy values are read off its summary
        DensePass dp;
        const uint8_t* colbase = src.tile + (long long)src.p * kUnitBytes;
        const long long bstride = (long long)src.NG * kTilePixels * kUnitBytes;
        if (need_avg) dense_pass_int<CT, true>(a, colbase, bstride, contig_f0, n, median, dp);
        else dense_pass_int<CT, false>(a, colbase, bstride, contig_f0, n, median, dp);
        k = dp.k;
        if (k > 0) {
            max_index = 4095 - (int)(dp.maxkey & 4095u);
            max_dist_sq = 0.25f * (float)(dp.maxkey >> 12);
            first_idx = dp.first_idx; last_idx = dp.last_idx;
            auto d4_at = [&](int s) {  // 4 * dist_sq of one frame (src/chrono.rs:265-278 on integers, see IntDist)
                rd.fetch(contig_f0 + s, px);
                uint32_t d4 = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (i < C && a.w[i] != 0.0f) {
                        const int lo = (int)median[i], hi = lo + (median[i] != (float)lo ? 1 : 0);
                        const int t = abs((int)px[i] - lo) + abs((int)px[i] - hi);
                        d4 += (uint32_t)(t * t);
                    }
                }
                return 0.25f * (float)d4;
            };
            if (k == 1) { first_d = max_dist_sq; last_d = max_dist_sq; }  // the only outlier is the maximum
            else {
                if (a.om == 0) first_d = d4_at(first_idx);
                if (a.om == 1) last_d = d4_at(last_idx);
            }
            if (need_avg) {
                mean_dist = dp.mean_dist;
#pragma unroll
                for (int i = 0; i < 4; i++) out_sum[i] = dp.out_sum[i];
            }
        }
        first_non = dp.first_non == 0x7fffffff ? -1 : dp.first_non;
    } else if (contig_f0 >= 0) {  // contiguous window: position s is frame contig_f0 + s; aligned words are pre-tested four frames at a time
        int s = 0;
        const bool long_window = n >= 64;
        while (s < n) {
            const int f = contig_f0 + s;
            if (long_window && (f & 15) == 0 && s + 16 <= n) {
                // a whole frame group inside the window: its four words tested with compile-time word selects; a group without
                // a candidate -- the frames before and after an object's visit -- is passed in one step (short windows, e.g.
                // the chrono-video queue's, hold one or two whole groups: the extra test measured 1 % slower there)
                rd.enter_group(f >> 4);
                uint32_t any, xg[4];
                rd.template group_word<0>(xg); any = may_exceed(ws, xg);
                rd.template group_word<1>(xg); any |= may_exceed(ws, xg);
                rd.template group_word<2>(xg); any |= may_exceed(ws, xg);
                rd.template group_word<3>(xg); any |= may_exceed(ws, xg);
                if (any == 0) {
                    if (first_non < 0) first_non = s;
                    s += 16;
                    continue;
                }
                if (class_only) {  // the group's four words classified in place
                    uint32_t sure, ex;
                    rd.template group_word<0>(xg); ex = may_exceed_sure(ws, xg, sure); if (ex) classify_word(xg, s, ex, sure); else if (first_non < 0) first_non = s;
                    rd.template group_word<1>(xg); ex = may_exceed_sure(ws, xg, sure); if (ex) classify_word(xg, s + 4, ex, sure); else if (first_non < 0) first_non = s + 4;
                    rd.template group_word<2>(xg); ex = may_exceed_sure(ws, xg, sure); if (ex) classify_word(xg, s + 8, ex, sure); else if (first_non < 0) first_non = s + 8;
                    rd.template group_word<3>(xg); ex = may_exceed_sure(ws, xg, sure); if (ex) classify_word(xg, s + 12, ex, sure); else if (first_non < 0) first_non = s + 12;
                    s += 16;
                    continue;
                }
            }
            if ((f & 3) == 0 && s + 4 <= n) {
                uint32_t xw[4];
                rd.fetch_word(f, xw);
                uint32_t sure = 0;
                const uint32_t ex = class_only ? may_exceed_sure(ws, xw, sure) : may_exceed(ws, xw);
                if (ex == 0) {
                    if (first_non < 0) first_non = s;
                } else if (a.int_dist) {  // the word's four distances as integers
                    uint32_t d4[4];
                    int_dist4(idc, xw, d4);
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
                        if ((int)d4[kk] >= a.thr4) {
#pragma unroll
                            for (int i = 0; i < 4; i++) px[i] = (uint8_t)((xw[i] >> (8 * kk)) & 0xffu);
                            visit(s + kk, 0.25f * (float)d4[kk]);
                        } else if (first_non < 0) {
                            first_non = s + kk;
                        }
                    }
                } else if (class_only) {
                    classify_word(xw, s, ex, sure);
                } else {
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
                        if ((ex >> (8 * kk + 7)) & 1u) {
#pragma unroll
                            for (int i = 0; i < 4; i++) px[i] = (uint8_t)((xw[i] >> (8 * kk)) & 0xffu);
                            visit(s + kk, dist_sq_px(dc, px));
                        } else if (first_non < 0) {
                            first_non = s + kk;
                        }
                    }
                }
                s += 4;
            } else {
                rd.fetch(f, px);
                visit(s, dist_sq_px(dc, px));
                s++;
            }
        }
    } else {
        for (int s = 0; s < n; s++) {
            rd.fetch(frame_at(s), px);
            visit(s, dist_sq_px(dc, px));
        }
    }
    n_out = k;
    warn = 0;
    const bool has_outliers = k > 0;

    // background (src/chrono.rs:294-375)
    if (a.bg == 2) {  // Average
        float mean[4];
#pragma unroll
        for (int i = 0; i < 4; i++) mean[i] = (float)band_sum[i] / (float)n;  // the reference's f32 running sum is exact (< 2^24)
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
                rd.fetch(frame_at(r), tmp);
                if (dist_sq_px(dc, tmp) >= thr_sq) {
                    int order = 0;  // number of outliers before r
                    for (int s = 0; s < r; s++) {
                        rd.fetch(frame_at(s), tmp);
                        if (dist_sq_px(dc, tmp) >= thr_sq) order++;
                    }
                    idx = n - 1 - order;
                }
            }
        }
        const int f = frame_at(idx);
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < C) pixel[i] = src.at(f, i);
    }

    if (!has_outliers) return 0;

    uint8_t sample[4] = {0, 0, 0, 0};
    if (k == 1) {  // src/chrono.rs:379-388
        if (class_only) {  // the classification-only pass did not keep the distance of the single outlier
            rd.fetch(frame_at(first_idx), px);
            first_d = dist_sq_px(dc, px);
        }
        const int f = frame_at(first_idx);
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < C) sample[i] = src.at(f, i);
        const float fade = fade_for(a.fade, first_idx, n, frame_offset);
        const float blend = fade * blend_value(a, sqrtf(first_d));
        blend_into_u8(pixel, sample, C, blend);
        return sat_u8(roundf(blend * 255.0f));
    }
    if (a.om == 4 || a.om == 5) {  // forward / backward (src/chrono.rs:391-427): second walk in list order
        float pix_new[4], blend_inv = 1.0f;
#pragma unroll
        for (int i = 0; i < 4; i++) pix_new[i] = (float)pixel[i];
        // An outlier blended at full weight (blend >= 1) copies its sample, whatever the accumulator held, and its factor
        // (1 - 1) zeroes blend_inv for good: everything the list holds before its LAST full-weight entry is dead. Without a
        // fade (blend = blend_value, never above 1 and never NaN) a short walk from the end of the list looks for that entry
        // and the chain starts there -- for an opaque object that is the list's last entry. Not found within 16 frames: the
        // whole list is walked as written.
        int t0 = 0;
        if (a.fade.is_none) {
            const int span = last_idx - first_idx;
            for (int t = span; t >= 0 && t > span - 16; t--) {
                const int s = (a.om == 4) ? first_idx + t : last_idx - t;
                rd.fetch(frame_at(s), sample);
                const float d = dist_sq_px(dc, sample);
                if (d >= thr_sq && blend_value(a, sqrtf(d)) >= 1.0f) { t0 = t; break; }
            }
        }
        // only the span [first_idx, last_idx] holds outliers; in a contiguous window whole words without a candidate are skipped
        for (int ss = first_idx + t0; ss <= last_idx; ss++) {
            const int s = (a.om == 4) ? ss : last_idx - (ss - first_idx);
            const int f = frame_at(s);
            if (contig_f0 >= 0) {
                const int wf = f & ~3;  // word of this frame; when entering a word from its far end, test it once
                const bool entering = (a.om == 4) ? ((f & 3) == 0) : ((f & 3) == 3);
                if (entering && wf - contig_f0 >= first_idx && wf + 3 - contig_f0 <= last_idx) {
                    uint32_t xw[4];
                    rd.fetch_word(wf, xw);
                    if (may_exceed(ws, xw) == 0) { ss += 3; continue; }
                }
            }
            rd.fetch(f, sample);
            const float d = dist_sq_px(dc, sample);
            if (d >= thr_sq) {
                const float fade = fade_for(a.fade, s, n, frame_offset);
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
        const int f = frame_at(sidx);
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < C) sample[i] = src.at(f, i);
        dist = sqrtf(dsq);
    }
    const float fade = fade_for(a.fade, sidx, n, frame_offset);  // src/chrono.rs:485-488
    const float blend = fade * blend_value(a, dist);
    blend_into_u8(pixel, sample, C, blend);
    return sat_u8(roundf(blend * 255.0f));
}

// Median (and, for relative thresholds, quartiles and the inverse IQR) of one pixel-band (src/chrono.rs:238-255).
template <int W4, int G>
__device__ __forceinline__ void band_stats(int cap, const uint32_t (&xs)[W4], uint32_t ssum, float inv_cnt, const OutlierArgs& a,
                                           int pad, float& median, float& q1o, float& q3o, float& iqr_inv, int& center, float& halfw) {
    BandRanks r;
    band_solve<W4, G>(xs, ssum, inv_cnt, a, pad, cap, r);
    median = (r.mlo == r.mhi) ? (float)r.mlo : 0.5f * ((float)r.mlo + (float)r.mhi);  // src/chrono.rs:582-591
    center = (r.mlo + r.mhi) >> 1;
    halfw = median - (float)center;
    if (!a.absolute) {  // quartiles (src/chrono.rs:559-579) and inverse IQR (:246-252)
        const float q1 = (a.rk[0] == a.rk[1]) ? (float)r.q1a : (1.0f - a.q1_frac) * (float)r.q1a + a.q1_frac * (float)r.q1b;
        const float q3 = (a.rk[4] == a.rk[5]) ? (float)r.q3a : (1.0f - a.q3_frac) * (float)r.q3a + a.q3_frac * (float)r.q3b;
        q1o = q1;
        q3o = q3;
        float iq = q3 - q1;
        if (iq == 0.0f) iq = 1.0f;
        iqr_inv = 1.0f / iq;
    }
}

// Straight-line median pair for the common case: F at NP consecutive values p..p+NP-1 around the band mean (independent
// accumulator chains, no loop, no divergence) gives the NP-1 exact counts #{x <= p..p+NP-2}; if both ranks fall inside
// the window the pair is read off directly. Returns false when a rank lies outside (the pixel then takes the iterative
// solver). The byte-range ends count as known: #{x <= -1} = 0, #{x <= 255} = cap. NP = 5 for the median pair; NP = 7
// when the counts also have to bound the inter-quartile range (relative thresholds).
template <int W4, int G, int NP>
__device__ __forceinline__ bool band_window(const uint32_t (&x)[W4], int center, int kp1, int kp2, int cap, int& v1, int& v2, bool want_f,
                                            uint32_t& f_mid, int& p_out, int (&cn)[NP - 1]) {
    constexpr int kHalf = NP / 2;
    int p = center - kHalf;
    p = p < 0 ? 0 : (p > 256 - NP ? 256 - NP : p);
    uint32_t cc[NP], f[NP];
#pragma unroll
    for (int k = 0; k < NP; k++) { cc[k] = rep4(p) + 0x01010101u * (uint32_t)k; f[k] = 0; }
    if (NP <= 5) {  // two chains per value
        uint32_t h[NP];
#pragma unroll
        for (int k = 0; k < NP; k++) h[k] = 0;
#pragma unroll
        for (int q = 0; q < W4; q += 2) {
#pragma unroll
            for (int k = 0; k < NP; k++) f[k] = sad4_acc(x[q], cc[k], f[k]);
#pragma unroll
            for (int k = 0; k < NP; k++) h[k] = sad4_acc(x[q + 1], cc[k], h[k]);
        }
#pragma unroll
        for (int k = 0; k < NP; k++) f[k] += h[k];
    } else {
#pragma unroll
        for (int q = 0; q < W4; q++) {
#pragma unroll
            for (int k = 0; k < NP; k++) f[k] = sad4_acc(x[q], cc[k], f[k]);
        }
    }
    if (G == 1) {
#pragma unroll
        for (int k = 0; k < NP - 1; k++) cn[k] = ((int)f[k + 1] - (int)f[k] + cap) >> 1;
    } else {
        // per lane |F(c+1) - F(c)| <= bytes per lane = 4*W4: bias, pack two differences per word, one reduction each
        constexpr uint32_t kBias = 4 * W4;
#pragma unroll
        for (int k = 0; k < NP - 1; k += 2) {
            uint32_t pk = (f[k + 1] - f[k] + kBias) | ((f[k + 2 < NP ? k + 2 : k + 1] - f[k + 1] + kBias) << 16);
            pk = group_sum<G>(pk);
            cn[k] = (int)((pk & 0xffffu) >> 1);  // sum(d + bias) = 2 * count because G * bias = cap
            if (k + 1 < NP - 1) cn[k + 1] = (int)(pk >> 17);
        }
    }
    f_mid = 0;
    p_out = p;
    if (want_f) f_mid = group_sum<G>(f[kHalf]);  // F(p + NP/2): spread estimate for the quartile guesses (relative thresholds)
    v1 = p; v2 = p;
    if (kp1 == kp2) {  // (uniform) odd sample counts: the median pair is one rank
#pragma unroll
        for (int k = 0; k < NP - 1; k++) v1 += (cn[k] <= kp1);
        v2 = v1;
    } else {
#pragma unroll
        for (int k = 0; k < NP - 1; k++) { v1 += (cn[k] <= kp1); v2 += (cn[k] <= kp2); }
    }
    return ((cn[0] <= kp1) || p == 0) && ((kp2 < cn[NP - 2]) || p == 256 - NP);
}

// Median pair of one pixel-band (lane = pixel, G == 1) by a few straight-line windows: five F values per window give
// four exact counts for 260 VABSDIFF4 and a few bookkeeping instructions, where an iteration of band_solve gives one or two
// for 104 and a long state machine. Window 1 sits on the band mean (it resolves the bands that were not the reason the pixel
// came to this tier). Window 2 sits where band_solve's first re-jump would go: the mean of the samples on the median's side of
// the first guess (an object rests on the pixel: two clusters) or, when those samples spread far beyond the rank (iid bytes,
// wide noise), the first guess moved by the samples still to be passed over the density the mean absolute deviation implies.
// From then on every window leaves exact counts on both sides of the rank that is still open and the next window is placed
// where a uniform spread of the samples between them puts the rank, kept adjacent to a known count so that it always covers
// three new values. Returns true when every lane of the warp is resolved (warp-uniform); otherwise the caller runs band_solve.
template <int W4>
__device__ __forceinline__ bool band_solve_windows(const uint32_t (&x)[W4], uint32_t bsum, const OutlierArgs& a, int pad, int cap, int& mlo, int& mhi) {
    constexpr int kMaxWindows = 8;  // iid bytes: 4.2 windows for the slowest of 32 lanes; two far-apart clusters: 5.1 (simulated)
    const int kp1 = a.rk[2] + pad, kp2 = a.rk[3] + pad;
    int g = __float2int_rn((float)bsum * a.inv_n_sub);
    // the two ranks of the pair are resolved separately (their values may lie far apart); the bracket belongs to the rank kt
    // that is still open: (xl, cl) the largest value known to have #{x <= xl} <= kt, (xh, ch) the smallest with #{x <= xh} > kt
    int kt = kp1, xl = -1, cl = 0, xh = 255, ch = cap;
    bool have1 = false, have2 = false;
    mlo = mhi = 0;
#pragma unroll 1
    for (int it = 0; it < kMaxWindows; it++) {
        int w1, w2, p0, cn[4];
        uint32_t fm;
        band_window<W4, 1, 5>(x, g, kp1, kp2, cap, w1, w2, true, fm, p0, cn);
        const bool lo_end = p0 == 0, hi_end = p0 == 256 - 5;  // the ends of the byte range count as known
        if (!have1 && (cn[0] <= kp1 || lo_end) && (kp1 < cn[3] || hi_end)) {
            mlo = w1; have1 = true;
            if (kp2 != kp1 && !have2) { kt = kp2; xh = 255; ch = cap; }  // (xl, cl) stays valid for the larger rank
        }
        if (!have2 && (cn[0] <= kp2 || lo_end) && (kp2 < cn[3] || hi_end)) { mhi = w2; have2 = true; }
        const bool open = !(have1 && have2);
        if (!__any_sync(0xffffffffu, open)) return true;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (cn[k] <= kt && p0 + k > xl) { xl = p0 + k; cl = cn[k]; }
#pragma unroll
        for (int k = 3; k >= 0; k--)
            if (cn[k] > kt && p0 + k < xh) { xh = p0 + k; ch = cn[k]; }
        int g2;
        if (it == 0) {
            const int gc = p0 + 2, nj = cn[2];  // F(gc) = fm, #{x <= gc} = nj, F(gc + 1) = F(gc) + 2 nj - cap
            const bool up = kp1 >= nj;
            int side, far;
            float fside;
            if (up) {  // mean of the samples above gc
                const int sum_up = ((int)fm + 2 * nj - cap + (int)bsum - cap * (gc + 1)) >> 1;  // sum of (x - gc - 1) over x > gc
                side = gc + 1 + __float2int_rn(__fdividef((float)sum_up, (float)max(cap - nj, 1)));
                far = kp1 + 1 - nj;
                fside = (float)((int)fm + 2 * nj - cap) - (float)pad * (float)(gc + 1);
            } else {   // mean of the real samples <= gc (the pad zeros taken out)
                const int sum_dn = (((int)fm - (int)bsum + cap * gc) >> 1) - pad * gc;  // sum of (gc - x) over real x <= gc
                side = gc - __float2int_rn(__fdividef((float)sum_dn, (float)max(nj - pad, 1)));
                far = nj - kp1;
                fside = (float)fm - (float)pad * (float)gc;
            }
            const int step = __float2int_rn((float)far * 3.5f * fside * a.inv_n_sub * a.inv_n_sub);
            const int dside = up ? side - gc : gc - side;
            g2 = (4 * step < dside) ? (up ? gc + step : gc - step) : side;
        } else if (it >= 3 && (it & 1)) {
            g2 = (xl + xh + 1) >> 1;  // a plain bisection now and then bounds the walk along a flat stretch of the counts
        } else {
            g2 = xl + __float2int_rn(__fdividef((float)(kt + 1 - cl) * (float)(xh - xl), (float)max(ch - cl, 1)));
        }
        // p0 = g - 2 in [xl, xh - 3]: count 0 of the window is then a known "<= kt" count (or the range end), count 3 a known
        // "> kt" one once the bracket is narrow; an open bracket shrinks by at least three values per window
        const int g_lo = xl + 2, g_hi = max(xl + 2, xh - 1);
        g2 = g2 < g_lo ? g_lo : (g2 > g_hi ? g_hi : g2);
        if (open) g = g2;
    }
    return !__any_sync(0xffffffffu, !(have1 && have2));
}

// ------------------------------------------------------------------------------------------------ K1
// One warp = one tile slice: 32/G pixels x G lanes per pixel. A pixel-band's whole time series (WPL 16-frame units per
// lane; slot i of lane j holds frame group g0 + i*G + j) is held in registers while its order statistics and its
// certificate term are computed. With G == 1 a tile's band is one contiguous slab of the stack, and the NEXT pixel-band
// is staged into shared memory by one TMA bulk copy (cp.async.bulk + mbarrier) while the current one is processed, so HBM
// latency overlaps the arithmetic; every byte of the stack is read from HBM exactly once.
// Three tiers per pixel:
//   fast    every band's median pair (and quartile pairs) inside their 5-value windows, certificate holds -> done from registers
//   hard    an order statistic outside its window: iterative solver (band_solve); such pixels are queued per
//           warp in shared memory and re-processed a warp-full at a time
//   exact   certificate fails: exact_pixel walks the frames, 32 queued pixels at a time
// GENERIC = false is the whole-stack launch whose leading slots are all real frame groups; GENERIC = true adds window
// masks, spare-capacity slots and the --sample subset.
#ifndef CHB_WARPS
#define CHB_WARPS 8
#endif
#ifndef CHB_MINB
#define CHB_MINB 2
#endif
#ifndef CHB_HARD_MINB
#define CHB_HARD_MINB 2
#endif
#ifndef CHB_EXACT_MINB
#define CHB_EXACT_MINB 4
#endif
constexpr int kWarpsPerCta = CHB_WARPS;
constexpr int kQueueCap = 64;  // per warp
struct QueueEntry {
    long long pix;
    float median[4];
    float iqr_inv[4];
    uint32_t sum[4];
};
constexpr int kBarBytes = 128;
constexpr int kAccBytes = kWarpsPerCta * 32 * 12 * 4;
static_assert(kWarpsPerCta * 32 * 16 == kRingStride, "the dense pass's ring is laid out for 256 threads");
constexpr int kMaskBytes = kWarpsPerCta * 32 * kMaskGroups * 2;  // per-thread outlier masks of the in-kernel dense pass
// dynamic shared memory of one CTA: one mbarrier per warp, per-thread result slots, one staged pixel-band per warp (G == 1: the tile's contiguous slab; G > 1: every lane's own units, [slot][lane])
__host__ __device__ constexpr int dense_ring_bytes(int C, int threads) { return dense_ring_depth(C) * C * threads * 16; }
__host__ __device__ constexpr int outlier_smem_bytes(int wpl, int g, int C) {
    // g == 1: one slab per warp; g > 1: wpl units per lane; then the mask slots and (lane = pixel kernels only) the dense pass's ring
    return kBarBytes + kAccBytes + kWarpsPerCta * wpl * 512 + kMaskBytes + (g == 1 ? dense_ring_bytes(C, kWarpsPerCta * 32) : 0);
}

template <int C>
__device__ __forceinline__ void store_pixel(const OutlierArgs& a, long long pix, const uint8_t (&pixel)[4], uint8_t mask) {
    // src/chrono.rs:183-191
#pragma unroll
    for (int c = 0; c < C; c++) {
        a.out_image[pix * C + c] = pixel[c];
        if (a.out_mask) a.out_mask[pix * C + c] = (c < 3) ? mask : 255;
    }
}

// Dense per-frame pass + finish for the 32 pixels (one per lane) whose unit columns start at colbase; lanes that are not
// `active` run along and their results are discarded. Returns the ballot of active lanes with an all-outlier warning.
template <int C>
__device__ __forceinline__ unsigned dense_pixels(const OutlierArgs& a, const uint8_t* colbase, long long pix, bool active, const float (&median)[4],
                                                 const uint32_t (&band_sum)[4], const MaskSlots ms) {
    const long long bstride = (long long)a.NG * kTilePixels * kUnitBytes;
    const IntMedians im = make_int_medians<C>(a, median);
    int k = 0, warn = 0;
    uint32_t maxkey = 0;
    dense_masks_int<C>(a, colbase, bstride, a.contig_f0, a.n, im, ms, k, maxkey);
    uint8_t pixel[4] = {0, 0, 0, 0};
    const uint8_t mask = finish_masks<C>(a, colbase, bstride, pixel_gid(a, pix), median, band_sum, im, ms, a.contig_f0, a.n,
                                         a.frame_offset, k, maxkey, pixel, warn);
    if (active) {
        store_pixel<C>(a, pix, pixel, mask);
        if (a.dbg_nout) a.dbg_nout[pix] = k;
    }
    return __ballot_sync(0xffffffffu, active && warn);
}

template <int C, bool MASK_ONLY = false>
__device__ __noinline__ void drain_queue(const OutlierArgs& a, const QueueEntry* slot, bool in_range, int lane, const MaskSlots ms) {
    long long pix = in_range ? slot->pix : -1;
    const bool active = pix >= 0;
    const unsigned act = __ballot_sync(0xffffffffu, active);
    if (act == 0) return;
    {   // lanes without a pixel run along on the first active lane's entry; their results are discarded
        const unsigned long long first = __shfl_sync(0xffffffffu, (unsigned long long)slot, __ffs(act) - 1);
        if (!active) { slot = reinterpret_cast<const QueueEntry*>(first); pix = slot->pix; }
    }
    const long long tile = pix >> 5;
    const uint8_t* tile_base = a.stack + tile * tile_bytes(C, a.NG);
    unsigned wb;
    if (MASK_ONLY || a.mask_path) {
        float median[4];
        uint32_t sum[4];
#pragma unroll
        for (int c = 0; c < 4; c++) { median[c] = slot->median[c]; sum[c] = slot->sum[c]; }
        wb = dense_pixels<C>(a, tile_base + (pix & 31) * kUnitBytes, pix, active, median, sum, ms);
    } else if (!MASK_ONLY) {
        const QueueEntry e = *slot;
        const PixelSrc src{tile_base, a.NG, C, (int)(pix & 31)};
        uint8_t pixel[4] = {0, 0, 0, 0};
        int n_out = 0, warn = 0;
        const uint8_t mask = exact_pixel<C, true>(a, src, pixel_gid(a, pix), e.median, e.iqr_inv, e.sum, pixel, n_out, warn, a.contig_f0, a.frame_offset);
        if (active) {
            store_pixel<C>(a, pix, pixel, mask);
            if (a.dbg_nout) a.dbg_nout[pix] = n_out;
        }
        wb = __ballot_sync(0xffffffffu, active && warn);
    }
    if (lane == 0) {
        if (wb) atomicAdd(a.counters, (unsigned long long)__popc(wb));
        atomicAdd(a.counters + 1, (unsigned long long)__popc(act));
    }
}

// What a pixel accumulates over its bands. The per-band results (median, 1/IQR, sum) are only needed when the pixel
// is finished, so they live in per-thread shared-memory slots ([field][thread], conflict-free) instead of registers.
constexpr int kAccWords = 12;
struct PixelAcc {
    uint32_t slot;      // shared-memory address of this thread's slots: field k at slot + k * kAccStride
    uint32_t first_px;  // bytes of window position 0, band c in byte c
    float bound;        // certificate: upper bound of dist_sq over the window's frames
    bool hard;          // a median pair fell outside the 5-value window
    bool approx;        // iqr_inv holds an upper bound, not the exact value (relative thresholds, fast tier)
    static constexpr uint32_t kAccStride = CHB_WARPS * 32 * 4;
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int i = 0; i < kAccWords; i++) sts32(slot + i * kAccStride, 0u);
        first_px = 0; bound = 0.0f; hard = false; approx = false;
    }
    __device__ __forceinline__ void set_median(int c, float v) { sts32(slot + c * kAccStride, __float_as_uint(v)); }
    __device__ __forceinline__ void set_iqr_inv(int c, float v) { sts32(slot + (4 + c) * kAccStride, __float_as_uint(v)); }
    __device__ __forceinline__ void set_sum(int c, uint32_t v) { sts32(slot + (8 + c) * kAccStride, v); }
    __device__ __forceinline__ float median(int c) const { return __uint_as_float(lds32(slot + c * kAccStride)); }
    __device__ __forceinline__ float iqr_inv(int c) const { return __uint_as_float(lds32(slot + (4 + c) * kAccStride)); }
    __device__ __forceinline__ uint32_t sum(int c) const { return lds32(slot + (8 + c) * kAccStride); }
};

// One pixel-band held in A: window mask, sum, order statistics, certificate term. FAST: try the straight-line window first
// (absolute thresholds only); otherwise run the iterative solver.
template <int C, int WPL, int G, int MODE, bool FAST>
__device__ __forceinline__ void process_band(const OutlierArgs& a, uint32_t (&A)[4 * WPL], int c, int j, long long pix, bool write_dbg,
                                             int cap, int pad, PixelAcc& acc, bool skip_cert = false) {
    constexpr int W4 = 4 * WPL;
    constexpr bool GENERIC = (MODE == 0);  // MODE 1 / 2: lean whole-stack kernels for absolute / relative thresholds
    const float w = a.w[c];
    if (GENERIC && a.window_masked) {  // frames outside the window must read as zero
#pragma unroll
        for (int i = 0; i < WPL; i++) {
            const uint4 m = __ldg(reinterpret_cast<const uint4*>(a.wmask) + (i * G + j));
            A[4 * i] &= m.x; A[4 * i + 1] &= m.y; A[4 * i + 2] &= m.z; A[4 * i + 3] &= m.w;
        }
    }
    {   // window position 0 lives in slot 0 of lane j == 0; its byte index is uniform
        const int wsel = (a.first_frame >> 2) & 3, sh = (a.first_frame & 3) * 8;
        const uint32_t wv = wsel == 0 ? A[0] : (wsel == 1 ? A[1] : (wsel == 2 ? A[2] : A[3]));
        acc.first_px |= ((wv >> sh) & 0xffu) << (8 * c);
    }
    uint32_t bsum = 0;
    if (a.bg == 2 || w != 0.0f) {  // window sum (IDP.4A: FMA pipe)
        uint32_t s0 = 0, s1 = 0;
#pragma unroll
        for (int q = 0; q < W4; q += 2) { s0 = __dp4a(A[q], 0x01010101u, s0); s1 = __dp4a(A[q + 1], 0x01010101u, s1); }
        bsum = group_sum<G>(s0 + s1);
        acc.set_sum(c, bsum);
    }
    if (w == 0.0f) return;
    float med = 0.0f, q1 = 0.0f, q3 = 0.0f, iqi = 0.0f, halfw = 0.0f;
    int center = 0;
    bool solved = false;
    if (GENERIC && a.smask) {  // --sample: order statistics on the subset only
        uint32_t xs[W4];
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < WPL; i++) {
            const uint4 m = __ldg(reinterpret_cast<const uint4*>(a.smask) + (i * G + j));
            xs[4 * i] = A[4 * i] & m.x; xs[4 * i + 1] = A[4 * i + 1] & m.y;
            xs[4 * i + 2] = A[4 * i + 2] & m.z; xs[4 * i + 3] = A[4 * i + 3] & m.w;
        }
#pragma unroll
        for (int q = 0; q < W4; q++) s = __dp4a(xs[q], 0x01010101u, s);
        band_stats<W4, G>(cap, xs, group_sum<G>(s), a.inv_n_sub, a, pad, med, q1, q3, iqi, center, halfw);
        solved = true;
    } else if (FAST && (MODE != 0 || a.absolute || !a.exact_quartiles)) {
        // (exact quartiles for the debug planes are only produced by the iterative solver below, GENERIC kernel)
        int mlo, mhi, p0;
        uint32_t fm;
        const bool rel = (MODE == 2) || (MODE == 0 && !a.absolute);
        const int guess = __float2int_rn((float)bsum * a.inv_n_sub);
        bool ok;
        if (!rel) {
            int cn[4];
            const int kp1 = a.rk[2] + pad, kp2 = a.rk[3] + pad;
            ok = band_window<W4, G, 5>(A, guess, kp1, kp2, cap, mlo, mhi, false, fm, p0, cn);
        } else {
            // The certificate only needs an UPPER bound of 1/IQR, i.e. a lower bound of the IQR, and the counts of a
            // 7-value window give one: Q1 <= d[hi rank of the Q1 pair] <= p + #{counts <= that rank} (valid when the rank is
            // reached inside the window) and Q3 >= d[lo rank of the Q3 pair] >= p + #{counts <= that rank} (valid when at
            // least the first count is). A pixel this bound cannot clear gets exact quartiles on the iterative tier.
            int cn[6];
            ok = band_window<W4, G, 7>(A, guess, a.rk[2] + pad, a.rk[3] + pad, cap, mlo, mhi, false, fm, p0, cn);
            const int r1 = a.rk[1] + pad, r4 = a.rk[4] + pad;
            int ub1 = p0, lb4 = p0;
#pragma unroll
            for (int k = 0; k < 6; k++) { ub1 += (cn[k] <= r1); lb4 += (cn[k] <= r4); }
            const bool ub_ok = cn[5] >= r1 + 1, lb_ok = (cn[0] <= r4) || p0 == 0;
            const int L = lb4 - ub1;
            if (ub_ok && lb_ok && L >= 1) iqi = 1.0f / (float)L;
            else ok = false;
            acc.approx = true;
        }
        med = (mlo == mhi) ? (float)mlo : 0.5f * ((float)mlo + (float)mhi);  // src/chrono.rs:582-591
        center = (mlo + mhi) >> 1;
        halfw = med - (float)center;
        acc.hard = acc.hard || !ok;
        solved = true;
    }
    if (!solved && !FAST && G == 1 && (MODE == 1 || (MODE == 0 && a.absolute)) && a.hard_window) {
        // Iterative tier, absolute thresholds (one rank pair per band): a short iteration of straight-line windows before the
        // solver's general loop (see band_solve_windows); only a warp-full that still holds an unresolved pair runs the solver.
        int mlo, mhi;
        if (band_solve_windows<W4>(A, bsum, a, pad, cap, mlo, mhi)) {
            med = (mlo == mhi) ? (float)mlo : 0.5f * ((float)mlo + (float)mhi);  // src/chrono.rs:582-591
            center = (mlo + mhi) >> 1;
            halfw = med - (float)center;
            solved = true;
        }
    }
    if (!solved) band_stats<W4, G>(cap, A, bsum, a.inv_n_sub, a, pad, med, q1, q3, iqi, center, halfw);
    acc.set_median(c, med);
    acc.set_iqr_inv(c, iqi);
    if (a.dbg_median && write_dbg) {  // per-band sub-results (planes are zeroed by the host)
        a.dbg_median[pix * 4 + c] = med;
        if (a.dbg_q1) a.dbg_q1[pix * 4 + c] = q1;
        if (a.dbg_q3) a.dbg_q3[pix * 4 + c] = q3;
    }
    if (skip_cert) {  // (warp-uniform) the tile is going to be classified frame by frame whatever this band's term is
        acc.bound = __int_as_float(0x7f800000);
    } else if (!(w < 0.0f)) {  // negative weights only lower dist_sq; NaN poisons the bound (-> exact path)
        // certificate term: an upper bound of |x - median| over the window's frames. Bytes that are not window frames
        // (zero in the registers) are replaced by the centre value, so they contribute 0.
        const uint32_t cc = rep4(center);
        {
            constexpr int i = WPL - 1;  // last slot: always (its mask is all ones when nothing needs patching)
            const uint4 m = __ldg(reinterpret_cast<const uint4*>(a.wmask) + (i * G + j));
            A[4 * i] |= cc & ~m.x; A[4 * i + 1] |= cc & ~m.y; A[4 * i + 2] |= cc & ~m.z; A[4 * i + 3] |= cc & ~m.w;
        }
        if (GENERIC && a.patch_slots) {
#pragma unroll
            for (int i = 0; i < WPL - 1; i++) {
                const uint4 m = __ldg(reinterpret_cast<const uint4*>(a.wmask) + (i * G + j));
                A[4 * i] |= cc & ~m.x; A[4 * i + 1] |= cc & ~m.y; A[4 * i + 2] |= cc & ~m.z; A[4 * i + 3] |= cc & ~m.w;
            }
        }
        uint32_t o0 = 0, o1 = 0;
#pragma unroll
        for (int q = 0; q < W4; q += 2) {
            o0 |= absdiff4(A[q], cc);
            o1 |= absdiff4(A[q + 1], cc);
        }
        uint32_t o = o0 | o1;
        o |= o >> 16;
        o |= o >> 8;
        o = group_or<G>(o & 0xffu);  // >= max over frames of |x - centre| (OR dominates max)
        const float aw = a.absolute ? w : w * iqi;
        float t = aw * ((float)o + halfw);
        // The OR overshoots the maximum by up to 2x (4 | 3 = 7, 8 | 7 = 15). Where that alone costs a pixel its certificate --
        // the bound fails with the OR but would hold with the highest set bit of the OR, which the maximum is at least -- the
        // warp takes a second pass for the exact byte-wise maximum (VIMNMX.U16x2: the high byte of each half is a byte maximum).
        // Series whose noise stays below the OR's next power of two (the S2 series at abs/0.05) never take it.
        // (absolute thresholds only: with relative ones the term also carries the IQR bound's slack and the pass rarely pays)
        if (MODE == 1 || (MODE == 0 && a.absolute)) {
            const bool fails = !((acc.bound + t * t) * 1.0001f < a.thr_sq);
            if (__any_sync(0xffffffffu, fails)) {  // (rare: a tile that touches an object or heavy noise)
                const float t_lo = aw * ((float)(o ? (1u << (31 - __clz(o))) : 0u) + halfw);
                const bool could = fails && (acc.bound + t_lo * t_lo) * 1.0001f < a.thr_sq;
                if (__any_sync(0xffffffffu, could)) {
                    uint32_t m0 = 0, m1 = 0;
#pragma unroll
                    for (int q = 0; q < W4; q += 2) {
                        const uint32_t d0 = absdiff4(A[q], cc), d1 = absdiff4(A[q + 1], cc);
                        m0 = __vmaxu2(m0, d0); m0 = __vmaxu2(m0, d0 << 8);
                        m1 = __vmaxu2(m1, d1); m1 = __vmaxu2(m1, d1 << 8);
                    }
                    const uint32_t m = __vmaxu2(m0, m1);
                    uint32_t mx = max((m >> 8) & 0xffu, m >> 24);
#pragma unroll
                    for (int sh = 32 / G; sh < 32; sh <<= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, sh));
                    t = aw * ((float)mx + halfw);
                }
            }
        }
        acc.bound += t * t;
    }
}

// Certified pixels are written from registers; the others are appended to the launch's exact-path queue.
template <int C>
__device__ __forceinline__ void finish_pixel(const OutlierArgs& a, const PixelAcc& acc, long long pix, int p_in_tile, bool owner, int lane,
                                             long long slot = -1) {
    const bool clean = acc.bound * 1.0001f < a.thr_sq;  // margin covers the f32 roundings of the reference's sum
    if (owner && clean) {
        uint8_t pixel[4] = {0, 0, 0, 0};
        if (a.bg == 2) {
#pragma unroll
            for (int c = 0; c < C; c++) pixel[c] = sat_u8(roundf((float)acc.sum(c) / (float)a.n));  // src/chrono.rs:297-306,335-337
        } else if (a.bg == 3) {
#pragma unroll
            for (int c = 0; c < C; c++) pixel[c] = sat_u8(roundf(acc.median(c)));  // :340-345
        } else if (a.bg == 0) {
#pragma unroll
            for (int c = 0; c < C; c++) pixel[c] = (uint8_t)((acc.first_px >> (8 * c)) & 0xffu);  // :348-350
        } else {
            const int pos = (int)rng_range(a.seed, pixel_gid(a, pix), 0, (uint32_t)a.n);  // :357
            const int f = __ldg(a.win_frames + pos);
            const PixelSrc src{a.stack + (pix >> 5) * tile_bytes(C, a.NG), a.NG, C, p_in_tile};
#pragma unroll
            for (int c = 0; c < C; c++) pixel[c] = src.at(f, c);
        }
        store_pixel<C>(a, pix, pixel, 0);
        if (a.dbg_nout) a.dbg_nout[pix] = 0;
    }
    // Uncertified pixels go to the launch's global queue (one slot per pixel of the band, so it cannot overflow) and are
    // finished by outlier_exact_kernel right after: the exact path never enters this kernel's instruction stream.
    const bool dirty = owner && !clean;
    if (slot >= 0) {  // iterative tier: the pixel's own slot (its position in that tier's queue), certified or not
        if (owner) {
            QueueEntry& e = a.gq[slot];
            e.pix = dirty ? pix : -1;
            if (dirty) {
#pragma unroll
                for (int c = 0; c < 4; c++) { e.median[c] = acc.median(c); e.iqr_inv[c] = acc.iqr_inv(c); e.sum[c] = acc.sum(c); }
            }
        }
        return;
    }
    const unsigned db = __ballot_sync(0xffffffffu, dirty);
    if (db) {  // streaming kernel: appended from the far end of the array
        unsigned int gbase = 0;
        if (lane == 0) gbase = atomicAdd(a.gq_count, (unsigned int)__popc(db));
        gbase = __shfl_sync(0xffffffffu, gbase, 0);
        if (dirty) {
            QueueEntry& e = a.gq[a.n_pixels - 1 - (long long)(gbase + __popc(db & ((1u << lane) - 1u)))];
            e.pix = pix;
#pragma unroll
            for (int c = 0; c < 4; c++) { e.median[c] = acc.median(c); e.iqr_inv[c] = acc.iqr_inv(c); e.sum[c] = acc.sum(c); }
        }
    }
}

// Hard pixels, 32/G at a time: each pixel group reloads its own bands (L2) and runs the iterative solver.
template <int C, int WPL, int G, int MODE>
__device__ __noinline__ void drain_hard(const OutlierArgs& a, unsigned int hbase, int count, int lane, int cap, int pad, uint32_t acc_slot, const MaskSlots ms) {
    const long long* hq = a.ghq + hbase;
    constexpr int W4 = 4 * WPL;
    constexpr long long kSlotStride = (long long)G * kTilePixels * kUnitBytes;
    const int j = lane / (32 / G), pl = lane % (32 / G);
    const bool active = pl < count;
    const long long pix = hq[active ? pl : 0];
    const long long tile = pix >> 5;
    const int p = (int)(pix & 31);
    const long long band_stride = (long long)a.NG * (kTilePixels * kUnitBytes);
    const uint8_t* base = a.stack + tile * tile_bytes(C, a.NG) + ((long long)(a.g0 + j) * kTilePixels + p) * kUnitBytes;
    PixelAcc acc;
    acc.slot = acc_slot;
    acc.reset();
    uint32_t A[W4];
#pragma unroll 1
    for (int c = 0; c < C; c++) {
#pragma unroll
        for (int i = 0; i < WPL; i++) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (i * G + j < a.n_groups) v = __ldg(reinterpret_cast<const uint4*>(base + c * band_stride + i * kSlotStride));
            A[4 * i + 0] = v.x; A[4 * i + 1] = v.y; A[4 * i + 2] = v.z; A[4 * i + 3] = v.w;
        }
        process_band<C, WPL, G, MODE, false>(a, A, c, j, pix, active && j == 0, cap, pad, acc);
    }
    // the uncertified pixels of the warp-full are classified frame by frame right here, while their units are hot in L1 / L2,
    // instead of going through the exact-path queue (their queue slots are marked empty): a pixel whose median needed the solver
    // almost always holds an outlier, and the last launch of the call is left with the streaming kernel's own few pixels
    bool finished_here = false;
    if (G == 1 && MODE != 2 && a.hard_inline_min > 0) {
        const bool dirty = active && !(acc.bound * 1.0001f < a.thr_sq);
        const unsigned db = __ballot_sync(0xffffffffu, dirty);
        if (__popc(db) >= a.hard_inline_min) {
            float median[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            uint32_t sum[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int c = 0; c < C; c++) { median[c] = acc.median(c); sum[c] = acc.sum(c); }
            const uint8_t* colbase = a.stack + tile * tile_bytes(C, a.NG) + p * kUnitBytes;
            const unsigned wb = dense_pixels<C>(a, colbase, pix, dirty, median, sum, ms);
            if (lane == 0) {
                if (wb) atomicAdd(a.counters, (unsigned long long)__popc(wb));
                atomicAdd(a.counters + 1, (unsigned long long)__popc(db));
            }
            if (dirty) a.gq[(long long)hbase + pl].pix = -1;
            finished_here = dirty;
        }
    }
    finish_pixel<C>(a, acc, pix, p, active && j == 0 && !finished_here, lane, (long long)hbase + pl);
}

// One tile finished inside the streaming kernel: lane = pixel of the tile, medians and sums from the thread's result slots.
template <int C>
__device__ __noinline__ void inline_dense_tile(const OutlierArgs& a, int tile, int lane, long long pix, bool dirty, const PixelAcc& acc, const MaskSlots ms) {
    float median[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    uint32_t sum[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int c = 0; c < C; c++) { median[c] = acc.median(c); sum[c] = acc.sum(c); }
    const uint8_t* colbase = a.stack + (long long)tile * tile_bytes(C, a.NG) + lane * kUnitBytes;
    const unsigned wb = dense_pixels<C>(a, colbase, pix, dirty, median, sum, ms);
    if (lane == 0 && wb) atomicAdd(a.counters, (unsigned long long)__popc(wb));
}

template <int C, int WPL, int G, int MODE>
__global__ void __launch_bounds__(kWarpsPerCta * 32, CHB_MINB) outlier_kernel(const __grid_constant__ OutlierArgs a) {
    constexpr bool GENERIC = (MODE == 0);
    constexpr int W4 = 4 * WPL;
    constexpr int PPW = 32 / G;
    constexpr long long kSlotStride = (long long)G * kTilePixels * kUnitBytes;
    // G == 1: a tile's band is one contiguous slab, staged by ONE TMA bulk copy per pixel-band. With G > 1 a slab is
    // n_groups rows of 512/G bytes (one small bulk copy per row measured slower than direct loads: 10.4 vs 8.7 ms on the
    // 1000-frame UHD stack); there every lane prefetches its OWN 16-byte units of the next pixel-band with cp.async into a
    // lane-private strip of shared memory while the current band is processed.
    constexpr bool kStage = (G == 1);
    constexpr int kRowBytes = 512 / G;  // bytes of one (band, group) row that belong to this warp's 32/G pixels
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp_in_cta = threadIdx.x >> 5;
    const int j = lane / PPW, pl = lane % PPW;
    const int n_warps = (int)((gridDim.x * blockDim.x) >> 5);
    const int n_tasks = (int)(a.n_tiles * G);  // the host keeps n_tiles * G below 2^31
    const int cap = W4 * 4 * G;     // bytes per pixel-band across the G lanes
    const int pad = cap - a.n_sub;  // zero bytes that take part in the selection
    uint64_t* const bar = reinterpret_cast<uint64_t*>(smem_raw) + warp_in_cta;
    const uint32_t acc_slot = smem_u32(smem_raw + kBarBytes) + threadIdx.x * 4;
    uint8_t* const stage = smem_raw + kBarBytes + kAccBytes + warp_in_cta * (WPL * 512);
    const uint32_t stage_lane = smem_u32(stage) + j * kRowBytes + pl * 16;  // this lane's 16 bytes of row (slot * G + j)
    const int staged_groups = a.n_groups < WPL * G ? a.n_groups : WPL * G;
    const MaskSlots mask_slots{smem_u32(smem_raw + kBarBytes + kAccBytes + kWarpsPerCta * (WPL * 512)) + threadIdx.x * 2, kWarpsPerCta * 32u * 2u,
                               smem_u32(smem_raw + kBarBytes + kAccBytes + kWarpsPerCta * (WPL * 512) + kMaskBytes) + threadIdx.x * 16};
    uint32_t parity = 0;
    if (kStage) {
        if (lane == 0) mbar_init(bar, 1);
        __syncwarp();
    }
    // Starts the copy of one pixel-band of a tile slice: with G == 1 the slab is contiguous (one bulk copy by one lane);
    // with G > 1 every frame group contributes a row of 512/G bytes, copied by the lanes in parallel onto one mbarrier.
    auto stage_band = [&](int task, int c) {
        const int tile = task / G;
        const uint8_t* src = a.stack + (long long)tile * tile_bytes(C, a.NG) + ((long long)c * a.NG + a.g0) * (kTilePixels * kUnitBytes) + (task % G) * kRowBytes;
        if (G == 1) {
            if (lane == 0) {
                // the warp's ld.shared reads of the slab (generic proxy, ordered before this point by __syncwarp) must be ordered
                // before the bulk copy's writes (async proxy): without the proxy fence the copy may overwrite bytes a delayed
                // read has not fetched yet -- seen as medians off by one in ~2 of 750 000 tiles when the load / store unit is busy
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, (uint32_t)staged_groups * 512u);
                bulk_g2s(stage, src, (uint32_t)staged_groups * 512u, bar);
            }
        } else {
            if (lane == 0) mbar_expect_tx(bar, (uint32_t)staged_groups * kRowBytes);
            __syncwarp();
            for (int r = lane; r < staged_groups; r += 32) bulk_g2s(stage + r * kRowBytes, src + (long long)r * 512, kRowBytes, bar);
        }
    };

    // G > 1: this lane's units of one pixel-band, slot i at strip + i * 512 (conflict-free 128-bit reads)
    const uint32_t strip_lane = smem_u32(stage) + lane * 16;
    auto prefetch_band = [&](int task, int c) {
        const int tile = task / G;
        const int p = (task % G) * PPW + pl;
        const uint8_t* cb = a.stack + (long long)tile * tile_bytes(C, a.NG) + ((long long)(c * a.NG + a.g0 + j) * kTilePixels + p) * kUnitBytes;
#pragma unroll
        for (int i = 0; i < WPL; i++)
            if ((!GENERIC && i < WPL - 1) || i * G + j < a.n_groups) cp_async16(strip_lane + i * 512, cb + i * kSlotStride);
        cp_async_commit();
    };

    uint32_t A[W4];  // the current pixel-band
    int task = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (task < n_tasks) {
        if (kStage) stage_band(task, 0);
        else prefetch_band(task, 0);
    }
    while (task < n_tasks) {
        const int tile = task / G;
        const int p = (task % G) * PPW + pl;
        const long long pix = (long long)tile * kTilePixels + p;
        const bool owner = pix < a.n_pixels && j == 0;
        PixelAcc acc;
        acc.slot = acc_slot;
        acc.reset();
#pragma unroll 1
        for (int c = 0; c < C; c++) {
            if (kStage) {
                // ---- the band was staged by a bulk copy issued one pixel-band ago: smem -> registers, then start the next copy
                mbar_wait(bar, parity);
                parity ^= 1u;
#pragma unroll
                for (int i = 0; i < WPL; i++) {
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if ((!GENERIC && i < WPL - 1) || i * G + j < staged_groups)
                        v = lds128(stage_lane + i * (G * kRowBytes));  // a quarter warp reads 128 contiguous bytes
                    A[4 * i + 0] = v.x; A[4 * i + 1] = v.y; A[4 * i + 2] = v.z; A[4 * i + 3] = v.w;
                }
                __syncwarp();  // every lane has read the slab before the next copy may overwrite it
                const bool last = (c == C - 1);
                const int nt = last ? task + n_warps : task;
                if (nt < n_tasks) stage_band(nt, last ? 0 : c + 1);
            } else {
                // ---- the lane's units were prefetched one pixel-band ago: strip -> registers, then start the next prefetch
                cp_async_wait_all();
#pragma unroll
                for (int i = 0; i < WPL; i++) {
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if ((!GENERIC && i < WPL - 1) || i * G + j < a.n_groups) v = lds128(strip_lane + i * 512);
                    A[4 * i + 0] = v.x; A[4 * i + 1] = v.y; A[4 * i + 2] = v.z; A[4 * i + 3] = v.w;
                }
                const bool last = (c == C - 1);
                const int nt = last ? task + n_warps : task;
                if (nt < n_tasks) prefetch_band(nt, last ? 0 : c + 1);
            }
            // once enough pixels of the tile have lost their certificate the tile is finished by the dense per-frame pass below,
            // which is exact for every pixel: the remaining bands' certificate terms are not computed
            bool skip_cert = false;
            if (G == 1 && MODE != 2 && c > 0 && a.inline_min > 0)
                skip_cert = __popc(__ballot_sync(0xffffffffu, owner && !acc.hard && !(acc.bound * 1.0001f < a.thr_sq))) >= a.inline_min;
            process_band<C, WPL, G, MODE, true>(a, A, c, j, pix, owner, cap, pad, acc, skip_cert);
        }
        // ---- hard pixels wait in the warp's queue until a warp-full can run the iterative solver together
        // iterative tier: an order statistic outside its window, or a pixel the IQR bound could not clear
        const bool to_hard = acc.hard || (acc.approx && !(acc.bound * 1.0001f < a.thr_sq));
        const bool hard = owner && to_hard;
        const unsigned hb = __ballot_sync(0xffffffffu, hard);
        if (hb) {
            // they are flagged in the tile's word; compact_hard_kernel turns the flags into the iterative tier's queue in tile
            // order and outlier_hard_kernel / outlier_hist_kernel run on it right after, so their code stays out of this loop
            // (owner lanes are lanes 0 .. PPW-1: lane pl holds pixel (task % G) * PPW + pl of the tile)
            if (lane == 0) {
                atomicAdd(a.counters + 2, (unsigned long long)__popc(hb));
                atomicOr(a.hflags + tile, hb << ((task % G) * PPW));
            }
        }
        // ---- tiles where many pixels stay uncertified (noise at the threshold): classify every frame of all 32 pixels right here,
        // while the tile's bytes are still in L2, instead of queueing the pixels one by one (G == 1: lane = pixel of the tile)
        bool finished_here = false;
        if (G == 1 && MODE != 2 && a.inline_min > 0) {
            const bool dirty = owner && !to_hard && !(acc.bound * 1.0001f < a.thr_sq);
            const unsigned db = __ballot_sync(0xffffffffu, dirty);
            if (__popc(db) >= a.inline_min) {
                inline_dense_tile<C>(a, tile, lane, pix, dirty, acc, mask_slots);
                if (lane == 0) atomicAdd(a.counters + 1, (unsigned long long)__popc(db));
                finished_here = dirty;
            }
        }
        finish_pixel<C>(a, acc, pix, p, owner && !to_hard && !finished_here, lane);
        task += n_warps;
    }
}

// Second launch of a compositing call: the pixels queued for the iterative tier, 32/G per warp pass; what their certificate
// cannot clear moves on to the exact-path queue.
template <int C, int WPL, int G, int MODE>
__global__ void __launch_bounds__(kWarpsPerCta * 32, CHB_HARD_MINB) outlier_hard_kernel(const __grid_constant__ OutlierArgs a) {
    constexpr int PPW = 32 / G;
    grid_dependency_wait();
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int cap = 4 * WPL * 4 * G, pad = cap - a.n_sub;
    const uint32_t acc_slot = smem_u32(smem_raw + kBarBytes) + threadIdx.x * 4;
    const MaskSlots ms{smem_u32(smem_raw + kBarBytes + kAccBytes + kWarpsPerCta * (WPL * 512)) + threadIdx.x * 2, kWarpsPerCta * 32u * 2u,
                       smem_u32(smem_raw + kBarBytes + kAccBytes + kWarpsPerCta * (WPL * 512) + kMaskBytes) + threadIdx.x * 16};
    const unsigned int total = a.ghq_count[0];
    const unsigned int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * PPW; base < total; base += n_warps * PPW)
        drain_hard<C, WPL, G, MODE>(a, base, (int)min((unsigned int)PPW, total - base), lane, cap, pad, acc_slot, ms);
    // The pixels the streaming kernel queued itself (certificate failed, median inside its window) sit at the far end of the
    // exact-path queue. With the dense per-frame pass available (mask_path; every pixel of this tier's own warp-fulls is then
    // finished in drain_hard) this kernel takes them as well -- starting with the warps that got no or few iterative-tier
    // pixels -- and the call needs no outlier_exact_kernel launch: one launch latency less per call.
    if (G == 1 && MODE != 2 && a.hard_drains_all) {
        const unsigned int extra = a.gq_count[0];
        const unsigned int w = n_warps - 1u - ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
        for (unsigned int base = w * 32u; base < extra; base += n_warps * 32u) {
            const unsigned int v = base + lane;
            const bool in_range = v < extra;
            drain_queue<C, true>(a, a.gq + (a.n_pixels - 1 - (long long)(in_range ? v : base)), in_range, lane, ms);
        }
    }
}

// Iterative tier for long whole-stack series (hundreds of frames): instead of the solver's repeated passes over the pixel's
// registers, one warp per queued pixel builds a 256-bin histogram of a band in shared memory (one shared-memory atomic per
// sample) and reads every order statistic the band needs -- median pair, quartile pairs, smallest and largest sample --
// off one warp-wide prefix sum (lane l owns bins 8l .. 8l+7). The certificate term uses the exact max |x - centre|.
template <int C>
__global__ void __launch_bounds__(kWarpsPerCta * 32) outlier_hist_kernel(const __grid_constant__ OutlierArgs a) {
    grid_dependency_wait();
    __shared__ __align__(16) uint32_t hist_all[kWarpsPerCta][256];
    __shared__ uint32_t acc_words[kAccWords * kWarpsPerCta * 32];
    const int lane = threadIdx.x & 31, warp_in_cta = threadIdx.x >> 5;
    uint32_t* const hist = hist_all[warp_in_cta];
    for (int b = lane; b < 256; b += 32) hist[b] = 0;
    __syncwarp();
    const unsigned int total = a.hist_all ? (unsigned int)a.n_pixels : a.ghq_count[0];
    if (a.hist_all && blockIdx.x == 0 && threadIdx.x == 0) a.ghq_count[0] = total;  // the per-frame path finds every slot of its queue mirrored
    const unsigned int n_warps = (gridDim.x * blockDim.x) >> 5;
    const bool rel = !a.absolute;
    for (unsigned int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; idx < total; idx += n_warps) {
        const long long pix = a.hist_all ? (long long)idx : a.ghq[idx];
        const long long tile = pix >> 5;
        const int p = (int)(pix & 31);
        PixelAcc acc;
        acc.slot = smem_u32(acc_words) + threadIdx.x * 4;
        acc.reset();
#pragma unroll 1
        for (int c = 0; c < C; c++) {
            const float w = a.w[c];
            const uint8_t* ub = a.stack + tile * tile_bytes(C, a.NG) + (((long long)c * a.NG + a.g0) * kTilePixels + p) * kUnitBytes;
            const bool need = (a.bg == 2 || w != 0.0f);
            uint32_t first = 0;
            for (int g = lane; g < a.n_groups; g += 32) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(ub + (long long)g * (kTilePixels * kUnitBytes)));
                if (g == 0) first = v.x & 0xffu;  // window position 0 (whole-stack launch: frame 0)
                if (need) {  // all sixteen bytes: the frames a last group lacks are zero bytes, taken out of bin 0 below
                    const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int q = 0; q < 4; q++) {
#pragma unroll
                        for (int k = 0; k < 4; k++) atomicAdd(&hist[(wv[q] >> (8 * k)) & 0xffu], 1u);
                    }
                }
            }
            __syncwarp();
            if (need && lane == 0) hist[0] -= (uint32_t)(a.n_groups * kGroupFrames - a.n);
            first = __shfl_sync(0xffffffffu, first, 0);
            acc.first_px |= first << (8 * c);
            __syncwarp();
            if (!need) continue;
            // ---- this lane's eight bins, then the warp-wide prefix sum
            uint32_t h[8];
            {
                const uint4 h0 = *reinterpret_cast<const uint4*>(hist + 8 * lane), h1 = *reinterpret_cast<const uint4*>(hist + 8 * lane + 4);
                h[0] = h0.x; h[1] = h0.y; h[2] = h0.z; h[3] = h0.w; h[4] = h1.x; h[5] = h1.y; h[6] = h1.z; h[7] = h1.w;
                *reinterpret_cast<uint4*>(hist + 8 * lane) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(hist + 8 * lane + 4) = make_uint4(0, 0, 0, 0);
            }
            uint32_t lane_tot = 0, lane_sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { lane_tot += h[k]; lane_sum += (uint32_t)(8 * lane + k) * h[k]; }
            uint32_t incl = lane_tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t excl = incl - lane_tot;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lane_sum += __shfl_xor_sync(0xffffffffu, lane_sum, o);
            acc.set_sum(c, lane_sum);
            __syncwarp();
            if (w == 0.0f) continue;
            uint32_t cum[8];  // cumulative counts at this lane's bins
            {
                uint32_t run = excl;
#pragma unroll
                for (int k = 0; k < 8; k++) { run += h[k]; cum[k] = run; }
            }
            auto stat = [&](int r) {  // the sample of 0-based rank r: the first bin whose cumulative count exceeds r
                const unsigned m = __ballot_sync(0xffffffffu, incl > (uint32_t)r);
                const int src = __ffs(m) - 1;
                int val = 8 * lane;  // in the owning lane: its first bin + the number of its bins still at or below r
#pragma unroll
                for (int k = 0; k < 8; k++) val += (cum[k] <= (uint32_t)r);
                return __shfl_sync(0xffffffffu, val, src);
            };
            const int mlo = stat(a.rk[2]);
            const int mhi = (a.rk[3] == a.rk[2]) ? mlo : stat(a.rk[3]);
            const float med = (mlo == mhi) ? (float)mlo : 0.5f * ((float)mlo + (float)mhi);  // src/chrono.rs:582-591
            const int center = (mlo + mhi) >> 1;
            const float halfw = med - (float)center;
            float iqi = 0.0f, q1 = 0.0f, q3 = 0.0f;
            if (rel) {  // quartiles (src/chrono.rs:559-579) and inverse IQR (:246-252)
                const int q1a = stat(a.rk[0]), q3a = stat(a.rk[4]);
                const int q1b = (a.rk[0] == a.rk[1]) ? q1a : stat(a.rk[1]);
                const int q3b = (a.rk[4] == a.rk[5]) ? q3a : stat(a.rk[5]);
                q1 = (a.rk[0] == a.rk[1]) ? (float)q1a : (1.0f - a.q1_frac) * (float)q1a + a.q1_frac * (float)q1b;
                q3 = (a.rk[4] == a.rk[5]) ? (float)q3a : (1.0f - a.q3_frac) * (float)q3a + a.q3_frac * (float)q3b;
                float iq = q3 - q1;
                if (iq == 0.0f) iq = 1.0f;
                iqi = 1.0f / iq;
            }
            acc.set_median(c, med);
            acc.set_iqr_inv(c, iqi);
            if (a.dbg_median && lane == 0) {  // per-band sub-results (planes are zeroed by the host)
                a.dbg_median[pix * 4 + c] = med;
                if (a.dbg_q1) a.dbg_q1[pix * 4 + c] = q1;
                if (a.dbg_q3) a.dbg_q3[pix * 4 + c] = q3;
            }
            if (!(w < 0.0f)) {  // exact max |x - centre| from the smallest and the largest sample
                const unsigned nz = __ballot_sync(0xffffffffu, lane_tot > 0);
                int my_lo = 8 * lane + 7, my_hi = 8 * lane;
#pragma unroll
                for (int k = 7; k >= 0; k--)
                    if (h[k] > 0) my_lo = 8 * lane + k;
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (h[k] > 0) my_hi = 8 * lane + k;
                const int minv = __shfl_sync(0xffffffffu, my_lo, __ffs(nz) - 1);
                const int maxv = __shfl_sync(0xffffffffu, my_hi, 31 - __clz(nz));
                const int odev = max(center - minv, maxv - center);
                const float aw = a.absolute ? w : w * iqi;
                const float t = aw * ((float)odev + halfw);
                acc.bound += t * t;
            }
        }
        __syncwarp();
        finish_pixel<C>(a, acc, pix, p, lane == 0, lane, (long long)idx);
        __syncwarp();
    }
}

// Last launch of a compositing call: the queued pixels, 32 per warp.
template <int C>
__global__ void __launch_bounds__(256, CHB_EXACT_MINB) outlier_exact_kernel(const __grid_constant__ OutlierArgs a) {
    grid_dependency_wait();
    __shared__ unsigned short mask_words[kMaskGroups][256];  // per-thread outlier masks of the dense pass
    __shared__ __align__(16) uint8_t ring[dense_ring_bytes(C, 256)];  // ... and its ring of frame groups
    const MaskSlots ms{smem_u32(&mask_words[0][threadIdx.x]), 256u * 2u, smem_u32(ring) + threadIdx.x * 16};
    const unsigned int mirrored = a.ghq_count[0], total = mirrored + a.gq_count[0];
    const int lane = threadIdx.x & 31;
    const unsigned int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u; base < total; base += n_warps * 32u) {
        const unsigned int v = base + lane;
        const bool in_range = v < total;
        const long long slot = !in_range ? 0 : (v < mirrored ? (long long)v : a.n_pixels - 1 - (long long)(v - mirrored));
        drain_queue<C>(a, a.gq + slot, in_range, lane, ms);
    }
}

// Between the streaming kernel and the iterative tier: the per-tile flag words become that tier's queue. A block takes 1024
// consecutive tiles and writes their flagged pixels in tile order (blocks land in the order of their atomicAdd), so entries
// that are neighbours in the queue are neighbours in the image.
// The flag words are cleared as they are read, so the next call finds them zero without a per-call memset of the whole array.
// It also clears the counter set the NEXT call on this slot will use (the host alternates between two sets), so that a call
// starts without a memset.
__global__ void __launch_bounds__(256) compact_hard_kernel(uint32_t* __restrict__ flags, long long n_tiles, long long* __restrict__ ghq,
                                                           unsigned int* __restrict__ ghq_count, unsigned int* __restrict__ next_set) {
    constexpr int kTilesPerThread = 4;
    __shared__ uint32_t warp_tot[8];
    __shared__ uint32_t block_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long n_chunks = (n_tiles + 256 * kTilesPerThread - 1) / (256 * kTilesPerThread);
    if (blockIdx.x == 0 && threadIdx.x < 16) next_set[threadIdx.x] = 0u;
    for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const long long t0 = (chunk * 256 + threadIdx.x) * kTilesPerThread;
        uint32_t f[kTilesPerThread];
        uint32_t cnt = 0;
#pragma unroll
        for (int k = 0; k < kTilesPerThread; k++) {
            f[k] = (t0 + k < n_tiles) ? flags[t0 + k] : 0u;
            if (f[k]) flags[t0 + k] = 0u;
            cnt += (uint32_t)__popc(f[k]);
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        uint32_t woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            if (w < wid) woff += warp_tot[w];
            tot += warp_tot[w];
        }
        if (threadIdx.x == 0) block_base = tot ? atomicAdd(ghq_count, tot) : 0u;
        __syncthreads();
        if (cnt) {
            uint32_t off = block_base + woff + incl - cnt;
#pragma unroll
            for (int k = 0; k < kTilesPerThread; k++) {
                uint32_t m = f[k];
                while (m) {
                    const int b = __ffs(m) - 1;
                    ghq[off++] = (t0 + k) * kTilePixels + b;
                    m &= m - 1;
                }
            }
        }
        __syncthreads();  // warp_tot / block_base are reused by the next chunk
    }
}

// ------------------------------------------------------------------------------------------------ K3 chrono-video
// create_video (src/main.rs:230-331) composites one window per output frame; the windows of a `--video-in a/b/1` run have
// the same length and start one frame apart. One launch composites a whole run of such windows: a warp takes a tile and a
// block of 16 consecutive window starts and loads the frame groups those windows span ONCE (lane = pixel). Phase 1 goes
// band by band with the band's bytes in registers: 16 packed counts #{x <= p + k} around the median summarise the window
// and are updated incrementally as it slides (see "sliding counts" below); median pair, quartiles and the smallest /
// largest sample -- hence the exact certificate term max |x - centre| -- are read off the counts. Counts are rebuilt from
// scratch (VABSDIFF4 F evaluations) at the start of a block and, after the iterative solver, for pixels whose order
// statistics left the counted values. Per (window, band) results are parked in shared memory. Phase 2 goes window by
// window: certificate, background pixel, or the exact-path queue (exact_pixel, 32 queued pixel-windows at a time).
#ifndef CHB_VIDEO_MINB
#define CHB_VIDEO_MINB 3
#endif
constexpr int kVideoWarps = 4;
constexpr int kVideoQueueCap = 64;
constexpr int kVideoBlock = 16;  // window starts per task
struct VideoArgs {
    OutlierArgs o;            // what every window shares: n, ranks, thresholds, weights, policies, fade, seed, stack, counters
    int first_start;          // window i covers frames [first_start + i, first_start + i + o.n)
    int n_windows;
    int blk0, n_blocks;       // blocks of 16 starts: block b holds starts [16 * (blk0 + b), 16 * (blk0 + b) + 16)
    int res_words;            // result words per band and window: 1, or 2 when sums / inter-quartile ranges are needed
    long long out_stride;     // bytes between the planes of consecutive windows
    uint8_t* out_images;      // [n_windows][n_pixels * C]
    uint8_t* out_masks;       // may be null
    unsigned long long* win_warnings;  // [n_windows] all-outlier pixels per window
    uint32_t mask_a, mask_b;  // byte masks of window words NW-2 and NW-1 (0xFF = position < n)
    struct VideoQueueEntry* gq;  // global exact-path queue of this launch, drained by video_exact_kernel
    unsigned int* gq_count;   // [0] entries requested, [1] end of the entries that were written (first refused request)
    unsigned int gq_cap;
};
struct VideoQueueEntry {
    long long pix;
    int win, unused;
    float median[4];
    float iqr_inv[4];
    uint32_t sum[4];
};
__host__ __device__ constexpr int video_smem_bytes(int C, int res_words) {
    return kVideoWarps * kVideoQueueCap * (int)sizeof(VideoQueueEntry) + kVideoBlock * C * res_words * kVideoWarps * 32 * 4;
}

template <int NW>
__device__ __noinline__ void video_solve(const uint32_t* xl, int guess, const OutlierArgs& a, int pad, int& med2, int& iq4) {
    uint32_t x[NW];
#pragma unroll
    for (int q = 0; q < NW; q++) x[q] = xl[q];
    BandRanks r;
    band_solve<NW, 1>(x, (uint32_t)(guess * a.n_sub), a.inv_n_sub, a, pad, 4 * NW, r);  // the solver only derives its first guess from the sum
    med2 = r.mlo + r.mhi;
    iq4 = 0;
    if (!a.absolute) {  // quartiles (src/chrono.rs:559-579) are multiples of 1/4: 4 * (q3 - q1) is an exact integer
        const float q1 = (a.rk[0] == a.rk[1]) ? (float)r.q1a : (1.0f - a.q1_frac) * (float)r.q1a + a.q1_frac * (float)r.q1b;
        const float q3 = (a.rk[4] == a.rk[5]) ? (float)r.q3a : (1.0f - a.q3_frac) * (float)r.q3a + a.q3_frac * (float)r.q3b;
        iq4 = __float2int_rn((q3 - q1) * 4.0f);
    }
}
// 1 / IQR from 4 * IQR (src/chrono.rs:246-252: an IQR of zero counts as one)
__device__ __forceinline__ float iqr_inv_of(int iq4) { return iq4 == 0 ? 1.0f : 1.0f / ((float)iq4 * 0.25f); }

// ---- sliding counts. A band's window is summarised by 16 counts cn_k = #{x <= p + k}, k = 0..15, packed one per byte in
// four words (window lengths <= 64 fit a byte). Sliding the window by one frame adds the entering sample's and removes the
// leaving sample's "<=" indicator -- a few integer instructions instead of a pass over the window -- and every order
// statistic whose value lies in [p + 1, p + 15] is read off the counts: d[r] = p + #{k : cn_k <= r}.
struct VideoCounts {
    uint32_t c0, c1, c2, c3;
    int p;
};
__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, int s) {  // shifts of 32 and more give 0
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
    return r;
}
// byte k of word j = 1 iff x <= p + 4j + k
__device__ __forceinline__ void le_masks(int x, int p, uint32_t (&m)[4]) {
    const int d8 = (x - p) * 8;
    m[0] = shl_clamp(0x01010101u, max(d8, 0));
    m[1] = shl_clamp(0x01010101u, __viaddmax_s32(d8, -32, 0));
    m[2] = shl_clamp(0x01010101u, __viaddmax_s32(d8, -64, 0));
    m[3] = shl_clamp(0x01010101u, __viaddmax_s32(d8, -96, 0));
}
// p + #{k : cn_k <= r} with kk = rep4(127 - r): adding 127 - r sets bit 7 of a count byte iff the count exceeds r
__device__ __forceinline__ int stat_at(const VideoCounts& v, uint32_t kk) {
    uint32_t acc0 = __dp4a((v.c0 + kk) & 0x80808080u, 0x01010101u, 0u);
    uint32_t acc1 = __dp4a((v.c1 + kk) & 0x80808080u, 0x01010101u, 0u);
    acc0 = __dp4a((v.c2 + kk) & 0x80808080u, 0x01010101u, acc0);
    acc1 = __dp4a((v.c3 + kk) & 0x80808080u, 0x01010101u, acc1);
    return v.p + 16 - (int)((acc0 + acc1) >> 7);
}
// Counts from scratch around `center`: F at the 17 values p .. p + 16 (VABSDIFF4.ACC), cn_k = (F(p+k+1) - F(p+k) + cap) / 2
// minus the zero bytes that pad the window words.
template <int NW>
__device__ __noinline__ VideoCounts video_recount(const uint32_t* xl, int center, int n) {
    uint32_t x[NW];
#pragma unroll
    for (int q = 0; q < NW; q++) x[q] = xl[q];
    int p = center - 8;
    p = p < 0 ? 0 : (p > 240 ? 240 : p);
    uint32_t f[17], cc[17];
#pragma unroll
    for (int k = 0; k < 17; k++) { cc[k] = rep4(min(p + k, 255)); f[k] = 0; }
#pragma unroll
    for (int q = 0; q < NW; q++) {
#pragma unroll
        for (int k = 0; k < 17; k++) f[k] = sad4_acc(x[q], cc[k], f[k]);
    }
    constexpr int cap = 4 * NW;
    const int pad = cap - n;
    uint32_t cn[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        cn[k] = (uint32_t)((((int)f[k + 1] - (int)f[k] + cap) >> 1) - pad);
        if (p + k >= 255) cn[k] = (uint32_t)n;  // F(256) does not exist: every byte is <= 255
    }
    VideoCounts r;
    r.p = p;
    r.c0 = cn[0] | (cn[1] << 8) | (cn[2] << 16) | (cn[3] << 24);
    r.c1 = cn[4] | (cn[5] << 8) | (cn[6] << 16) | (cn[7] << 24);
    r.c2 = cn[8] | (cn[9] << 8) | (cn[10] << 16) | (cn[11] << 24);
    r.c3 = cn[12] | (cn[13] << 8) | (cn[14] << 16) | (cn[15] << 24);
    return r;
}

// 4 * IQR of the window from its counts (quartiles, src/chrono.rs:559-579, are multiples of 1/4)
__device__ __forceinline__ int video_iq4(const OutlierArgs& a, const VideoCounts& cn) {
    const int q1a = stat_at(cn, rep4(127 - a.rk[0])), q3a = stat_at(cn, rep4(127 - a.rk[4]));
    const int q1b = (a.rk[0] == a.rk[1]) ? q1a : stat_at(cn, rep4(127 - a.rk[1]));
    const int q3b = (a.rk[4] == a.rk[5]) ? q3a : stat_at(cn, rep4(127 - a.rk[5]));
    const float q1 = (a.rk[0] == a.rk[1]) ? (float)q1a : (1.0f - a.q1_frac) * (float)q1a + a.q1_frac * (float)q1b;
    const float q3 = (a.rk[4] == a.rk[5]) ? (float)q3a : (1.0f - a.q3_frac) * (float)q3a + a.q3_frac * (float)q3b;
    return __float2int_rn((q3 - q1) * 4.0f);
}

// The uncommon cases of one (window, band), the whole warp together: a rank outside the counted values -> iterative solver,
// then counts around the exact median; samples beyond the counted values -> the window is scanned for the exact maximum
// deviation (or, for one or two such pixels, they get the pessimistic bound and take the exact path: most of them hold an
// outlier anyway). X: the window's words, positions >= n zeroed.
struct VideoSlowIO {  // (lives in local memory, and only on the uncommon path: the caller copies in and out)
    VideoCounts cn;
    int med2, iq4;
    uint32_t odev;
};
template <int NW>
__device__ __noinline__ void video_window_slow(const VideoArgs& v, const uint32_t* X, float w, bool live, int lane, int pad, VideoSlowIO* io) {
    const OutlierArgs& a = v.o;
    const int n = a.n;
    VideoCounts cn = io->cn;
    const int r_lo = a.absolute ? a.rk[2] : a.rk[0], r_hi = a.absolute ? a.rk[3] : a.rk[5];  // smallest / largest rank a band needs
    const bool ok = ((int)(cn.c0 & 0xffu) <= r_lo || cn.p == 0) && ((int)(cn.c3 >> 24) > r_hi);
    int med2 = 0, iq4 = 0;
    uint32_t odev = 0;
    bool solved = false;
    const unsigned nb = __ballot_sync(0xffffffffu, !ok && live);
    if (nb) {
        if (lane == 0) atomicAdd(a.counters + 2, (unsigned long long)__popc(nb));
        // A median leaves the counted values when an object has come to cover (or has just uncovered) more than half of the
        // window: the newest sample then belongs to the cluster that now holds the median. Counts around it usually resolve
        // every rank (they are exact wherever they are centred); only if they do not, the iterative solver runs.
        const uint32_t newest = (X[(n - 1) >> 2] >> (8 * ((n - 1) & 3))) & 0xffu;
        VideoCounts rc = video_recount<NW>(X, (int)newest, n);
        const bool ok2 = ((int)(rc.c0 & 0xffu) <= r_lo || rc.p == 0) && ((int)(rc.c3 >> 24) > r_hi);
        if (__any_sync(0xffffffffu, !ok && live && !ok2)) {
            int m2 = 0, i4 = 0;
            video_solve<NW>(X, min(cn.p + 8, 254), a, pad, m2, i4);
            rc = video_recount<NW>(X, m2 >> 1, n);
            if (!ok) { cn = rc; med2 = m2; iq4 = i4; solved = true; }
        } else if (!ok) {
            cn = rc;
        }
    }
    if (!solved) {
        const int mlo = stat_at(cn, rep4(127 - a.rk[2]));
        const int mhi = (a.rk[2] == a.rk[3]) ? mlo : stat_at(cn, rep4(127 - a.rk[3]));
        med2 = mlo + mhi;
        if (!a.absolute) iq4 = video_iq4(a, cn);
    }
    if (!(w < 0.0f)) {
        const int center = med2 >> 1;
        const bool in_range = ((cn.c0 & 0xffu) == 0u || cn.p == 0) && ((int)(cn.c3 >> 24) == n);
        const int minv = stat_at(cn, rep4(127)), maxv = stat_at(cn, rep4(127 - (n - 1)));
        odev = (uint32_t)max(center - minv, maxv - center);
        const unsigned ob = __ballot_sync(0xffffffffu, !in_range && live);
        if (__popc(ob) >= 3) {  // many pixels with samples beyond the counted values: scan the window
            const uint32_t cc = rep4(center);
            uint32_t m0 = 0, m1 = 0;
#pragma unroll
            for (int q = 0; q < NW; q += 2) {
                uint32_t x0 = X[q], x1 = X[q + 1];
                if (q == NW - 2) { x0 |= cc & ~v.mask_a; x1 |= cc & ~v.mask_b; }
                const uint32_t d0 = absdiff4(x0, cc), d1 = absdiff4(x1, cc);
                m0 = __vmaxu2(m0, d0); m0 = __vmaxu2(m0, d0 << 8);  // the high byte of each half is a byte-wise maximum
                m1 = __vmaxu2(m1, d1); m1 = __vmaxu2(m1, d1 << 8);
            }
            const uint32_t m = __vmaxu2(m0, m1);
            if (!in_range) odev = max((m >> 8) & 0xffu, m >> 24);
        } else if (!in_range) {
            odev = 255;
        }
    }
    io->cn = cn; io->med2 = med2; io->iq4 = iq4; io->odev = odev;
}

template <int C>
__device__ __noinline__ void drain_video_queue(const VideoArgs& v, const VideoQueueEntry* q, int count, int lane) {
    const OutlierArgs& a = v.o;
    const bool active = lane < count;
    const VideoQueueEntry e = q[active ? lane : 0];
    const PixelSrc src{a.stack + (e.pix >> 5) * tile_bytes(C, a.NG), a.NG, C, (int)(e.pix & 31)};
    const int f0 = v.first_start + e.win;  // frame of window position 0 = the window's frame_offset (src/chrono.rs:102-103)
    uint8_t pixel[4] = {0, 0, 0, 0};
    int n_out = 0, warn = 0;
    const uint8_t mask = exact_pixel(a, src, pixel_gid(a, e.pix), e.median, e.iqr_inv, e.sum, pixel, n_out, warn, f0, f0);
    if (active) {
        uint8_t* oi = v.out_images + (long long)e.win * v.out_stride + e.pix * C;
#pragma unroll
        for (int c = 0; c < C; c++) oi[c] = pixel[c];
        if (v.out_masks) {
            uint8_t* om = v.out_masks + (long long)e.win * v.out_stride + e.pix * C;
#pragma unroll
            for (int c = 0; c < C; c++) om[c] = (c < 3) ? mask : 255;  // src/chrono.rs:183-191
        }
        if (warn) atomicAdd(v.win_warnings + e.win, 1ULL);
    }
    if (lane == 0) atomicAdd(a.counters + 1, (unsigned long long)count);
}

// Pixel-windows the certificate cannot clear go to the launch's global queue and are finished by video_exact_kernel, so
// the (large) exact-path code stays out of video_kernel's instruction stream. Only when that queue is full does the warp
// fall back to its own shared-memory queue and drain it in place. (Result words passed by value: no local memory.)
template <int C>
__device__ __noinline__ void video_enqueue(const VideoArgs& v, VideoQueueEntry* queue, int& qcount, int lane, unsigned db, bool dirty, long long pix,
                                           int win, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
    const OutlierArgs& a = v.o;
    const int nd = __popc(db);
    unsigned int gbase = 0;
    if (lane == 0) gbase = atomicAdd(v.gq_count, (unsigned int)nd);
    gbase = __shfl_sync(0xffffffffu, gbase, 0);
    const bool to_global = gbase + (unsigned int)nd <= v.gq_cap;
    if (!to_global && lane == 0) atomicMin(v.gq_count + 1, gbase);
    if (!to_global && qcount + nd > kVideoQueueCap) {
        __syncwarp();
        while (qcount >= 32) { drain_video_queue<C>(v, queue + (qcount - 32), 32, lane); qcount -= 32; }
        __syncwarp();
    }
    if (dirty) {
        const int rank = __popc(db & ((1u << lane) - 1u));
        VideoQueueEntry& e = to_global ? v.gq[gbase + rank] : queue[qcount + rank];
        const uint32_t w0[4] = {a0, a1, a2, a3}, w1[4] = {b0, b1, b2, b3};
        e.pix = pix; e.win = win; e.unused = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            e.median[c] = 0.5f * (float)(w0[c] & 0x1ffu);
            e.iqr_inv[c] = (c < C && !a.absolute && a.w[c] != 0.0f) ? iqr_inv_of((int)(w1[c] >> 14)) : 0.0f;
            e.sum[c] = w1[c] & 0x3fffu;
        }
    }
    if (!to_global) {
        qcount += nd;
        __syncwarp();
        while (qcount >= 32) { drain_video_queue<C>(v, queue + (qcount - 32), 32, lane); qcount -= 32; }
        __syncwarp();
    }
}

// Phase 2 of a chrono-video task, window by window over the per-(window, band) result words parked in shared memory:
// certificate, then the background pixel -- or the exact-path queue.
template <int C>
__device__ __forceinline__ void video_phase2(const VideoArgs& v, const uint32_t* res, int rw, int i_lo, int i_hi, int blk, int tile, long long pix, bool owner,
                                             int lane, VideoQueueEntry* queue, int& qcount) {
    constexpr int kThreads = kVideoWarps * 32;
    const OutlierArgs& a = v.o;
    const int n = a.n;
    // ---- phase 2, window by window: certificate, background pixel or exact-path queue
#pragma unroll 1
    for (int i = i_lo; i < i_hi; i++) {
        const int win = blk * kVideoBlock + i - v.first_start;
        uint32_t r0[C], r1[C];
        bool clean;
        if (a.int_dist) {
            // absolute thresholds, weights 0 / 1: |2 x - 2 median| <= 2 max|x - centre| + (median - centre) * 2 in integers, and
            // 4 dist_sq < thr4 <=> dist_sq < thr_sq exactly (see IntDist), so the certificate is an integer comparison
            uint32_t acc = 0;
#pragma unroll
            for (int c = 0; c < C; c++) {
                const uint32_t* r = res + ((i * C + c) * rw) * kThreads;
                r0[c] = r[0];
                r1[c] = rw > 1 ? r[kThreads] : 0u;
                const uint32_t t = a.w[c] != 0.0f ? 2u * ((r0[c] >> 9) & 0xffu) + (r0[c] & 1u) : 0u;
                acc += t * t;
            }
            clean = (int)acc < a.thr4;
        } else {
            float bound = 0.0f;
#pragma unroll
            for (int c = 0; c < C; c++) {
                const uint32_t* r = res + ((i * C + c) * rw) * kThreads;
                r0[c] = r[0];
                r1[c] = rw > 1 ? r[kThreads] : 0u;
                const float w = a.w[c];
                if (w != 0.0f && !(w < 0.0f)) {
                    const float aw = a.absolute ? w : w * iqr_inv_of((int)(r1[c] >> 14));
                    const float t = aw * ((float)((r0[c] >> 9) & 0xffu) + 0.5f * (float)(r0[c] & 1u));
                    bound += t * t;
                }
            }
            clean = bound * 1.0001f < a.thr_sq;  // margin covers the f32 roundings of the reference's sum
        }
        if (owner && clean) {
            uint8_t pixel[4] = {0, 0, 0, 0};
            if (a.bg == 2) {
#pragma unroll
                for (int c = 0; c < C; c++) pixel[c] = sat_u8(roundf((float)(r1[c] & 0x3fffu) / (float)n));  // src/chrono.rs:297-306,335-337
            } else if (a.bg == 3) {
#pragma unroll
                for (int c = 0; c < C; c++) pixel[c] = sat_u8(roundf(0.5f * (float)(r0[c] & 0x1ffu)));  // :340-345
            } else if (a.bg == 0) {
#pragma unroll
                for (int c = 0; c < C; c++) pixel[c] = (uint8_t)((r0[c] >> 17) & 0xffu);  // :348-350: window position 0
            } else {
                const int pos = (int)rng_range(a.seed, pixel_gid(a, pix), 0, (uint32_t)n);  // :357
                const PixelSrc src{a.stack + (long long)tile * tile_bytes(C, a.NG), a.NG, C, lane};
#pragma unroll
                for (int c = 0; c < C; c++) pixel[c] = src.at(v.first_start + win + pos, c);
            }
            uint8_t* oi = v.out_images + (long long)win * v.out_stride + pix * C;
#pragma unroll
            for (int c = 0; c < C; c++) oi[c] = pixel[c];
            if (v.out_masks) {
                uint8_t* om = v.out_masks + (long long)win * v.out_stride + pix * C;
#pragma unroll
                for (int c = 0; c < C; c++) om[c] = (c < 3) ? 0 : 255;
            }
        }
        const bool dirty = owner && !clean;
        const unsigned db = __ballot_sync(0xffffffffu, dirty);
        if (db) {
            uint32_t w0[4] = {0, 0, 0, 0}, w1[4] = {0, 0, 0, 0};
#pragma unroll
            for (int c = 0; c < C; c++) { w0[c] = r0[c]; w1[c] = r1[c]; }
            video_enqueue<C>(v, queue, qcount, lane, db, dirty, pix, win, w0[0], w0[1], w0[2], w0[3], w1[0], w1[1], w1[2], w1[3]);
        }
    }
}

// Result word 0 of a (window, band): bits 0-8 mlo + mhi (twice the median), 9-16 max |x - centre|, 17-24 the byte of
// window position 0. Word 1 (when present): bits 0-13 band sum, 14-23 4 * IQR.
template <int C, int NW>
__global__ void __launch_bounds__(kVideoWarps * 32, CHB_VIDEO_MINB) video_kernel(const __grid_constant__ VideoArgs v) {
    constexpr int KG = (NW + 3) / 4 + 1;  // frame groups a block of 16 starts spans
    constexpr int NWL = 4 * KG;           // words per lane
    constexpr int kThreads = kVideoWarps * 32;
    const OutlierArgs& a = v.o;
    extern __shared__ __align__(16) uint8_t vsm[];
    const int lane = threadIdx.x & 31, warp_in_cta = threadIdx.x >> 5;
    VideoQueueEntry* const queue = reinterpret_cast<VideoQueueEntry*>(vsm) + warp_in_cta * kVideoQueueCap;
    uint32_t* const res = reinterpret_cast<uint32_t*>(vsm + kVideoWarps * kVideoQueueCap * sizeof(VideoQueueEntry)) + threadIdx.x;
    const int rw = v.res_words;
    const int n_warps = (int)((gridDim.x * blockDim.x) >> 5);
    const int n_tasks = (int)a.n_tiles * v.n_blocks;  // the host keeps this below 2^31
    constexpr int cap = 4 * NW;
    const int n = a.n;
    const int pad = cap - n;
    const bool rel = !a.absolute;
    // rank constants of stat_at(): median pair, quartile pairs, smallest and largest sample
    const uint32_t kk_m1 = rep4(127 - a.rk[2]), kk_m2 = rep4(127 - a.rk[3]);
    const uint32_t kk_min = rep4(127), kk_max = rep4(127 - (n - 1));
    const int r_lo = rel ? a.rk[0] : a.rk[2], r_hi = rel ? a.rk[5] : a.rk[3];  // smallest / largest rank a band needs
    const bool one_rank = a.rk[2] == a.rk[3];
    const int n_word = n >> 2, n_shift = 8 * (n & 3);  // window position n: word NW-2 .. NW of the window's words, and its byte
    int qcount = 0;
    for (int task = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); task < n_tasks; task += n_warps) {
        const int tile = task / v.n_blocks;
        const int blk = v.blk0 + task % v.n_blocks;
        const long long pix = (long long)tile * kTilePixels + lane;
        const bool owner = pix < a.n_pixels;
        const uint8_t* tb = a.stack + (long long)tile * tile_bytes(C, a.NG) + lane * kUnitBytes;
        const int i_lo = max(0, v.first_start - blk * kVideoBlock);                            // windows of this block that belong to the run
        const int i_hi = min(kVideoBlock, v.first_start + v.n_windows - blk * kVideoBlock);
        // windows of the run (bit i), none for lanes beyond the band's last pixel: only those vote for the uncommon path
        const uint32_t live_mask = (owner && i_hi > i_lo) ? (((1u << i_hi) - 1u) & ~((1u << i_lo) - 1u)) : 0u;  // 0 <= i_lo < i_hi <= 16
        // ---- phase 1, band by band: the band's bytes in registers, the window slides over them. All 16 windows of the block
        // are evaluated (phase 2 only reads those of the run); the words rotate once per four windows. The common case -- every
        // rank and the smallest / largest sample inside the counted values -- is straight-line code behind ONE warp vote;
        // everything else (solver, re-centred counts, window scan) lives in video_window_slow.
#pragma unroll 1
        for (int c = 0; c < C; c++) {
            uint32_t A[NWL];
#pragma unroll
            for (int k = 0; k < KG; k++) {
                uint4 u = make_uint4(0, 0, 0, 0);
                if (blk + k < a.NG) u = ldg_stream(tb + ((long long)c * a.NG + (blk + k)) * (kTilePixels * kUnitBytes));
                A[4 * k] = u.x; A[4 * k + 1] = u.y; A[4 * k + 2] = u.z; A[4 * k + 3] = u.w;
            }
            const float w = a.w[c];
            const bool stats = (w != 0.0f);
            const bool want_dev = stats && !(w < 0.0f);
            VideoCounts cn = {0, 0, 0, 0, 0};
            uint32_t bsum = 0;
            if (a.bg == 2 || stats) {  // window 0 of the block: sum and counts from scratch
                uint32_t X[NW];
#pragma unroll
                for (int q = 0; q < NW; q++) X[q] = A[q];
                X[NW - 2] &= v.mask_a;
                X[NW - 1] &= v.mask_b;
                uint32_t s0 = 0, s1 = 0;
#pragma unroll
                for (int q = 0; q < NW; q += 2) { s0 = __dp4a(X[q], 0x01010101u, s0); s1 = __dp4a(X[q + 1], 0x01010101u, s1); }
                bsum = s0 + s1;
                if (stats) cn = video_recount<NW>(X, __float2int_rn((float)bsum * a.inv_n_sub), n);
            }
            // ONE copy of the window body in the instruction stream (the four-fold unrolled version, with the window's byte offset a
            // compile-time constant, was 18 KB of loop and spent 59 % of its stall samples waiting for instructions): two nested
            // loops, the words rotate in the outer one
#pragma unroll 1
            for (int quad = 0; quad < kVideoBlock / 4; quad++) {
                // the four samples that enter during this quad's slides (window positions n .. n + 3 of its first window)
                const uint32_t wlo = n_word == NW - 2 ? A[NW - 2] : (n_word == NW - 1 ? A[NW - 1] : A[NW]);
                const uint32_t whi = n_word == NW - 2 ? A[NW - 1] : (n_word == NW - 1 ? A[NW] : A[NW + 1]);
                const uint32_t in4 = __funnelshift_r(wlo, whi, n_shift);
                const uint32_t live_quad = live_mask >> (4 * quad);
                uint32_t* const res_quad = res + (4 * quad * C + c) * rw * kThreads;
#pragma unroll 1
                for (int bo = 0; bo < 4; bo++) {  // bo: byte of A[0] the window starts at
                const uint32_t x_first = __byte_perm(A[0], 0, 0x4440 + bo);
                int med2 = 0, iq4 = 0;
                uint32_t odev = 0;
                if (stats) {
                    const int cn_lo = (int)(cn.c0 & 0xffu), cn_hi = (int)(cn.c3 >> 24);
                    // every rank the band needs must resolve inside the counted values (the byte-range ends count as known),
                    // and so must the smallest and the largest sample for the exact certificate term -- which implies the former
                    const bool fine = want_dev ? ((cn_lo == 0 || cn.p == 0) && cn_hi == n) : ((cn_lo <= r_lo || cn.p == 0) && cn_hi > r_hi);
                    const bool live = (live_quad >> bo) & 1u;
                    if (__any_sync(0xffffffffu, live && !fine)) {
                        uint32_t X[NW];
#pragma unroll
                        for (int q = 0; q < NW; q++) X[q] = __funnelshift_r(A[q], A[q + 1], 8 * bo);
                        X[NW - 2] &= v.mask_a;
                        X[NW - 1] &= v.mask_b;
                        VideoSlowIO io;
                        io.cn = cn;
                        video_window_slow<NW>(v, X, w, live, lane, pad, &io);
                        cn = io.cn; med2 = io.med2; iq4 = io.iq4; odev = io.odev;
                    } else {
                        const int mlo = stat_at(cn, kk_m1);
                        const int mhi = one_rank ? mlo : stat_at(cn, kk_m2);
                        med2 = mlo + mhi;
                        if (rel) iq4 = video_iq4(a, cn);
                        if (want_dev) {  // max |x - centre| from the smallest and largest sample
                            const int center = med2 >> 1;
                            const int minv = stat_at(cn, kk_min), maxv = stat_at(cn, kk_max);
                            odev = (uint32_t)max(center - minv, maxv - center);
                        }
                    }
                }
                uint32_t* r = res_quad + bo * (C * rw * kThreads);
                r[0] = (uint32_t)med2 | (odev << 9) | (x_first << 17);
                if (rw > 1) r[kThreads] = bsum | ((uint32_t)iq4 << 14);
                // ---- slide by one frame: the sample at window position n enters, position 0 leaves
                const uint32_t x_in = __byte_perm(in4, 0, 0x4440 + bo);
                bsum += x_in - x_first;
                if (stats) {
                    uint32_t mi[4], mo[4];
                    le_masks((int)x_in, cn.p, mi);
                    le_masks((int)x_first, cn.p, mo);
                    cn.c0 += mi[0] - mo[0]; cn.c1 += mi[1] - mo[1]; cn.c2 += mi[2] - mo[2]; cn.c3 += mi[3] - mo[3];
                }
                }
#pragma unroll
                for (int q = 0; q < NWL - 1; q++) A[q] = A[q + 1];
                A[NWL - 1] = 0;
            }
        }
        video_phase2<C>(v, res, rw, i_lo, i_hi, blk, tile, pix, owner, lane, queue, qcount);
    }
    __syncwarp();
    if (qcount > 0) drain_video_queue<C>(v, queue, qcount, lane);
}

// Second launch of a chrono-video chunk: the queued pixel-windows, 32 per warp.
template <int C>
__global__ void __launch_bounds__(128) video_exact_kernel(const __grid_constant__ VideoArgs v) {
    const unsigned int total = min(v.gq_count[0], v.gq_count[1]);  // requests that did not fit were finished in place
    const int lane = threadIdx.x & 31;
    const unsigned int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u; base < total; base += n_warps * 32u)
        drain_video_queue<C>(v, v.gq + base, (int)min(32u, total - base), lane);
}

// ------------------------------------------------------------------------------------------------ K4 shake analysis
// ShakeAnalyzer::calc_diffs (src/shake.rs:338-386): for every offset of the search square, the sum of squared differences
// between the first frame's anchor windows and the frame's pixels at that offset, over all anchors and bands; i32 with
// wrap-around like the release build of the reference. Only the (2(r+s)+1)^2 patch around every anchor is on the device.
// One CTA per search offset; its threads stride over the window bytes of every anchor (a row of a window is contiguous in
// both the window and the patch), then reduce.
struct ShakeArgs {
    const uint8_t* windows;  // [n_anchors][size][size * C]
    const uint8_t* patches;  // [n_anchors][psize][psize * C], psize = 2 (r + s) + 1: the frame around every anchor
    int n_anchors, size, psize, search_size, C;
    int32_t* diffs;          // [search_size^2]
    int32_t* result;         // [2]: index of the first minimum, its value
};
__global__ void __launch_bounds__(256) shake_diff_kernel(const ShakeArgs a) {
    const int ox = blockIdx.x % a.search_size, oy = blockIdx.x / a.search_size;
    const int row_bytes = a.size * a.C;
    const int per_anchor = a.size * row_bytes;
    const int total = a.n_anchors * per_anchor;
    uint32_t acc = 0;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int i = e / per_anchor, r = e - i * per_anchor;
        const int dy = r / row_bytes, xb = r - dy * row_bytes;
        const int w = a.windows[e];
        const int p = a.patches[((size_t)i * a.psize + (oy + dy)) * (a.psize * a.C) + ox * a.C + xb];
        const int d = w - p;
        acc += (uint32_t)(d * d);
    }
    __shared__ uint32_t part[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) s += part[k];
        a.diffs[blockIdx.x] = (int32_t)s;
    }
}
// First minimum of the diff table (Iterator::min_by_key keeps the first of equal minima, src/shake.rs:275-276).
__global__ void __launch_bounds__(256) shake_argmin_kernel(const ShakeArgs a) {
    const int n = a.search_size * a.search_size;
    long long best = (long long)0x7fffffff << 32 | 0x7fffffff;  // (value, index) packed so that one min orders by value, then index
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const long long key = ((long long)a.diffs[i] << 32) | (unsigned int)i;
        best = key < best ? key : best;
    }
    __shared__ long long part[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); k++) best = part[k] < best ? part[k] : best;
        if (part[0] < best) best = part[0];
        a.result[0] = (int32_t)(best & 0xffffffffLL);
        a.result[1] = (int32_t)(best >> 32);
    }
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
// src/simple.rs:103-118 (all values are exact integers). The winner's bytes are picked from the chunk's registers.
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
                if (g_base + gg >= a.n_groups) break;  // uniform: the last chunk may be short
                uint4 m = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                if (a.wmask) m = __ldg(reinterpret_cast<const uint4*>(a.wmask) + (g_base + gg));
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
                // the winner's bytes are still in this chunk's registers: select unit, word, byte (no second trip to memory)
                const int gsel = c_frame >> 4, wsel = (c_frame >> 2) & 3, sh = (c_frame & 3) * 8;
#pragma unroll
                for (int c = 0; c < C; c++) {
                    uint4 uu = u[0][c];
#pragma unroll
                    for (int gg = 1; gg < kChunkGroups; gg++) {
                        if (gsel == gg) uu = u[gg][c];
                    }
                    const uint32_t wv = wsel == 0 ? uu.x : (wsel == 1 ? uu.y : (wsel == 2 ? uu.z : uu.w));
                    // bands with weight 0 are not loaded by this kernel: fetch those from memory (rare configuration)
                    if ((a.use_mask >> c) & 1u) outp[c] = (uint8_t)((wv >> sh) & 0xffu);
                    else outp[c] = PixelSrc{a.stack + tile * tbytes, a.NG, C, p}.at((a.g0 + g_base) * kGroupFrames + c_frame, c);
                }
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
                                                         int row0_global, int full_height, int block_rows, int block_skip_rows) {
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
            const int yl = (int)(pix / width), xx = (int)(pix % width);
            const int y = row0_global + yl + (block_rows > 0 ? (yl / block_rows) * block_skip_rows : 0);  // interleaved row-block shards
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
