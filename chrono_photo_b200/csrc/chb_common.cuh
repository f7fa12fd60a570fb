// chb_common.cuh -- helpers shared by host and device code of libchrono_b200 (sm_100a).
// Reference citations are relative to mlange-42/chrono-photo v0.6.5.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define CHB_HD __host__ __device__ __forceinline__

namespace chb {

// ---- stack layout in HBM ("time-sliced", replaces src/slicer.rs temp files) -----------------------------------
// A band of n_pixels pixels (row-major rows of one GPU's shard) is cut into tiles of 32 consecutive pixels.
// For tile t, band (channel) c, frame group g (16 consecutive frames) and pixel p in the tile, one 16-byte unit
// holds that pixel-band's samples of frames 16g .. 16g+15 (byte b = frame 16g+b):
//     unit(t, c, g, p) = base + (((t * C + c) * NG + g) * 32 + p) * 16
// so every tile is one contiguous block of C*NG*512 bytes holding the complete time series of its 32 pixels, and a
// warp reading one (c, g) row moves 512 contiguous bytes. Frames >= n_frames in the last group are zero.
constexpr int kTilePixels = 32;
constexpr int kGroupFrames = 16;
constexpr int kUnitBytes = 16;

CHB_HD long long tile_bytes(int C, int NG) { return (long long)C * NG * kTilePixels * kUnitBytes; }
CHB_HD long long unit_offset(long long tile, int C, int NG, int c, int g, int p) {
    return ((((tile * C + c) * (long long)NG + g) * kTilePixels) + p) * kUnitBytes;
}

// ---- counter-based RNG shared with the oracle (replaces rand::thread_rng, src/chrono.rs:68) -------------------
CHB_HD uint32_t rng_u32(uint64_t seed, uint64_t pixel, uint32_t draw) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (pixel + 1) + 0xD1B54A32D192ED03ULL * (uint64_t)draw;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}
CHB_HD uint32_t rng_range(uint64_t seed, uint64_t pixel, uint32_t draw, uint32_t n) {
    return (uint32_t)(((uint64_t)rng_u32(seed, pixel, draw) * (uint64_t)n) >> 32);
}

// ---- Fade lookup (src/options.rs:113-139) ----------------------------------------------------------------------
struct FadeDev {
    int is_none, mode, absolute, offset, n_values;
    const float* values;  // device pointer inside kernels
};
CHB_HD float fade_get(const FadeDev& f, int frame) {
    if (f.is_none) return 1.0f;
    int i = frame - f.offset;
    int len = f.n_values;
    if (i >= 0 && i < len) return f.values[i];
    if (f.mode == 0) return i < 0 ? f.values[0] : f.values[len - 1];
    while (i < 0) i += len;
    i = i % len;
    return f.values[i];
}
// OutlierProcessor::fade / SimpleProcessor::fade (src/chrono.rs:496-502, src/simple.rs:170-176)
CHB_HD float fade_for(const FadeDev& f, int frame, int total, int offset) {
    return f.absolute ? fade_get(f, offset + frame) : fade_get(f, total - frame - 1);
}

// Rust `f32 as u8`: saturating, NaN -> 0
CHB_HD uint8_t sat_u8(float v) {
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}

// ---- synthetic series (DESIGN.md "Synthetic inputs"); one function for the device generator and the host twin --
CHB_HD uint32_t hash32(uint64_t seed, uint32_t f, uint32_t y, uint32_t x, uint32_t ch) {
    uint64_t z = seed ^ (0x9E3779B97F4A7C15ULL * ((uint64_t)f + 1));
    z ^= ((uint64_t)y << 32) | (uint64_t)x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z += (uint64_t)ch * 0xD1B54A32D192ED03ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}

enum SynthKind { kSynthS1 = 1, kSynthS2 = 2, kSynthUniform = 3, kSynthGauss = 4 };

// kind 1 (S1): recipe of src/util/create_example_data.rs:10-49 without the JPEG round trip: bands 0,1 uniform in
//   [240,250), band 2 uniform in [140,150); a 17x17 square centred (100+10f, H/3+5f) and a fixed one centred
//   (W-24, H-68) with band 0 set to 0 (the reference writes only the first byte of the pixel, :36-38, :44-48).
// kind 2 (S2): smooth gradient + uniform noise in [-5,5] + 8 discs of radius 40 on linear tracks of fixed length.
// kind 3: iid uniform bytes (worst case for the selection search).
// kind 4: gradient + approximately Gaussian noise (sigma ~ 3, sum of 4 uniforms) + the discs of kind 2.
CHB_HD uint8_t synth_byte(int kind, uint64_t seed, int f, int n_frames, int y, int x, int ch, int W, int H) {
    uint32_t h = hash32(seed, (uint32_t)f, (uint32_t)y, (uint32_t)x, (uint32_t)ch);
    if (kind == kSynthUniform) return (uint8_t)(h >> 24);
    if (kind == kSynthS1) {
        int v = (ch == 2 ? 140 : 240) + (int)(((uint64_t)h * 10) >> 32);
        if (ch == 0) {
            int cx = 100 + f * 10, cy = H / 3 + f * 5;
            int dx = x - cx, dy = y - cy;
            if (dx >= -8 && dx <= 8 && dy >= -8 && dy <= 8) v = 0;
            dx = x - (W - 24);
            dy = y - (H - 68);
            if (dx >= -8 && dx <= 8 && dy >= -8 && dy <= 8) v = 0;
        }
        return (uint8_t)v;
    }
    // kinds 2 and 4
    int base = 48 + (int)(((long long)x * 96) / W) + (int)(((long long)y * 64) / H) + 16 * ch;
    int noise;
    if (kind == kSynthS2) {
        noise = (int)(((uint64_t)h * 11) >> 32) - 5;
    } else {
        int s = (int)(h & 7) + (int)((h >> 8) & 7) + (int)((h >> 16) & 7) + (int)((h >> 24) & 7);  // 0..28, sigma ~ 4.6
        noise = s - 14;
    }
    int v = base + noise;
#pragma unroll 1
    for (int k = 0; k < 8; k++) {
        // disc k: start (x0, y0) and a linear track of fixed length over the whole clip (200*(1 + k%3) px in x, +-200 px in y),
        // wrapping around the image: 1..3 px/frame for a 200-frame series, slower for longer clips, so the share of pixels
        // a disc ever touches (~1.3 % at 24 MP) does not depend on the frame count
        long long x0 = ((long long)W * (2 * k + 1)) / 16, y0 = ((long long)H * ((5 * k + 3) % 16)) / 16;
        long long lx = 200LL * (1 + (k % 3)), ly = (k & 1) ? 200LL : -200LL;
        long long nf = n_frames > 0 ? n_frames : 1;
        long long dxk = (lx * f) / nf, dyk = (ly * f) / nf;  // C division truncates toward zero on host and device alike
        long long cx = (x0 + dxk) % W, cy = ((y0 + dyk) % H + H) % H;
        long long dx = x - cx, dy = y - cy;
        if (dx * dx + dy * dy <= 1600) v = (k & 1) ? 232 - 8 * ch : 24 + 8 * ch;
    }
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

}  // namespace chb
