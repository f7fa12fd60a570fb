"""Second, independent restatement of chrono-photo's compositing arithmetic (numpy float32 scalars, pure-Python loops;
small stacks only). TEST INFRASTRUCTURE: cross-checks oracle/chrono_oracle.c, which is otherwise pinned only by the
reference's single quartile vector. Written from the Rust source (src/chrono.rs:208-591, src/simple.rs:43-136,
src/color.rs:4-44, src/options.rs:113-139,223-231), not from the C oracle.
"""
import math

import numpy as np

f32 = np.float32
MASK64 = (1 << 64) - 1


def rust_round(x):
    """f32::round: half away from zero."""
    x = float(x)
    if math.isnan(x) or math.isinf(x):
        return x
    return float(math.floor(abs(x) + 0.5) * (1.0 if x >= 0 else -1.0))


def as_u8(x):
    """`f32 as u8`: saturating, NaN -> 0."""
    x = float(x)
    if math.isnan(x):
        return 0
    return int(max(0.0, min(255.0, math.trunc(x)))) if not math.isinf(x) else (255 if x > 0 else 0)


def rng_range(seed, pixel, draw, n):
    z = (seed + 0x9E3779B97F4A7C15 * (pixel + 1) + 0xD1B54A32D192ED03 * draw) & MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    z ^= z >> 31
    return ((z >> 32) * n) >> 32


class Fade:
    def __init__(self, is_none=True, mode=0, absolute=True, offset=0, values=()):
        self.is_none, self.mode, self.absolute, self.offset = is_none, mode, absolute, offset
        self.values = [f32(v) for v in values]

    @staticmethod
    def build(mode, absolute, frames):
        offset = frames[0][0]
        length = frames[-1][0] - offset
        values = []
        idx = 0
        for i in range(length + 1):
            f1, v1 = frames[idx]
            f2, v2 = frames[idx + 1]
            frame = i + offset
            values.append(f32(v1) + (f32(v2) - f32(v1)) * f32(frame - f1) / f32(f2 - f1))
            if frame == f2 and idx + 2 < len(frames):
                idx += 1
        return Fade(False, mode, absolute, offset, values)

    def get(self, frame):
        if self.is_none:
            return f32(1.0)
        i = frame - self.offset
        n = len(self.values)
        if 0 <= i < n:
            return self.values[i]
        if self.mode == 0:
            return self.values[0] if i < 0 else self.values[n - 1]
        while i < 0:
            i += n
        return self.values[i % n]

    def at(self, frame, total, offset):
        return self.get(offset + frame) if self.absolute else self.get(total - frame - 1)


def blend_value(thr, dist):
    mn, mx, scale = thr
    if dist <= mn:
        return f32(0.0)
    if dist >= mx:
        return f32(1.0)
    return (dist - mn) * scale


def blend_into_u8(a, b, blend):
    if blend <= 0:
        return
    for i in range(len(a)):
        if blend >= 1:
            a[i] = b[i]
        else:
            aa = f32(a[i])
            a[i] = as_u8(rust_round(aa + (f32(b[i]) - aa) * blend))


def median_sorted(d):
    n = len(d)
    if (n + 1) % 2 == 0:
        return f32(d[(n + 1) // 2 - 1])
    i = (n + 1) // 2
    return f32(0.5) * (f32(d[i - 1]) + f32(d[i]))


def quantile_sorted(d, q):
    pos = f32(len(d) + 1) * f32(q)
    p1 = int(pos) - 1
    frac = pos - f32(math.trunc(float(pos)))
    if frac < f32(0.001):
        return f32(d[p1])
    if frac > f32(0.999):
        return f32(d[p1 + 1])
    return (f32(1.0) - frac) * f32(d[p1]) + frac * f32(d[p1 + 1])


def outlier(stack, absolute, thr_min, thr_max, thr_scale, bg, om, weights=(1, 1, 1, 1), fade=None, indices=None,
            sample_pos=None, seed=0, pixel_offset=0):
    """stack (N,H,W,C) uint8 -> (image, mask, warnings). bg: 0 first 1 random 2 average 3 median; om: 0..5."""
    N, H, W, C = stack.shape
    fade = fade or Fade()
    frames = list(range(N)) if indices is None else list(indices)
    n = len(frames)
    spos = list(range(n)) if sample_pos is None else list(sample_pos)
    frame_offset = 0 if indices is None else frames[0]
    thr = (f32(thr_min), f32(thr_max), f32(thr_scale))
    thr_sq = thr[0] * thr[0]
    w = [f32(x) for x in weights]
    img = np.zeros((H, W, C), np.uint8)
    msk = np.zeros((H, W, C), np.uint8)
    warnings = 0
    for y in range(H):
        for x in range(W):
            col = [[int(stack[f, y, x, c]) for c in range(C)] for f in frames]
            pixel_id = pixel_offset + y * W + x
            med = [f32(0)] * 4
            iqr_inv = [f32(0)] * 4
            for c in range(C):
                if w[c] != 0:
                    d = sorted(col[s][c] for s in spos)
                    med[c] = median_sorted(d)
                    if not absolute:
                        r = quantile_sorted(d, 0.75) - quantile_sorted(d, 0.25)
                        if r == 0:
                            r = f32(1.0)
                        iqr_inv[c] = f32(1.0) / r
            outl = []
            max_d, max_i = f32(0), 0
            for s in range(n):
                dsq = f32(0)
                for c in range(C):
                    if w[c] != 0:
                        diff = med[c] - f32(col[s][c])
                        if diff != 0:
                            t = w[c] * diff if absolute else (w[c] * iqr_inv[c]) * diff
                            sg = f32(-1.0) if math.copysign(1.0, float(w[c])) < 0 else f32(1.0)
                            dsq = dsq + sg * (t * t)
                        else:
                            dsq = dsq + f32(0)
                if dsq >= thr_sq:
                    outl.append((s, dsq))
                    if dsq > max_d:
                        max_d, max_i = dsq, s
            k = len(outl)
            warn = False
            if bg == 2:
                mean = [f32(0)] * C
                for s in range(n):
                    for c in range(C):
                        mean[c] = mean[c] + f32(col[s][c])
                mean = [m / f32(n) for m in mean]
                if k == 0:
                    pixel = [as_u8(rust_round(m)) for m in mean]
                elif k == 1:
                    smp = col[outl[0][0]]
                    pixel = [as_u8(rust_round(mean[c] * (f32(n) / f32(n - 1)) - f32(smp[c]) / f32(n))) for c in range(C)]
                else:
                    osum = [f32(0)] * C
                    for s, _ in outl:
                        for c in range(C):
                            osum[c] = osum[c] + f32(col[s][c])
                    with np.errstate(divide="ignore", invalid="ignore"):
                        pixel = [as_u8(rust_round(mean[c] * (f32(n) / f32(n - k)) - osum[c] / f32(n))) for c in range(C)]
            elif bg == 3:
                pixel = [as_u8(rust_round(med[c])) for c in range(C)]
            else:
                if bg == 0:
                    if k == 0:
                        idx = 0
                    elif k == n:
                        idx, warn = 0, True
                    else:
                        oset = {s for s, _ in outl}
                        idx = next(i for i in range(n) if i not in oset)
                else:
                    if k == 0:
                        idx = rng_range(seed, pixel_id, 0, n)
                    elif k == n:
                        idx, warn = rng_range(seed, pixel_id, 0, n), True
                    else:
                        perm = list(range(n))
                        cand = n
                        for s, _ in outl:
                            perm[s], perm[cand - 1] = perm[cand - 1], perm[s]
                            cand -= 1
                        idx = perm[rng_range(seed, pixel_id, 0, cand)]
                pixel = list(col[idx])
            blend_byte = 0
            if k == 1:
                s, dsq = outl[0]
                blend = fade.at(s, n, frame_offset) * blend_value(thr, np.sqrt(dsq))
                blend_into_u8(pixel, col[s], blend)
                blend_byte = as_u8(rust_round(blend * f32(255.0)))
            elif k > 1:
                if om in (4, 5):
                    acc = [f32(p) for p in pixel]
                    binv = f32(1.0)
                    for s, dsq in (outl if om == 4 else reversed(outl)):
                        blend = fade.at(s, n, frame_offset) * blend_value(thr, np.sqrt(dsq))
                        if blend > 0:
                            for c in range(C):
                                acc[c] = f32(col[s][c]) if blend >= 1 else acc[c] + (f32(col[s][c]) - acc[c]) * blend
                        binv = binv * (f32(1.0) - blend)
                    pixel = [as_u8(rust_round(a)) for a in acc]
                    blend_byte = as_u8(rust_round((f32(1.0) - binv) * f32(255.0)))
                else:
                    if om == 3:
                        msum = [f32(0)] * C
                        mdist = f32(0)
                        for s, dsq in outl:
                            for c in range(C):
                                msum[c] = msum[c] + f32(col[s][c])
                            mdist = mdist + np.sqrt(dsq)
                        smp = [as_u8(rust_round(m / f32(k))) for m in msum]
                        sidx, dist = 0, mdist / f32(k)
                    else:
                        sidx, dsq = outl[0] if om == 0 else (outl[-1] if om == 1 else (max_i, max_d))
                        smp, dist = col[sidx], np.sqrt(dsq)
                    blend = fade.at(sidx, n, frame_offset) * blend_value(thr, dist)
                    blend_into_u8(pixel, smp, blend)
                    blend_byte = as_u8(rust_round(blend * f32(255.0)))
            if warn:
                warnings += 1
            for c in range(C):
                img[y, x, c] = pixel[c]
                msk[y, x, c] = blend_byte if c < 3 else 255
    return img, msk, warnings


def simple(stack, darker, weights=(1, 1, 1, 1), fade=None, indices=None):
    N, H, W, C = stack.shape
    fade = fade or Fade()
    frames = list(range(N)) if indices is None else list(indices)
    n = len(frames)
    frame_offset = 0 if indices is None else frames[0]
    w = [f32(x) for x in weights]
    out = np.zeros((H, W, C), np.uint8)
    ext = np.full((H, W), np.finfo(np.float32).max if darker else np.finfo(np.float32).min, np.float32)
    for s, f in enumerate(frames):
        fd = fade.at(s, n, frame_offset)
        for y in range(H):
            for x in range(W):
                v = f32(0)
                for c in range(C):
                    v = v + f32(stack[f, y, x, c]) * w[c]
                if (v < ext[y, x]) if darker else (v > ext[y, x]):
                    ext[y, x] = v
                    if fd > 0:
                        px = [int(p) for p in out[y, x]]
                        blend_into_u8(px, [int(p) for p in stack[f, y, x]], fd)
                        out[y, x] = px
    return out
