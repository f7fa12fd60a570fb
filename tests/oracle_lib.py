"""ctypes wrapper of the CPU oracle (oracle/chrono_oracle.c). TEST INFRASTRUCTURE: only tests/, smoke() and the
cpu_baseline / --impl reference legs of bench.py import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC_DIR = os.path.join(ROOT, "oracle")
ORC_SO = os.path.join(ORC_DIR, "_build", "liborc.so")


class OrcThreshold(C.Structure):
    _fields_ = [("absolute", C.c_int32), ("min", C.c_float), ("max", C.c_float), ("scale", C.c_float)]


class OrcFade(C.Structure):
    _fields_ = [("is_none", C.c_int32), ("mode", C.c_int32), ("absolute", C.c_int32), ("offset", C.c_int32),
                ("n_values", C.c_int32), ("values", C.POINTER(C.c_float))]


class OrcOutlierParams(C.Structure):
    _fields_ = [("threshold", OrcThreshold), ("background", C.c_int32), ("outlier", C.c_int32), ("weights", C.c_float * 4),
                ("fade", OrcFade), ("seed", C.c_uint64), ("pixel_offset", C.c_uint64)]


class OrcDebug(C.Structure):
    _fields_ = [("median", C.POINTER(C.c_float)), ("q1", C.POINTER(C.c_float)), ("q3", C.POINTER(C.c_float)),
                ("n_outliers", C.POINTER(C.c_int32)), ("sel_index", C.POINTER(C.c_int32)), ("bg_index", C.POINTER(C.c_int32))]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORC_DIR])
    return ORC_SO


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ORC_DIR, "chrono_oracle.c")
        if not os.path.exists(ORC_SO) or os.path.getmtime(ORC_SO) < os.path.getmtime(src):
            build()
        l = C.CDLL(ORC_SO)
        l.orc_threshold_blend_value.restype = C.c_float
        l.orc_threshold_blend_value.argtypes = [C.POINTER(OrcThreshold), C.c_float]
        l.orc_fade_get.restype = C.c_float
        l.orc_fade_get.argtypes = [C.POINTER(OrcFade), C.c_int32]
        l.orc_median.restype = C.c_float
        l.orc_median.argtypes = [C.c_void_p, C.c_size_t]
        l.orc_quantile.restype = C.c_float
        l.orc_quantile.argtypes = [C.c_void_p, C.c_size_t, C.c_float]
        l.orc_quartiles.argtypes = [C.c_void_p, C.c_size_t] + [C.POINTER(C.c_float)] * 3
        l.orc_threshold_new.argtypes = [C.c_int, C.c_float, C.c_float, C.POINTER(OrcThreshold)]
        l.orc_blend_into_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float]
        l.orc_blend_into_f32_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float]
        l.orc_rng_range.restype = C.c_uint32
        l.orc_rng_range.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        l.orc_outlier.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(OrcOutlierParams), C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(OrcDebug), C.c_int]
        l.orc_simple.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(OrcFade),
                                 C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        l.orc_fade_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int32)]
        l.orc_crop_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        l.orc_video_windows.argtypes = [C.c_int] * 11 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        l.orc_synth_frames.argtypes = [C.c_int, C.c_uint64] + [C.c_int] * 8 + [C.c_void_p, C.c_int]
        _lib = l
    return _lib


def threshold(absolute, mn, mx):
    t = OrcThreshold()
    lib().orc_threshold_new(1 if absolute else 0, mn, mx, C.byref(t))
    return t


def fade_none():
    return OrcFade(1, 0, 1, 0, 0, None)


def fade(mode, absolute, frames):
    """frames: list of (frame, value). Returns (OrcFade, keepalive array)."""
    fr = np.asarray([f for f, _ in frames], np.int32)
    va = np.asarray([v for _, v in frames], np.float32)
    cap = int(fr[-1] - fr[0]) + 1
    out = np.zeros(cap, np.float32)
    off = C.c_int32()
    n = lib().orc_fade_build(fr.ctypes.data, va.ctypes.data, len(frames), out.ctypes.data, cap, C.byref(off))
    assert n > 0
    return OrcFade(0, int(mode), 1 if absolute else 0, off.value, n, out.ctypes.data_as(C.POINTER(C.c_float))), out


def quartiles(values):
    a = np.ascontiguousarray(values, np.uint8)
    q1, m, q3 = C.c_float(), C.c_float(), C.c_float()
    lib().orc_quartiles(a.ctypes.data, len(a), C.byref(q1), C.byref(m), C.byref(q3))
    return q1.value, m.value, q3.value


def median(values):
    a = np.ascontiguousarray(values, np.uint8)
    return lib().orc_median(a.ctypes.data, len(a))


def blend_into_u8(a, b, blend):
    a = np.array(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    lib().orc_blend_into_u8(a.ctypes.data, b.ctypes.data, len(a), blend)
    return a


def outlier(stack, thr, background, outlier_mode, weights=(1, 1, 1, 1), fade_=None, indices=None, sample_pos=None, seed=0,
            pixel_offset=0, want_debug=False, n_threads=1):
    """stack: (N, H, W, C) uint8. Returns (image, mask, warnings[, debug dict])."""
    stack = np.ascontiguousarray(stack, np.uint8)
    N, H, W, Cc = stack.shape
    prm = OrcOutlierParams()
    prm.threshold = thr
    prm.background = int(background)
    prm.outlier = int(outlier_mode)
    for i in range(4):
        prm.weights[i] = float(weights[i]) if i < len(weights) else 1.0
    keep = None
    if fade_ is None:
        prm.fade = fade_none()
    else:
        prm.fade, keep = fade_
    prm.seed = seed
    prm.pixel_offset = pixel_offset
    out = np.zeros((H, W, Cc), np.uint8)
    mask = np.zeros((H, W, Cc), np.uint8)
    idx = None if indices is None else np.ascontiguousarray(indices, np.int32)
    sp = None if sample_pos is None else np.ascontiguousarray(sample_pos, np.int32)
    warn = C.c_uint64(0)
    dbg = None
    d = {}
    if want_debug:
        P = H * W
        d = {"median": np.zeros((P, 4), np.float32), "q1": np.zeros((P, 4), np.float32), "q3": np.zeros((P, 4), np.float32),
             "n_outliers": np.zeros(P, np.int32), "sel_index": np.zeros(P, np.int32), "bg_index": np.zeros(P, np.int32)}
        fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        dbg = OrcDebug(fp(d["median"]), fp(d["q1"]), fp(d["q3"]), ip(d["n_outliers"]), ip(d["sel_index"]), ip(d["bg_index"]))
    rc = lib().orc_outlier(stack.ctypes.data, N, H, W, Cc, C.byref(prm), idx.ctypes.data if idx is not None else None,
                           len(idx) if idx is not None else 0, sp.ctypes.data if sp is not None else None,
                           len(sp) if sp is not None else 0, out.ctypes.data, mask.ctypes.data, C.byref(warn),
                           C.byref(dbg) if dbg is not None else None, n_threads)
    if rc != 0:
        raise ValueError(f"orc_outlier returned {rc}")
    del keep
    if want_debug:
        return out, mask, warn.value, d
    return out, mask, warn.value


def simple(stack, darker, weights=(1, 1, 1, 1), fade_=None, indices=None, n_threads=1):
    stack = np.ascontiguousarray(stack, np.uint8)
    N, H, W, Cc = stack.shape
    w = (C.c_float * 4)(*[float(weights[i]) if i < len(weights) else 1.0 for i in range(4)])
    keep = None
    if fade_ is None:
        f = fade_none()
    else:
        f, keep = fade_
    out = np.zeros((H, W, Cc), np.uint8)
    idx = None if indices is None else np.ascontiguousarray(indices, np.int32)
    rc = lib().orc_simple(stack.ctypes.data, N, H, W, Cc, 1 if darker else 0, w, C.byref(f),
                          idx.ctypes.data if idx is not None else None, len(idx) if idx is not None else 0, out.ctypes.data, n_threads)
    if rc != 0:
        raise ValueError(f"orc_simple returned {rc}")
    del keep
    return out


def shake_analyze(frames, anchors, anchor_radius, search_radius, want_diffs=False):
    """ShakeAnalyzer::analyze (src/shake.rs:190-305) on decoded frames (N, H, W, C) uint8 -> offsets [(0, 0), (dx, dy), ...]
    (and, optionally, the per-frame diff tables). Raises IndexError where the reference panics (coordinate out of range)."""
    frames = np.ascontiguousarray(frames, np.uint8)
    N, H, W, Cc = frames.shape
    anc = np.ascontiguousarray(anchors, np.int32).reshape(-1, 2)
    size, ss = 2 * anchor_radius + 1, 2 * search_radius + 1
    wins = np.zeros(len(anc) * size * size * Cc, np.uint8)
    l = lib()
    l.orc_shake_fill_windows.restype = C.c_int
    l.orc_shake_offset.restype = C.c_int
    if l.orc_shake_fill_windows(C.c_void_p(frames[0].ctypes.data), W, H, Cc, C.c_void_p(anc.ctypes.data), len(anc), anchor_radius,
                                C.c_void_p(wins.ctypes.data)) != 0:
        raise IndexError("Image coordinate out of range")
    out, tables = [(0, 0)], []
    for i in range(1, N):
        diffs = np.zeros(ss * ss, np.int32)
        dx, dy = C.c_int32(), C.c_int32()
        if l.orc_shake_offset(C.c_void_p(frames[i].ctypes.data), W, H, Cc, C.c_void_p(anc.ctypes.data), len(anc), anchor_radius, search_radius,
                              C.c_void_p(wins.ctypes.data), C.c_void_p(diffs.ctypes.data), C.byref(dx), C.byref(dy)) != 0:
            raise IndexError("Image coordinate out of range")
        out.append((dx.value, dy.value))
        tables.append(diffs)
    return (out, tables) if want_diffs else out


def crop_create(offsets, width, height):
    off = np.ascontiguousarray(offsets, np.int32).reshape(-1, 2)
    xy = np.zeros_like(off)
    w, h = C.c_int32(), C.c_int32()
    r = lib().orc_crop_create(off.ctypes.data, len(off), width, height, xy.ctypes.data, C.byref(w), C.byref(h))
    return None if r == 0 else (xy, w.value, h.value)


def video_windows(image_count, vin, vout, cap=4096):
    """vin/vout: (start|None, end|None, step). Returns (count, starts, ends, numbers)."""
    ws, we, num = (np.zeros(cap, np.int32) for _ in range(3))
    n = lib().orc_video_windows(image_count, vin[0] is not None, vin[0] or 0, vin[1] is not None, vin[1] or 0, vin[2],
                                vout[0] is not None, vout[0] or 0, vout[1] is not None, vout[1] or 0, vout[2],
                                ws.ctypes.data, we.ctypes.data, num.ctypes.data, cap)
    m = min(n, cap)
    return n, ws[:m], we[:m], num[:m]


def synth_frames(kind, seed, n_frames, width, full_height, channels=3, row0=0, rows=None, f0=0, n_out=None, n_threads=None):
    """Synthetic series of the bench / tests generated on the host by the oracle's twin of the device generator:
    frames [f0, f0 + n_out) of an n_frames series, rows [row0, row0 + rows) -> uint8 [n_out][rows][width][channels]."""
    rows = full_height - row0 if rows is None else rows
    n_out = n_frames - f0 if n_out is None else n_out
    n_threads = (os.cpu_count() or 1) if n_threads is None else n_threads
    out = np.empty((n_out, rows, width, channels), dtype=np.uint8)
    rc = lib().orc_synth_frames(int(kind), int(seed), int(f0), int(n_out), int(n_frames), int(width), int(full_height), int(channels),
                                int(row0), int(rows), out.ctypes.data, int(n_threads))
    if rc != 0:
        raise ValueError("orc_synth_frames: bad argument")
    return out
