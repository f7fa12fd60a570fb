"""Scratch: which part of a compositing call is not deterministic? Repeats calls on a resident full-size stack and compares
(a) composite + mask, (b) the per-pixel median plane and outlier counts (lean kernel: only those two debug planes are requested).
python tools/determinism2.py kind reps inline_min pdl"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import chrono_photo_b200 as cp
from chrono_photo_b200 import _lib
kind, reps, inline_min, pdl = (int(x) for x in sys.argv[1:5])
H, W, N = 4000, 6000, 200
ctx = cp.Context([0]); fs = cp.FrameStack(ctx, W, H, 3, N); fs.fill_synthetic(kind, 42)
cp.set_tuning("inline_min", inline_min); cp.set_tuning("pdl", pdl)
proc = cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2)
P = H * W
def call():
    out = np.empty((H, W, 3), np.uint8); msk = np.empty((H, W, 3), np.uint8)
    med = np.zeros((P, 4), np.float32); nout = np.zeros(P, np.int32)
    planes = _lib.DebugPlanes(med.ctypes.data_as(C.POINTER(C.c_float)), None, None, nout.ctypes.data_as(C.POINTER(C.c_int32)))
    warn = C.c_uint64(0); p = proc._params()
    _lib.check(_lib.lib().chb_outlier_debug(fs._h, C.byref(p), None, 0, C.c_void_p(out.ctypes.data), C.c_void_p(msk.ctypes.data), C.byref(warn), C.byref(planes)))
    return out, msk, med, nout
ref = call()
stats = {"img": 0, "median": 0, "nout": 0}
for i in range(reps):
    r = call()
    d_img = (r[0] != ref[0]).any(axis=2).ravel() | (r[1] != ref[1]).any(axis=2).ravel()
    d_med = (r[2] != ref[2]).any(axis=1)
    d_n = r[3] != ref[3]
    if d_img.any() or d_med.any() or d_n.any():
        stats["img"] += int(d_img.any()); stats["median"] += int(d_med.any()); stats["nout"] += int(d_n.any())
        px = np.nonzero(d_img | d_med | d_n)[0]
        print(f"run {i}: img {int(d_img.sum())} median {int(d_med.sum())} nout {int(d_n.sum())} tiles {sorted(set((px // 32).tolist()))[:6]}",
              "example px", int(px[0]), "median", r[2][px[0]], "vs", ref[2][px[0]], "nout", int(r[3][px[0]]), "vs", int(ref[3][px[0]]))
print(f"kind {kind} inline_min {inline_min} pdl {pdl}: runs differing in img/median/nout: {stats} of {reps}")
