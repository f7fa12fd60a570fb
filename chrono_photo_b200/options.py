"""Option value types of the compositing path, mirroring chrono-photo's `options.rs` / `flist.rs` surface:
same names, same string grammars, same f32 arithmetic (done in the C library, not in Python floats).

    SelectionMode, FadeMode, Fade, Threshold, OutlierSelectionMode, BackgroundMode   src/options.rs
    FrameRange                                                                        src/flist.rs:9-79
"""
import ctypes as C
import enum

import numpy as np

from . import _lib


class ParseEnumError(ValueError):
    """src/lib.rs ParseEnumError"""


class ParseOptionError(ValueError):
    """src/lib.rs ParseOptionError"""


class SelectionMode(enum.Enum):
    """src/options.rs:8-31"""
    OUTLIER = "outlier"
    LIGHTER = "lighter"
    DARKER = "darker"

    @classmethod
    def from_str(cls, s):
        try:
            return cls(s)
        except ValueError:
            raise ParseEnumError(f"Not a pixel selection mode: {s}. Must be one of (lighter|darker|outlier)")


class FadeMode(enum.IntEnum):
    """src/options.rs:35-56"""
    CLAMP = 0
    REPEAT = 1

    @classmethod
    def from_str(cls, s):
        if s == "repeat":
            return cls.REPEAT
        if s == "clamp":
            return cls.CLAMP
        raise ParseEnumError(f"Not a fade mode: {s}. Must be one of (repeat|clamp)")


class OutlierSelectionMode(enum.IntEnum):
    """src/options.rs:284-315"""
    FIRST = 0
    LAST = 1
    EXTREME = 2
    AVERAGE = 3
    ALL_FORWARD = 4
    ALL_BACKWARD = 5

    @classmethod
    def from_str(cls, s):
        table = {"first": cls.FIRST, "last": cls.LAST, "extreme": cls.EXTREME, "average": cls.AVERAGE,
                 "forward": cls.ALL_FORWARD, "backward": cls.ALL_BACKWARD}
        if s not in table:
            raise ParseEnumError(f"Not an outlier selection mode: {s}. Must be one of (first|last|extreme|average|forward|backward)")
        return table[s]


class BackgroundMode(enum.IntEnum):
    """src/options.rs:319-344"""
    FIRST = 0
    RANDOM = 1
    AVERAGE = 2
    MEDIAN = 3

    @classmethod
    def from_str(cls, s):
        table = {"first": cls.FIRST, "random": cls.RANDOM, "average": cls.AVERAGE, "median": cls.MEDIAN}
        if s not in table:
            raise ParseEnumError(f"Not a background pixel selection mode: {s}. Must be one of (first|random|average|median)")
        return table[s]


class Threshold:
    """src/options.rs:188-280. `min`/`max`/`scale` are the reference's internal f32 values (abs thresholds * 255)."""

    def __init__(self, absolute, min, max):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        _lib.lib().chb_threshold_new(1 if absolute else 0, C.c_float(min), C.c_float(max), C.byref(a), C.byref(b), C.byref(c))
        self.absolute = bool(absolute)
        self.min, self.max, self.scale = a.value, b.value, c.value

    @classmethod
    def abs(cls, min, max):
        return cls(True, min, max)

    @classmethod
    def rel(cls, min, max):
        return cls(False, min, max)

    @classmethod
    def from_str(cls, s):
        parts = s.split("/")
        if parts[0] in ("absolute", "abs"):
            absolute = True
        elif parts[0] in ("relative", "rel"):
            absolute = False
        else:
            raise ParseOptionError(f"Not a pixel outlier detection mode: {s}. Must be one of (abs[olute]|rel[ative])/<min>[/<max>]")
        if len(parts) < 2:
            raise ParseOptionError(f"Unexpected format in {s}")
        try:
            mn = float(parts[1])
            mx = float(parts[2]) if len(parts) > 2 else mn  # single value: max = min (:269-276)
        except ValueError:
            raise ParseOptionError(f"Unable to parse threshold for outlier detection: {s}")
        return cls(absolute, mn, mx)


class Fade:
    """src/options.rs:59-185. The LUT is built by the library's chb_fade_build (f32, same evaluation order)."""

    def __init__(self, mode, absolute, frames):
        frames = list(frames)
        if len(frames) < 2:
            raise ParseOptionError("Fade requires at least two frames specified.")
        fr = np.asarray([f for f, _ in frames], dtype=np.int32)
        va = np.asarray([v for _, v in frames], dtype=np.float32)
        cap = int(fr[-1] - fr[0]) + 1
        if cap < 1:
            raise ParseOptionError("Fade frames must be ordered by frame")
        out = np.zeros(cap, dtype=np.float32)
        off = C.c_int32()
        n = _lib.lib().chb_fade_build(fr.ctypes.data_as(C.POINTER(C.c_int32)), va.ctypes.data_as(C.POINTER(C.c_float)), len(frames),
                                      out.ctypes.data_as(C.POINTER(C.c_float)), cap, C.byref(off))
        if n < 0:
            raise ParseOptionError("Invalid fade specification")
        self.is_none = False
        self.mode = FadeMode(mode)
        self.absolute = bool(absolute)
        self.offset = off.value
        self.values = out[:n]

    @classmethod
    def none(cls):
        """Fade::none(), src/options.rs:97-105"""
        f = cls.__new__(cls)
        f.is_none, f.mode, f.absolute, f.offset = True, FadeMode.CLAMP, True, 0
        f.values = np.zeros(0, dtype=np.float32)
        return f

    @classmethod
    def from_str(cls, s):
        parts = s.split("/")
        if len(parts) < 2:
            raise ParseOptionError(f"Unexpected format in {s}")
        mode = FadeMode.from_str(parts[0])
        if parts[1] in ("absolute", "abs"):
            absolute = True
        elif parts[1] in ("relative", "rel"):
            absolute = False
        else:
            raise ParseOptionError(f"Not a frame fade spec: {s}")
        frames = []
        for p in parts[2:]:
            fv = p.split(",")
            if len(fv) != 2:
                raise ParseOptionError(f"Expected (int,float) per frame for fade. Got: {s}")
            try:
                frames.append((int(fv[0]), float(fv[1])))
            except ValueError:
                raise ParseOptionError(f"Expected (int,float) per frame for fade. Got: {s}")
        return cls(mode, absolute, frames)

    def get(self, frame):
        """Fade::get, src/options.rs:113-139 (host mirror, used by the video helpers and tests)."""
        if self.is_none:
            return 1.0
        i = frame - self.offset
        n = len(self.values)
        if 0 <= i < n:
            return float(self.values[i])
        if self.mode == FadeMode.CLAMP:
            return float(self.values[0] if i < 0 else self.values[n - 1])
        while i < 0:
            i += n
        return float(self.values[i % n])

    def to_c(self):
        f = _lib.Fade()
        f.is_none = 1 if self.is_none else 0
        f.mode = int(self.mode)
        f.absolute = 1 if self.absolute else 0
        f.offset = int(self.offset)
        f.n_values = len(self.values)
        self._keep = np.ascontiguousarray(self.values, dtype=np.float32)
        f.values = self._keep.ctypes.data_as(C.POINTER(C.c_float)) if len(self._keep) else None
        return f


class FrameRange:
    """src/flist.rs:9-79: start/end/step with '.' for an open end."""

    def __init__(self, start=None, end=None, step=1):
        self.start, self.end, self.step = start, end, int(step)

    @classmethod
    def empty(cls):
        return cls(None, None, 1)

    def range(self):
        if self.start is None or self.end is None:
            return None
        return self.end - self.start

    @classmethod
    def from_str(cls, s):
        parts = s.split("/")
        if len(parts) != 3:
            raise ParseOptionError(f"Option --frames expects 3 elements: start/end/step, {len(parts)} were suppied")
        vals = []
        for i, p in enumerate(parts):
            if p == ".":
                vals.append(None)
            else:
                try:
                    vals.append(int(p))
                except ValueError:
                    raise ParseOptionError(f"Can't parse element {i} in option --frames (start/end/step), got '{p}'.")
        return cls(vals[0], vals[1], vals[2] if vals[2] is not None else 1)


class ShakeParams:
    """src/shake.rs:44-86: `--shake <anchor-radius>/<search-radius>`."""

    def __init__(self, anchor_radius, search_radius):
        self.anchor_radius, self.search_radius = int(anchor_radius), int(search_radius)

    @classmethod
    def from_str(cls, s):
        parts = s.split("/")
        if len(parts) != 2:
            raise ParseOptionError(f"Unexpected format in shake parameters, expected <rad>/<search-rad>: {s}")
        try:
            rad, search = int(parts[0]), int(parts[1])
            if rad < 0 or search < 0:
                raise ValueError
        except ValueError:  # the reference panics here (`expect`), u32 parse
            raise ParseOptionError(f"Unexpected format in shake parameter: {s}")
        return cls(rad, search)


class ShakeAnchor:
    """src/shake.rs:88-119: `--shake-anchors x/y`."""

    def __init__(self, x, y):
        self.anchor = (int(x), int(y))

    @classmethod
    def from_str(cls, s):
        parts = s.split("/")
        if len(parts) != 2:
            raise ParseOptionError(f"Unexpected format in shake anchor, expected x/y: {s}")
        try:
            return cls(int(parts[0]), int(parts[1]))
        except ValueError:
            raise ParseOptionError(f"Unexpected format in shake anchor, expected x/y: {s}")
