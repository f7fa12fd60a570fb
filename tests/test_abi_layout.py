"""One layout for three views of the C ABI's structs: include/chrono_b200.h as gcc lays it out, the ctypes mirror
(chrono_photo_b200/_lib.py) and the Rust binding (bindings/rust/src/ffi.rs, whose `assert!(size_of::<..>() == N)` and
offset assertions are parsed here because no Rust toolchain exists in this image)."""
import ctypes as C
import os
import re
import subprocess

from chrono_photo_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STRUCTS = {
    "chb_fade": (_lib.Fade, "ChbFade", ["is_none", "mode", "absolute", "offset", "n_values", "values"]),
    "chb_outlier_params": (_lib.OutlierParams, "ChbOutlierParams",
                           ["thr_absolute", "background", "outlier", "thr_min", "thr_max", "thr_scale", "weights", "fade", "sample_count", "seed", "pixel_offset", "block_pixels", "block_skip"]),
    "chb_simple_params": (_lib.SimpleParams, "ChbSimpleParams", ["darker", "weights", "fade"]),
    "chb_debug_planes": (_lib.DebugPlanes, "ChbDebugPlanes", ["median", "q1", "q3", "n_outliers"]),
}


def c_layout(tmp_path):
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "chrono_b200.h"', 'int main(void) {']
    for cname, (_, _, fields) in STRUCTS.items():
        src.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for f in fields:
            src.append(f'  printf("{cname} {f} %zu\\n", offsetof({cname}, {f}));')
    src += ['  return 0;', '}']
    cfile, exe = tmp_path / "layout.c", tmp_path / "layout"
    cfile.write_text("\n".join(src))
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(cfile)])
    out = {}
    for line in subprocess.check_output([str(exe)], text=True).splitlines():
        s, f, v = line.split()
        out[(s, f)] = int(v)
    return out


def test_header_ctypes_and_rust_agree(tmp_path):
    lay = c_layout(tmp_path)
    rust = open(os.path.join(ROOT, "bindings", "rust", "src", "ffi.rs")).read()
    rust_sizes = {m.group(1): int(m.group(2)) for m in re.finditer(r"assert!\(size_of::<(\w+)>\(\) == (\d+)\)", rust)}
    rust_offsets = {(m.group(1), m.group(2)): int(m.group(3)) for m in re.finditer(r"assert_eq!\(offset_of!\((\w+), (\w+)\), (\d+)\)", rust)}
    assert rust_offsets, "no offset assertions found in ffi.rs"
    for cname, (ct, rname, fields) in STRUCTS.items():
        assert C.sizeof(ct) == lay[(cname, "size")], cname
        assert rust_sizes[rname] == lay[(cname, "size")], rname
        for f in fields:
            assert getattr(ct, f).offset == lay[(cname, f)], (cname, f)
            if (rname, f) in rust_offsets:
                assert rust_offsets[(rname, f)] == lay[(cname, f)], (rname, f)
    # every offset the Rust file pins exists in the header
    for (rname, f), off in rust_offsets.items():
        cname = next(c for c, (_, r, _) in STRUCTS.items() if r == rname)
        assert lay[(cname, f)] == off


def test_rust_binding_declares_only_exported_symbols():
    rust = open(os.path.join(ROOT, "bindings", "rust", "src", "ffi.rs")).read()
    declared = set(re.findall(r"pub fn (chb_\w+)\(", rust))
    assert declared and declared <= set(_lib.SYMBOLS), declared - set(_lib.SYMBOLS)
    # the entry points SURVEY 8(b) lists are all bound
    for name in ("chb_ctx_create", "chb_stack_create", "chb_stack_upload", "chb_stack_sync", "chb_outlier", "chb_simple", "chb_stack_destroy",
                 "chb_ctx_destroy", "chb_last_error"):
        assert name in declared, name
