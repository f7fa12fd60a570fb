"""Builds libchrono_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "chb_api.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "chb_kernels.cuh"), os.path.join(HERE, "csrc", "chb_common.cuh"), os.path.join(HERE, "csrc", "chb_jpeg.inc"),
        os.path.join(HERE, "..", "include", "chrono_b200.h")]
OUT = os.path.join(HERE, "libchrono_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # Rust never contracts a*b+c; blended bytes must round like the reference
    "-shared", "-Xcompiler", "-fPIC", "--threads", "0",
    "-lnvjpeg",  # JPEG ingest / output encode on the GPU (chb_jpeg.inc)
]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    print("[chrono_photo_b200] " + " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
