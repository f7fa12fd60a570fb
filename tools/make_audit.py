"""Renders oracle/AUDIT.md from oracle/AUDIT.md.in: {k:name} -> `name :line` in chb_kernels.cuh, {api:text} -> `:line` of the first
line of chb_api.cu that contains the text. Run after editing the kernels: python tools/make_audit.py"""
import os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
kern = open(os.path.join(ROOT, "chrono_photo_b200", "csrc", "chb_kernels.cuh")).read().split("\n")
api = open(os.path.join(ROOT, "chrono_photo_b200", "csrc", "chb_api.cu")).read().split("\n")

def kline(name):
    pat = re.compile(r"(__global__|__device__).*\b" + re.escape(name) + r"\(")
    for i, l in enumerate(kern, 1):
        if pat.search(l):
            return i
    raise SystemExit(f"make_audit: no definition of {name} in chb_kernels.cuh")

def aline(text):
    for i, l in enumerate(api, 1):
        if text in l:
            return i
    raise SystemExit(f"make_audit: '{text}' not found in chb_api.cu")

src = open(os.path.join(ROOT, "oracle", "AUDIT.md.in")).read()
out = re.sub(r"\{k:([A-Za-z0-9_]+)\}", lambda m: f"`{m.group(1)} :{kline(m.group(1))}`", src)
out = re.sub(r"\{api:([^}]+)\}", lambda m: f"`:{aline(m.group(1))}`", out)
open(os.path.join(ROOT, "oracle", "AUDIT.md"), "w").write(out)
print("oracle/AUDIT.md written")
