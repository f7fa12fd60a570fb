"""Static SASS opcode histogram per kernel of libchrono_b200.so (cuobjdump -sass; needs no GPU):
python tools/sass_histogram.py > profiles/r2_sass_opcodes.txt
Lists, per kernel, the instruction count, the top opcodes and the Blackwell-specific ones the design leans on
(UBLKCP / SYNCS = TMA bulk copy + mbarrier, LDGSTS = cp.async, ACQBULK = programmatic dependent launch wait, VABSDIFF4, IDP)."""
import os, re, subprocess, sys
from collections import Counter, OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "chrono_photo_b200", "libchrono_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kernels = OrderedDict()
cur = None
arch = set()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); kernels[cur] = Counter(); continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m: arch.add(m.group(1))
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur: kernels[cur][m.group(2)] += 1
KEY = ["UBLKCP", "SYNCS", "LDGSTS", "ACQBULK", "VABSDIFF4", "IDP", "PRMT", "VIMNMX", "VIMNMX3", "ATOMS", "SHFL", "LDL", "STL"]
print("libchrono_b200.so: cubin architectures", sorted(arch), "-", len(kernels), "kernels")
print("(static instruction counts; UBLKCP/SYNCS = TMA bulk copy + mbarrier, LDGSTS = cp.async, ACQBULK = griddepcontrol.wait)\n")
total = Counter()
for name, c in kernels.items():
    n = sum(c.values()); total.update(c)
    d = demangle(name)
    d = re.sub(r"\(chb::(OutlierArgs|VideoArgs|SimpleArgs|ShakeArgs)\)", "", d).replace("chb::", "").replace("(int)", "")
    top = ", ".join(f"{o} {v}" for o, v in c.most_common(6))
    key = ", ".join(f"{k} {c[k]}" for k in KEY if c[k])
    print(f"{d[:64]:64s} {n:6d} instr | {top} | {key}")
print("\nall kernels:", sum(total.values()), "instructions;", ", ".join(f"{k} {total[k]}" for k in KEY if total[k]))
print("tensor-core / wgmma opcodes (HMMA, IMMA, UTCMMA, ...):", sum(v for k, v in total.items() if "MMA" in k), "(none: there is no contraction on this path)")
