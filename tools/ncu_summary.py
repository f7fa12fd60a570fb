"""Summarise an .ncu-rep (raw page + SASS page): python tools/ncu_summary.py report.ncu-rep [n_top_lines]"""
import csv, re, subprocess, sys
from collections import Counter
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__cycles_active.avg', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum']
for i, h in enumerate(hdr):
    if h in keys:
        print(f"{h:70s} {units[i]:14s} {vals[i]}")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
hdr = rows[1]; data = rows[2:]; ix = {h: i for i, h in enumerate(hdr)}
op = Counter(); tot = 0; st = Counter()
for r in data:
    try: n = int(r[ix['Instructions Executed']])
    except Exception: continue
    tot += n
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']])
    op[m.group(2).split('.')[0] if m else '?'] += n
    for h in hdr:
        if h.startswith('stall_') and 'Not Issued' not in h:
            try: st[h] += int(r[ix[h]] or 0)
            except Exception: pass
print("total warp-instructions", tot)
print("opcodes:", ", ".join(f"{o} {100*n/tot:.1f}%" for o, n in op.most_common(16)))
ts = sum(st.values()) or 1
print("stalls:", ", ".join(f"{h[6:]} {100*n/ts:.1f}%" for h, n in st.most_common(8)))
both = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(both.splitlines()))
for i, r in enumerate(rows[:6]):
    if 'Source' in r: start = i + 1; break
agg = []
for r in rows[start:]:
    if r[0] != '':
        try: agg.append([int(r[0]), r[1].strip()[:100], int(r[7] or 0), int(r[6] or 0)])
        except Exception: pass
t2 = sum(a[2] for a in agg) or 1
ts2 = sum(a[3] for a in agg) or 1
agg.sort(key=lambda a: -a[3])
print("top source lines by stall samples (line, inst%, samples%):")
for a in agg[:ntop]: print(f"{a[0]:4d} {100*a[2]/t2:5.1f}% {100*a[3]/ts2:5.1f}% | {a[1]}")
