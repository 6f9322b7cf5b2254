#!/usr/bin/env python
"""Headline metrics, stall reasons, and samples per opcode / top instructions of one .ncu-rep.  usage: ncu_ops.py <rep> [top]"""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h, v = raw[0], raw[-1]
want = ['gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'l1tex__throughput.avg.pct', 'lts__throughput.avg.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for i, n in enumerate(h):
    if any((n + ' ').startswith(w) for w in want): print(f"{n:75s} {v[i]}")
st = [(float(v[i]), n) for i, n in enumerate(h) if 'average_warps_issue_stalled' in n and n.endswith('per_issue_active.ratio')]
print("stalls:", ", ".join(f"{n.split('stalled_')[1].split('_per_')[0]}={x:.2f}" for x, n in sorted(st, reverse=True)[:7]))
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr = src[1]; data = [r for r in src[2:] if len(r) > 10]
isrc, ismp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
tot = sum(int(r[ismp]) for r in data); g = defaultdict(int); ge = defaultdict(int)
for r in data:
    t = r[isrc].split(); op = t[1] if t[0].startswith('@') else t[0]
    g[op] += int(r[ismp]); ge[op] += int(r[iex])
print("SASS instrs", len(data), "samples", tot)
for k, x in sorted(g.items(), key=lambda x: -x[1])[:12]: print(f"  {k:22s} {100 * x / tot:5.1f}%  executed {ge[k]}")
for i, r in sorted(sorted(enumerate(data), key=lambda x: -int(x[1][ismp]))[:top]): print(f"  {i:5d} {r[isrc][:64]:64s} {100 * int(r[ismp]) / tot:5.1f}% {r[iex]}")
