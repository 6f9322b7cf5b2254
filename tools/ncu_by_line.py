#!/usr/bin/env python
"""Aggregates an `ncu --page source --csv` export by CUDA source line using nvdisasm line info of the same build.
usage: ncu_by_line.py <source_page.csv> <nvdisasm --print-line-info output> <kernel substring> [top]"""
import csv, re, sys
from collections import defaultdict
src_csv, disasm, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(disasm).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('//--------------------- .text.') and kname in l)
cur = None; instrs = []
for l in lines[start + 1:]:
    if l.startswith('//--------------------- .'): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: instrs.append(cur)
rows = list(csv.reader(open(src_csv)))
secs = []; c = None
for r in rows:
    if r and r[0] == 'Kernel Name': c = {'name': r[1], 'rows': []}; secs.append(c); continue
    if r and r[0] == 'Address': c['hdr'] = r; continue
    if c is not None and len(r) > 10: c['rows'].append(r)
plain = kname.split('ILi')[0]                      # template instances: mangled name for nvdisasm (seg_match_kILi2), demangled one in the ncu export
s = next(x for x in secs if kname in x['name'] or plain in x['name'])
hdr, data = s['hdr'], s['rows']
iex, ismp = hdr.index('Instructions Executed'), hdr.index('# Samples')
assert len(data) == len(instrs), (len(data), len(instrs))
ex = defaultdict(int); sm = defaultdict(int)
for i, r in enumerate(data):
    ex[instrs[i]] += int(r[iex]); sm[instrs[i]] += int(r[ismp])
tot, tots = sum(ex.values()), sum(sm.values())
print(f"kernel {s['name'][:60]}  warp-instr {tot}  samples {tots}  sass {len(data)}")
import os
cache = {}
def text(k):
    if not k: return ''
    f = k[0]
    if f not in cache:
        cache[f] = []
        for root in ('/root/repo/sccg-genome-compression_b200/csrc',):
            p = os.path.join(root, f)
            if os.path.exists(p): cache[f] = open(p).read().split('\n')
    return cache[f][k[1] - 1].strip()[:80] if 0 < k[1] <= len(cache[f]) else ''
for k, v in sorted(ex.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{str(k):30s} ex {100*v/tot:5.1f}%  smp {100*sm[k]/max(tots,1):5.1f}%  {text(k)}")
