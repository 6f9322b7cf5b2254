#!/usr/bin/env python
"""Development aid: matcher time of prefixes / slices of the chr1-sized pair (same content, different sizes and places)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import sccg_b200
from sccg_genome_compression_b200 import synth
ctx = sccg_b200.Context(0, lib_path=sys.argv[1]) if len(sys.argv) > 1 else sccg_b200.Context(0)
pad = torch.zeros(64, dtype=torch.uint8)
n = synth.CHR1_LEN
ref, tgt = synth.local_pair(n, synth.seed_for(2, 0))
d_ref = torch.cat([torch.from_numpy(ref), pad]).cuda(); d_tgt = torch.cat([torch.from_numpy(tgt), pad]).cuda()
def run(off, ln):
    cm, mm = [], []
    for it in range(7):
        ctx.compress_device(d_ref.data_ptr() + off, ln, d_tgt.data_ptr() + off, ln, b">x")
        p = ctx.profile()
        if it >= 2: cm.append(p["kernels_ms"]); mm.append(p["match_ms"])
    print(f"off={off:>11d} len={ln:>11d}: matcher {min(mm)*1e3:7.1f} us  {ln/1000/(min(mm)*1e3):6.0f} seg/us   compress {min(cm)*1e3:7.1f} us  tail {(min(cm)-min(mm))*1e3:5.1f}", flush=True)
for ln in (n, 146_364_000, 48_000_000, 20_000_000):
    run(0, ln)
run(0, n)
ptr, ln, mode = ctx.compress_device(d_ref.data_ptr(), n, d_tgt.data_ptr(), n, b">x")
dev = ctx.download(ptr, ln)
host, hmode = ctx.compress(ref, tgt, b">x")                  # pipelined upload: one matcher launch per chunk, everything inline
print("device-resident image == host-path image:", dev == host and mode == hmode, len(dev))
