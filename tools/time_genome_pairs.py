#!/usr/bin/env python
"""Development aid: per-pair device-resident compress times of the 24-pair bench genome (which pairs are slower than their size says?).
usage: time_genome_pairs.py [first last]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import sccg_b200
from sccg_genome_compression_b200 import synth
a = int(sys.argv[1]) if len(sys.argv) > 1 else 0
b = int(sys.argv[2]) if len(sys.argv) > 2 else len(synth.HG19_LENGTHS)
ctx = sccg_b200.Context(0)
pad = torch.zeros(64, dtype=torch.uint8)
tc = tm = 0.0
for i in range(a, b):
    n = synth.HG19_LENGTHS[i]
    ref, tgt = synth.local_pair(n, synth.seed_for(2, i))
    d_ref = torch.cat([torch.from_numpy(ref), pad]).cuda(); d_tgt = torch.cat([torch.from_numpy(tgt), pad]).cuda()
    cm, mm = [], []
    for it in range(6):
        ctx.compress_device(d_ref.data_ptr(), n, d_tgt.data_ptr(), n, b">x")
        p = ctx.profile()
        if it >= 2: cm.append(p["kernels_ms"]); mm.append(p["match_ms"])
    tc += min(cm); tm += min(mm)
    print(f"pair {i:2d} n={n:>10d}: compress {min(cm)*1e3:7.1f} us  matcher {min(mm)*1e3:7.1f} us  {n/1000/(min(mm)*1e3):6.0f} seg/us  tail {(min(cm)-min(mm))*1e3:5.1f}", flush=True)
    del d_ref, d_tgt
print(f"sum compress {tc:.3f} ms, matcher {tm:.3f} ms")
