#!/usr/bin/env python
"""One global-mode compress of a BASELINE config shape (for ncu launch lists).  usage: one_global.py gap|divergent [scale]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import sccg_b200
from sccg_genome_compression_b200 import synth
shape = sys.argv[1]; scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
ref, tgt = (synth.global_gap_pair(int(63_811_651 * scale), int(59_128_983 * scale), synth.seed_for(1)) if shape == "gap"
            else synth.divergent_pair(int(48_129_895 * scale), synth.seed_for(3)))
ctx = sccg_b200.Context(0)
pad = torch.zeros(64, dtype=torch.uint8)
d_ref = torch.cat([torch.from_numpy(ref), pad]).cuda(); d_tgt = torch.cat([torch.from_numpy(tgt), pad]).cuda()
for _ in range(2):
    ptr, n, mode = ctx.compress_device(d_ref.data_ptr(), ref.size, d_tgt.data_ptr(), tgt.size, b">x")
print(mode, n, ctx.profile())
