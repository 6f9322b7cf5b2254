#!/usr/bin/env python
"""Development aid: device-resident compress times of local-mode pairs of several sizes (chr1-, chr8-, chr21-sized) with the
library as built; environment overrides (SCCG_LM_TWO_PHASE_MIN, SCCG_LM_QUEUE_CTAS, ...) apply.  usage: time_local.py [lib.so]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import sccg_b200
from sccg_genome_compression_b200 import synth
ctx = sccg_b200.Context(0, lib_path=sys.argv[1]) if len(sys.argv) > 1 else sccg_b200.Context(0)
pad = torch.zeros(64, dtype=torch.uint8)
tot_c = 0.0
for n in (synth.CHR1_LEN, 146_364_022, 48_129_895):
    ref, tgt = synth.local_pair(n, synth.seed_for(2, n % 97))
    d_ref = torch.cat([torch.from_numpy(ref), pad]).cuda(); d_tgt = torch.cat([torch.from_numpy(tgt), pad]).cuda()
    cm, mm = [], []
    for it in range(8):
        ptr, ln, mode = ctx.compress_device(d_ref.data_ptr(), n, d_tgt.data_ptr(), n, b">x")
        p = ctx.profile()
        if it >= 3: cm.append(p["kernels_ms"]); mm.append(p["match_ms"])
    print(f"n={n}: compress {min(cm):.4f} ms (matcher {min(mm):.4f}), tail {min(cm) - min(mm):.4f}", flush=True)
    tot_c += min(cm)
print(f"sum: compress {tot_c:.4f} ms")
