#!/usr/bin/env python
"""Development aid: times compress_device of the bench workload for several builds of the library."""
import sys, glob
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import sccg_b200
from sccg_genome_compression_b200 import synth
ref, tgt = synth.local_pair(synth.CHR1_LEN, synth.seed_for(2, 0))
pad = torch.zeros(64, dtype=torch.uint8)
d_ref = torch.cat([torch.from_numpy(ref), pad]).cuda(); d_tgt = torch.cat([torch.from_numpy(tgt), pad]).cuda()
for so in sorted(sys.argv[1:]):
    ctx = sccg_b200.Context(0, lib_path=so)
    ms = []; tot = []
    for it in range(8):
        ctx.compress_device(d_ref.data_ptr(), ref.size, d_tgt.data_ptr(), tgt.size, b">x")
        p = ctx.profile()
        if it >= 3: ms.append(p["match_ms"]); tot.append(p["kernels_ms"])
    print(f"{so}: match {sum(ms)/len(ms):.3f} ms  total {sum(tot)/len(tot):.3f} ms", flush=True)
    ctx.close()
