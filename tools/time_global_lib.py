#!/usr/bin/env python
"""Development aid: device-resident time of the divergent chr21-shaped pair for prebuilt library variants.
usage: time_global_lib.py <lib.so> [<lib.so> ...]   (environment overrides such as SCCG_GP_SCAN_GRID apply to all)"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import sccg_b200
from sccg_genome_compression_b200 import synth
ref, tgt = synth.divergent_pair(48_129_895, synth.seed_for(3))
pad = torch.zeros(64, dtype=torch.uint8)
d_ref = torch.cat([torch.from_numpy(ref), pad]).cuda(); d_tgt = torch.cat([torch.from_numpy(tgt), pad]).cuda()
for so in sys.argv[1:]:
    ctx = sccg_b200.Context(0, lib_path=so)
    best = None
    for it in range(5):
        ctx.compress_device(d_ref.data_ptr(), ref.size, d_tgt.data_ptr(), tgt.size, b">x")
        p = ctx.profile()
        if it >= 2 and (best is None or p["kernels_ms"] < best["kernels_ms"]): best = p
    print(so, {k: round(v, 3) if isinstance(v, float) else v for k, v in best.items() if k in ("kernels_ms", "index_ms", "parse_ms", "spec_rounds", "front_steps")}, flush=True)
    ctx.close()
