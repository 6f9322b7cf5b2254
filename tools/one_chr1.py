#!/usr/bin/env python
"""One chr1-sized pair (BASELINE configs[1]) through the device-resident entry points, a few times: the program ncu profiles
(tools/gpu_ncu_chr1.sh).  usage: one_chr1.py [reps] [size]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import sccg_b200
import oracle_lib as ol
from sccg_genome_compression_b200 import synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else synth.CHR1_LEN
ref, tgt = synth.local_pair(n, synth.seed_for(2, 0))
pad = torch.zeros(64, dtype=torch.uint8)
d_ref = torch.cat([torch.from_numpy(ref), pad]).cuda(); d_tgt = torch.cat([torch.from_numpy(tgt), pad]).cuda()
ctx = sccg_b200.Context(0)
H = b">chr1 synthetic hg19-vs-hg18 shape"
for _ in range(reps):
    ptr, ln, mode = ctx.compress_device(d_ref.data_ptr(), n, d_tgt.data_ptr(), n, H)
    p = ctx.profile()
    print("compress kernels_ms %.4f match_ms %.4f launches %d" % (p["kernels_ms"], p["match_ms"], p["launches"]))
enc = ctx.download(ptr, ln)
header, low, nline, body = ol.split_intermediate(enc)
d_body = torch.cat([torch.frombuffer(bytearray(body), dtype=torch.uint8), pad]).cuda()
d_low = torch.cat([torch.frombuffer(bytearray(low), dtype=torch.uint8), pad]).cuda()
d_n = torch.zeros(64, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    ctx.reconstruct_device(d_ref.data_ptr(), n, d_body.data_ptr(), len(body), d_n.data_ptr(), 0, d_low.data_ptr(), len(low))
    p = ctx.profile()
    print("decompress kernels_ms %.4f tokenizer_ms %.4f gather_ms %.4f launches %d" % (p["kernels_ms"], p["serialize_ms"], p["gather_ms"], p["launches"]))
ctx.close()
