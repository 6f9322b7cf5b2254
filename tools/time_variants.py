#!/usr/bin/env python
"""Development aid: compress kernel times of the bench workload for builds with different -D flags.
usage: time_variants.py "<flags A>" "<flags B>" ...   (an argument that ends in .so is a library built beforehand, e.g. under build/)"""
import subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import sccg_b200
from sccg_genome_compression_b200 import synth
ref, tgt = synth.local_pair(synth.CHR1_LEN, synth.seed_for(2, 0))
pad = torch.zeros(64, dtype=torch.uint8)
d_ref = torch.cat([torch.from_numpy(ref), pad]).cuda(); d_tgt = torch.cat([torch.from_numpy(tgt), pad]).cuda()
for i, flags in enumerate(sys.argv[1:] or [""]):
    so = f"/tmp/libsccg_cvar{i}.so"
    if flags.endswith(".so"): so = flags
    else: subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "--cudart", "static",
                           "-ccbin", "/usr/bin/g++", *flags.split(), "-o", so, str(ROOT / "sccg-genome-compression_b200/csrc/sccg_b200.cu")])
    ctx = sccg_b200.Context(0, lib_path=so)
    ms = []; tot = []
    for it in range(10):
        ctx.compress_device(d_ref.data_ptr(), ref.size, d_tgt.data_ptr(), tgt.size, b">x")
        p = ctx.profile()
        if it >= 4: ms.append(p["match_ms"]); tot.append(p["kernels_ms"])
    print(f"[{flags}] match {min(ms):.4f} ms  compress kernels {min(tot):.4f} ms", flush=True)
    ctx.close()
