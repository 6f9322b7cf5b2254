#!/usr/bin/env python
"""Development aid: decompress kernel times of the bench workload for builds with different -D flags.
usage: time_gather_variants.py "<flags A>" "<flags B>" ..."""
import subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import sccg_b200, oracle_lib as ol
from sccg_genome_compression_b200 import synth
ref, tgt = synth.local_pair(synth.CHR1_LEN, synth.seed_for(2, 0))
pad = torch.zeros(64, dtype=torch.uint8)
d_ref = torch.cat([torch.from_numpy(ref), pad]).cuda(); d_tgt = torch.cat([torch.from_numpy(tgt), pad]).cuda()
base = sccg_b200.Context(0)
ptr, n, mode = base.compress_device(d_ref.data_ptr(), ref.size, d_tgt.data_ptr(), tgt.size, b">x")
enc = base.download(ptr, n); base.close()
header, low, nline, body = ol.split_intermediate(enc)
d_body = torch.cat([torch.frombuffer(bytearray(body), dtype=torch.uint8), pad]).cuda()
d_low = torch.cat([torch.frombuffer(bytearray(low), dtype=torch.uint8), pad]).cuda()
d_n = torch.zeros(64, dtype=torch.uint8, device="cuda")
d_refu = torch.cat([torch.from_numpy(np.frombuffer(ref.tobytes().upper(), dtype=np.uint8).copy()), pad]).cuda()
for i, flags in enumerate(sys.argv[1:] or [""]):
    so = f"/tmp/libsccg_var{i}.so"
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "--cudart", "static",
                           "-ccbin", "/usr/bin/g++", *flags.split(), "-o", so, str(ROOT / "sccg-genome-compression_b200/csrc/sccg_b200.cu")])
    ctx = sccg_b200.Context(0, lib_path=so)
    g, k = [], []
    for it in range(8):
        ctx.reconstruct_device(d_refu.data_ptr(), ref.size, d_body.data_ptr(), len(body), d_n.data_ptr(), 0, d_low.data_ptr(), len(low))
        p = ctx.profile()
        if it >= 3: g.append(p["gather_ms"]); k.append(p["kernels_ms"])
    print(f"[{flags}] gather {min(g):.4f} ms  decompress kernels {min(k):.4f} ms", flush=True)
    ctx.close()
