#!/usr/bin/env python
"""Development aid: which path seg_match_k takes per segment on the bench workload (-DSCCG_SEG_STATS build in /tmp)
and the kernel time with / without the diagonal-hypothesis path.  usage: seg_stats.py [chrom] [size]"""
import ctypes, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import sccg_b200
from sccg_genome_compression_b200 import synth
so = "/tmp/libsccg_stats.so"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "--cudart", "static",
                       "-ccbin", "/usr/bin/g++", "-DSCCG_SEG_STATS", "-o", so, str(ROOT / "sccg-genome-compression_b200/csrc/sccg_b200.cu")])
chrom = int(sys.argv[1]) if len(sys.argv) > 1 else 0
size = int(sys.argv[2]) if len(sys.argv) > 2 else synth.CHR1_LEN
ref, tgt = synth.local_pair(size, synth.seed_for(2, chrom))
pad = torch.zeros(64, dtype=torch.uint8)
d_ref = torch.cat([torch.from_numpy(ref), pad]).cuda(); d_tgt = torch.cat([torch.from_numpy(tgt), pad]).cuda()
for nodiag in (0, 1):
    os.environ["SCCG_NO_DIAG"] = str(nodiag)
    ctx = sccg_b200.Context(0, lib_path=so)
    ctx.lib.sccg_debug_seg_stats.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
    out = (ctypes.c_ulonglong * 8)()
    ms = []
    for it in range(6):
        ctx.lib.sccg_debug_seg_stats(out, 1)
        ctx.compress_device(d_ref.data_ptr(), ref.size, d_tgt.data_ptr(), tgt.size, b">x")
        ms.append(ctx.profile()["match_ms"])
    ctx.lib.sccg_debug_seg_stats(out, 0)
    print(f"SCCG_NO_DIAG={nodiag}: identical {out[0]} diagonal {out[1]} generic {out[2]} second-pass {out[3]}  match_ms {min(ms[2:]):.4f} (stats build)", flush=True)
    ctx.close()
