#!/usr/bin/env python
"""Development aid: device-resident compress times of the two global-mode configs for builds with different -D flags.
usage: time_global_variants.py "<flags A>" "<flags B>" ..."""
import subprocess, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import sccg_b200
from sccg_genome_compression_b200 import synth
pairs = {"gap": synth.global_gap_pair(63_811_651, 59_128_983, synth.seed_for(1)), "divergent": synth.divergent_pair(48_129_895, synth.seed_for(3))}
pad = torch.zeros(64, dtype=torch.uint8)
dev = {k: (torch.cat([torch.from_numpy(r), pad]).cuda(), torch.cat([torch.from_numpy(t), pad]).cuda(), r.size, t.size) for k, (r, t) in pairs.items()}
for i, flags in enumerate(sys.argv[1:] or [""]):
    so = f"/tmp/libsccg_gvar{i}.so"
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "--cudart", "static",
                           "-ccbin", "/usr/bin/g++", *flags.split(), "-o", so, str(ROOT / "sccg-genome-compression_b200/csrc/sccg_b200.cu")])
    ctx = sccg_b200.Context(0, lib_path=so)
    out = []
    for name, (dr, dt, nr, nt) in dev.items():
        ms = []
        for it in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            _, n, mode = ctx.compress_device(dr.data_ptr(), nr, dt.data_ptr(), nt, b">x")
            ms.append((time.perf_counter() - t0) * 1e3)
        out.append(f"{name} {min(ms[1:]):.3f} ms ({n} B, mode {mode})")
    print(f"[{flags}] " + "; ".join(out), flush=True)
    ctx.close()
