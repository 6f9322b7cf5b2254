#!/usr/bin/env python
"""Times the global-mode configurations of BASELINE.json (configs[0] chr19-shaped gap pair, configs[2]
chr21-shaped divergent pair) through the device-resident entry point.  Development aid."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
import sccg_b200
from sccg_genome_compression_b200 import synth

args = [a for a in sys.argv[1:] if not a.startswith("--")]
scale = float(args[0]) if args else 1.0
verify = "--verify" in sys.argv
ctx = sccg_b200.Context(0)
for name, (ref, tgt) in {
    "cfg1_chr19_gap": synth.global_gap_pair(int(63_811_651 * scale), int(59_128_983 * scale), synth.seed_for(1)),
    "cfg3_chr21_divergent": synth.divergent_pair(int(48_129_895 * scale), synth.seed_for(3)),
}.items():
    d_ref = torch.from_numpy(ref).cuda(); d_tgt = torch.from_numpy(tgt).cuda()
    res = []
    for it in range(3):
        t0 = time.perf_counter()
        ptr, n, mode = ctx.compress_device(d_ref.data_ptr(), ref.size, d_tgt.data_ptr(), tgt.size, b">" + name.encode())
        wall = time.perf_counter() - t0
        p = ctx.profile()
        res.append({"wall_ms": wall * 1e3, **p})
    out = {"config": name, "ref_bp": int(ref.size), "tgt_bp": int(tgt.size), "mode": mode, "encoded_bytes": n, "runs": res,
           "Mbp_per_s": tgt.size / (min(r["wall_ms"] for r in res) / 1e3) / 1e6}
    if verify:
        import oracle_lib as ol
        t0 = time.perf_counter()
        rc, exp, emode = ol.orc_compress(ref.tobytes(), tgt.tobytes(), b">" + name.encode())
        out["oracle_s"] = time.perf_counter() - t0
        out["matches_oracle"] = bool(exp == ctx.download(ptr, n) and emode == mode)
    print(json.dumps(out), flush=True)
