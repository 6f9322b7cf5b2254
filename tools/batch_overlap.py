#!/usr/bin/env python
"""many pairs in one process with the upload of pair i+1 overlapping the kernels / download of pair i.
Two contexts on the same GPU (own streams and buffers each), two host threads that take pairs alternately (ctypes releases
the GIL during a call).  Compares that with the plain one-pair-after-the-other loop on the same page-locked inputs.

    python tools/batch_overlap.py [--scale S] [--pairs N]
"""
import argparse, ctypes, json, sys, threading, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import sccg_b200
from sccg_genome_compression_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.25)
ap.add_argument("--pairs", type=int, default=12)
args = ap.parse_args()
lengths = [max(20_000, int(n * args.scale)) for n in synth.HG19_LENGTHS[:args.pairs]]


def cbuf(t, n):
    return (ctypes.c_char * n).from_address(t.data_ptr())


pairs = []
for idx, n in enumerate(lengths):
    ref, tgt = synth.local_pair(n, synth.seed_for(4, idx))
    h_ref = torch.from_numpy(ref).pin_memory(); h_tgt = torch.from_numpy(tgt).pin_memory()
    h_enc = torch.empty(n // 4 + (1 << 20), dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n + n // 50 + 4096, dtype=torch.uint8).pin_memory()
    pairs.append({"n": n, "ref": h_ref, "tgt": h_tgt, "enc": h_enc, "out": h_out, "hdr": b">chr%d" % (idx + 1)})
ctxs = [sccg_b200.Context(0), sccg_b200.Context(0)]


def compress(ctx, p):
    p["e_len"], p["mode"] = ctx.compress_into(cbuf(p["ref"], p["n"]), cbuf(p["tgt"], p["n"]), p["hdr"], p["enc"].data_ptr(), p["enc"].numel())


def decompress(ctx, p):
    p["d_len"] = ctx.decompress_into(cbuf(p["ref"], p["n"]), cbuf(p["enc"], p["e_len"]), p["out"].data_ptr(), p["out"].numel())


def run(fn, overlapped: bool) -> float:
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if not overlapped:
        for p in pairs:
            fn(ctxs[0], p)
    else:
        def worker(w):
            for p in pairs[w::2]:
                fn(ctxs[w], p)
        th = [threading.Thread(target=worker, args=(w,)) for w in range(2)]
        for t in th: t.start()
        for t in th: t.join()
    return time.perf_counter() - t0


for c in ctxs:                                        # warm both contexts up on the largest pair
    compress(c, pairs[0]); decompress(c, pairs[0])
res = {}
for name, fn in (("compress", compress), ("decompress", decompress)):
    seq = min(run(fn, False) for _ in range(3))
    if name == "compress":
        ref_enc = [bytes(p["enc"][:p["e_len"]].numpy()) for p in pairs]
    ovl = min(run(fn, True) for _ in range(3))
    if name == "compress":
        assert ref_enc == [bytes(p["enc"][:p["e_len"]].numpy()) for p in pairs], "overlapped batch differs"
    res[name] = {"sequential_ms": seq * 1e3, "overlapped_ms": ovl * 1e3, "gain": seq / ovl}
bp = sum(lengths)
print(json.dumps({"pairs": len(pairs), "total_bp": bp, "scale": args.scale, **res,
                  "note": "one GPU, two contexts / two host threads vs one; page-locked buffers; wall clock around the whole batch"}))
for c in ctxs:
    c.close()
