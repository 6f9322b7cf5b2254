#!/usr/bin/env python
"""Development aid: one pair close to the 2^31 symbol limit of the reference's int positions through the host-pointer ABI and
back (round trip = the target FASTA image).  usage: big_pair.py [n]"""
import ctypes, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import sccg_b200
from sccg_genome_compression_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000_000
t0 = time.time(); ref, tgt = synth.local_pair(n, synth.seed_for(2, 92)); print("generated", n, round(time.time() - t0, 1), "s", flush=True)
ctx = sccg_b200.Context(0)
header = b">pair near the int limit"
h_enc = torch.empty(n // 8, dtype=torch.uint8); h_out = torch.empty(n + n // 50 + 4096, dtype=torch.uint8)
cbuf = lambda a: (ctypes.c_char * a.size).from_address(a.ctypes.data)
t0 = time.time(); e_len, mode = ctx.compress_into(cbuf(ref), cbuf(tgt), header, h_enc.data_ptr(), h_enc.numel()); t1 = time.time()
d_len = ctx.decompress_into(cbuf(ref), (ctypes.c_char * e_len).from_address(h_enc.data_ptr()), h_out.data_ptr(), h_out.numel()); t2 = time.time()
got = h_out[:d_len].numpy(); hl = len(header) + 1; full = n // 50 * 50
body = got[hl:hl + full // 50 * 51].reshape(-1, 51)
ok = bytes(got[:hl]) == header + b"\n" and np.array_equal(body[:, :50].reshape(-1), tgt[:full]) and bool((body[:, 50] == 10).all()) and \
     bytes(got[hl + full // 50 * 51:]) == (tgt[full:].tobytes() + b"\n" if n > full else b"")
print({"n": n, "mode": mode, "encoded": e_len, "compress_s": round(t1 - t0, 3), "decompress_s": round(t2 - t1, 3), "roundtrip_identical": bool(ok)}, flush=True)
