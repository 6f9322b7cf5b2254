#!/usr/bin/env python
"""BASELINE.json configs[3] / configs[4]: whole-genome 24-chromosome synthetic hg19-vs-hg18 shape (3.1 Gbp), compression and
decompression partitioned by chromosome over the GPUs of one box (LPT packing, one process per GPU, no data-path collective).
Every pair goes end to end through the host-pointer C ABI (page-locked buffers in, page-locked buffer out); the round trip is
checked against the target FASTA image.  Synthetic pairs are generated on the host between the timed calls (not timed).

    python tools/whole_genome.py [--scale S]                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/whole_genome.py
"""
import argparse, ctypes, json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import torch.distributed as dist
import sccg_b200
from sccg_genome_compression_b200 import sharding, synth

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0, help="scale every chromosome length (debug)")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lengths = [max(20_000, int(n * args.scale)) for n in synth.HG19_LENGTHS]
mine = sharding.assign_chromosomes(lengths, world)[rank]
ctx = sccg_b200.Context(local)
cap = max(lengths[i] for i in mine) if mine else 1


def cbuf(t, n):
    return (ctypes.c_char * n).from_address(t.data_ptr())


h_ref = torch.empty(cap + 64, dtype=torch.uint8).pin_memory(); h_tgt = torch.empty(cap + 64, dtype=torch.uint8).pin_memory()
h_enc = torch.empty(cap // 4 + (1 << 20), dtype=torch.uint8).pin_memory(); h_out = torch.empty(cap + cap // 50 + 4096, dtype=torch.uint8).pin_memory()
t_comp = t_dec = 0.0
bp = enc_total = 0
ok = True
modes = []
for w in range(2):                                   # pass 0 warms the context up on the largest pair of this rank (untimed)
    for idx in (mine[:1] if w == 0 else mine):
        n = lengths[idx]
        ref, tgt = synth.local_pair(n, synth.seed_for(4, idx))
        h_ref[:n] = torch.from_numpy(ref); h_tgt[:n] = torch.from_numpy(tgt)
        header = b">chr%d synthetic hg19 shape" % (idx + 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e_len, mode = ctx.compress_into(cbuf(h_ref, n), cbuf(h_tgt, n), header, h_enc.data_ptr(), h_enc.numel())
        t1 = time.perf_counter()
        d_len = ctx.decompress_into(cbuf(h_ref, n), cbuf(h_enc, e_len), h_out.data_ptr(), h_out.numel())
        t2 = time.perf_counter()
        if w == 0:
            continue
        t_comp += t1 - t0; t_dec += t2 - t1; bp += n; enc_total += e_len; modes.append(mode)
        # round trip: "<header>\n" + 50-column lines of the target
        got = h_out[:d_len].numpy()
        hl = len(header) + 1
        full = n // 50 * 50
        body = got[hl:hl + full // 50 * 51].reshape(-1, 51)
        good = bytes(got[:hl]) == header + b"\n" and np.array_equal(body[:, :50].reshape(-1), tgt[:full]) and bool((body[:, 50] == 10).all())
        tail = bytes(got[hl + full // 50 * 51:])
        good = good and tail == (tgt[full:].tobytes() + b"\n" if n > full else b"")
        ok = ok and good
vals = torch.tensor([t_comp, t_dec, float(bp), float(enc_total), 1.0 if ok else 0.0], dtype=torch.float64, device="cuda")
if world > 1:
    mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    mn = vals.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
else:
    mx = sm = mn = vals
if rank == 0:
    tc, td, total_bp, total_enc = float(mx[0]), float(mx[1]), float(sm[2]), float(sm[3])
    print(json.dumps({"config": "whole-genome 24-chromosome synthetic hg19 shape, one pair per call, LPT over the GPUs", "n_gpus": world, "scale": args.scale,
                      "total_bp": int(total_bp), "encoded_bytes": int(total_enc), "compress_s_max_over_ranks": tc, "compress_Mbp_s": total_bp / tc / 1e6,
                      "decompress_s_max_over_ranks": td, "decompress_Gbp_s": total_bp / td / 1e9, "roundtrip_identical": bool(float(mn[4]) == 1.0),
                      "timing": "wall clock around sccg_compress_into / sccg_decompress_into (pinned host buffers), summed per rank, max over ranks"}), flush=True)
ctx.close()
if world > 1:
    dist.destroy_process_group()
