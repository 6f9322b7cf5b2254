#!/usr/bin/env python
"""Decompression of one chromosome over N GPUs by output range: every rank produces the part-th piece of the
file image (sccg_decompress_part, page-locked buffers) and uploads only the reference chunks that piece copies from.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/parts_bench.py [size]"""
import ctypes, hashlib, json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist
import sccg_b200
from sccg_genome_compression_b200 import synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
size = int(sys.argv[1]) if len(sys.argv) > 1 else synth.CHR1_LEN
ref_np, tgt_np = synth.local_pair(size, synth.seed_for(2, 0))
header = b">chr1 synthetic hg19-vs-hg18 shape"
ctx = sccg_b200.Context(local)
h_ref = torch.from_numpy(ref_np).pin_memory()
inter, mode = ctx.compress(h_ref.numpy(), tgt_np, header)
cap = size + size // 50 + 4096
h_out = torch.empty(cap // max(1, world) + (64 << 20), dtype=torch.uint8).pin_memory()
times = []
for it in range(6):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    off, n, total = ctx.decompress_part(h_ref.numpy(), inter, rank, world, h_out.data_ptr(), h_out.numel())
    dist.barrier(); torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
# check: every rank hashes its piece, rank 0 compares with the pieces of its own full decompression
digest = hashlib.sha256(bytes(h_out[:n].numpy())).digest()
info = [None] * world
dist.all_gather_object(info, (off, n, digest))
if rank == 0:
    t1 = []
    h_full = torch.empty(cap, dtype=torch.uint8).pin_memory()
    for it in range(4):
        t0 = time.perf_counter()
        full_len = ctx.decompress_into((ctypes.c_char * size).from_address(h_ref.data_ptr()), inter, h_full.data_ptr(), h_full.numel())
        t1.append(time.perf_counter() - t0)
    whole = h_full[:full_len].numpy()
    ok = sum(x[1] for x in info) == full_len and all(hashlib.sha256(bytes(whole[o:o + m])).digest() == d for o, m, d in info)
    print(json.dumps({"n_gpus": world, "bp": size, "pieces_ms": 1e3 * min(times[2:]), "single_gpu_ms": 1e3 * min(t1[1:]), "Gbp_s": size / min(times[2:]) / 1e9,
                      "pieces_identical_to_whole": bool(ok)}), flush=True)
ctx.close()
dist.destroy_process_group()
