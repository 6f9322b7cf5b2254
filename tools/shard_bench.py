#!/usr/bin/env python
"""One chromosome over N GPUs (segment-range shards): end-to-end time of sharding.compress_sharded on the
chr1-sized synthetic pair versus sccg_compress on one GPU, host buffers in, encoded file on rank 0 out.  Launch with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/shard_bench.py [size]"""
import json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist
import sccg_b200
from sccg_genome_compression_b200 import sharding, synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
size = int(sys.argv[1]) if len(sys.argv) > 1 else synth.CHR1_LEN
ref_np, tgt_np = synth.local_pair(size, synth.seed_for(2, 0))
# page-locked host copies (full PCIe speed); slices of them are zero-copy views
h_ref = torch.from_numpy(ref_np).pin_memory(); h_tgt = torch.from_numpy(tgt_np).pin_memory()
ref, tgt = h_ref.numpy(), h_tgt.numpy()
header = b">chr1 synthetic hg19-vs-hg18 shape"
ctx = sccg_b200.Context(local)
times = []
for it in range(6):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = sharding.compress_sharded(ctx, ref, tgt, header)
    dist.barrier(); torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
single = None
if rank == 0:
    ts = []
    for it in range(4):
        t0 = time.perf_counter(); one, mode = ctx.compress(ref, tgt, header); ts.append(time.perf_counter() - t0)
    single = min(ts[1:])
    assert out[0] == one, "sharded output differs from the single-GPU output"
    print(json.dumps({"n_gpus": world, "bp": len(tgt), "path": sharding.last_path, "sharded_ms": 1e3 * min(times[2:]), "single_gpu_ms": 1e3 * single,
                      "sharded_Mbp_s": len(tgt) / min(times[2:]) / 1e6, "note": "page-locked host buffers; includes the all_gather of the border reports and the gather of the parts to rank 0"}), flush=True)
ctx.close()
dist.destroy_process_group()
