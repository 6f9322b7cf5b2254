"""Segment-range shards on the GPU (`-m gpu`): two ranks (gloo transport) that share cuda:0 run sccg_shard_match /
sccg_shard_write through the C ABI; rank 0's concatenation must be the unsharded file byte for byte."""
import os
import sys

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import sccg_b200  # noqa: F401
from sccg_genome_compression_b200 import sharding, synth
from test_sharding_gloo import _free_port, _shard_cases

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import sccg_b200 as sb
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = sb.Context(0)
    res = {}
    cases = _shard_cases()
    ref, tgt = synth.local_pair(6_000_000, synth.seed_for(2, 71))
    cases.append(("local_6mbp", ref.tobytes(), tgt.tobytes()))
    for name, ref, tgt in cases:
        out = sharding.compress_sharded(ctx, ref, tgt, b">sharded " + name.encode())
        if rank == 0:
            res[name] = out
            res[name + ":path"] = sharding.last_path
    ctx.close()
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_compress_sharded_on_gpu(world):
    import oracle_lib as ol
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    cases = _shard_cases()
    ref, tgt = synth.local_pair(6_000_000, synth.seed_for(2, 71))
    cases.append(("local_6mbp", ref.tobytes(), tgt.tobytes()))
    for name, ref, tgt in cases:
        rc, exp, mode = ol.orc_compress(ref, tgt, b">sharded " + name.encode())
        assert rc == 0
        assert got[name] == (exp, mode), name
    assert got["local_6mbp:path"] == "sharded"
