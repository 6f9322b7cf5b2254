"""Host-side multi-GPU logic on the CPU: LPT chromosome assignment and the gather of the encoded
record streams to rank 0 over torch.distributed (gloo, world_size 2 and 3)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sccg_b200  # noqa: F401  (registers sccg_genome_compression_b200)
from sccg_genome_compression_b200 import sharding, synth


def test_lpt_assignment_hg19():
    for world in (1, 2, 4, 8):
        parts = sharding.assign_chromosomes(synth.HG19_LENGTHS, world)
        assert sorted(i for p in parts for i in p) == list(range(24))
        loads = [sum(synth.HG19_LENGTHS[i] for i in p) for p in parts]
        assert max(loads) <= 1.12 * sum(loads) / world           # LPT: balanced to within the largest chromosome
    assert sharding.assign_chromosomes([5, 5, 5], 2) == [[0, 2], [1]]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = [7, 3, 9, 1, 4, 4, 8]
    mine = sharding.assign_chromosomes(lengths, world)[rank]
    local = {i: (b"chr%d:" % i) + bytes([65 + i]) * (lengths[i] * 1000 + i) for i in mine}
    if rank == world - 1:
        local[99] = b""                                              # empty stream survives the gather
    got = sharding.gather_streams(local, dst=0)
    if rank == 0:
        q.put({k: (len(v), v[:8]) for k, v in got.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_streams_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lengths = [7, 3, 9, 1, 4, 4, 8]
    assert sorted(got) == list(range(7)) + [99]
    for i in range(7):
        assert got[i] == (len(b"chr%d:" % i) + lengths[i] * 1000 + i, ((b"chr%d:" % i) + bytes([65 + i]) * 8)[:8])
    assert got[99] == (0, b"")
