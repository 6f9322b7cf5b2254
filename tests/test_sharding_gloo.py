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


# ------------------------------------------------------------------------------------------------
# one pair over several ranks (segment-range shards): carries planned from the border reports, parts
# concatenated on rank 0 -> must be the unsharded file, byte for byte.  Kernels run in the SIMT emulator.
# ------------------------------------------------------------------------------------------------
def _shard_cases():
    import random
    from cases import rnd
    out = []
    ref, tgt = synth.local_pair(40_000, synth.seed_for(2, 61))
    out.append(("local_synth", ref.tobytes(), tgt.tobytes()))
    r = rnd(30_000, "shard")
    t = bytearray(r)
    rr = random.Random("shardm")
    for p in rr.sample(range(len(t)), 60):
        t[p] = rr.choice(b"ACGT")
    # lowercase runs that touch / cross / cover shard borders (world 2: border 15000; world 3: 10000, 20000)
    for a, b in ((9_990, 10_000), (14_000, 16_000), (19_999, 20_001), (29_990, 30_000), (0, 7)):
        t[a:b] = bytes(t[a:b]).lower()
    out.append(("runs_on_borders", r, bytes(t)))
    t2 = bytearray(r.lower())                                           # one run over everything
    out.append(("all_lowercase", r, bytes(t2)))
    t3 = bytearray(r)
    t3[10_000:20_000] = bytes(t3[10_000:20_000]).lower()                # a run that is exactly one middle shard (world 3)
    t3[9_000:9_990] = rnd(990, "nomatch")                               # segments without any match next to a border
    out.append(("run_equals_shard", r, bytes(t3)))
    out.append(("target_longer", r[:20_000], bytes(t[:27_345])))        # leftover target segments go to the last shard
    out.append(("reference_longer", r, bytes(t[:18_500])))
    t4 = bytearray(r)
    for i in range(12, 20):                                             # unrelated stretch: T2 abort -> global mode (fallback path)
        t4[i * 1000:(i + 1) * 1000] = rnd(1000, ("abort", i))
    out.append(("abort_to_global", r, bytes(t4)))
    t5 = bytearray(t); t5[12_345:12_348] = b"(7,"
    out.append(("paren_fallback", r, bytes(t5)))
    return out


def _shard_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from emu_lib import emu_context
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = emu_context()
    res = {}
    for name, ref, tgt in _shard_cases():
        out = sharding.compress_sharded(ctx, ref, tgt, b">sharded " + name.encode())
        if rank == 0:
            res[name] = out
            res[name + ":path"] = sharding.last_path
    # and back: every rank writes its piece of the reconstructed file at its offset
    import tempfile
    tmp = os.environ["SCCG_TEST_TMP"]
    for name, ref, tgt in _shard_cases()[:3]:
        inter = res.get(name, (None,))[0] if rank == 0 else None
        box = [inter]
        dist.broadcast_object_list(box, src=0)
        path = os.path.join(tmp, name + ".fa")
        sharding.decompress_sharded(ctx, ref, box[0], path)
    ctx.close()
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_compress_sharded_matches_unsharded(world, tmp_path, monkeypatch):
    monkeypatch.setenv("SCCG_TEST_TMP", str(tmp_path))
    monkeypatch.setenv("SCCG_PIPE_CHUNK", "4096")
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle_lib as ol
    from emu_lib import emu_context
    emu_context().close()                                               # build the emulator library once, before the ranks start
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for name, ref, tgt in _shard_cases():
        rc, exp, mode = ol.orc_compress(ref, tgt, b">sharded " + name.encode())
        assert rc == 0
        assert got[name] == (exp, mode), name
        fallback = name in ("abort_to_global", "paren_fallback") or (world == 3 and name in ("target_longer", "reference_longer"))
        assert got[name + ":path"] == ("unsharded" if fallback else "sharded"), name
    for name, ref, tgt in _shard_cases()[:3]:                           # the pieces written by the ranks form the reference's output file
        rc, exp = ol.orc_decompress(ref, got[name][0])
        assert rc == 0 and (tmp_path / (name + ".fa")).read_bytes() == exp, name


def test_plan_carries_unit():
    ranges = [(0, 10), (10, 20), (20, 30)]
    base = {"abort_inside": 0, "has_paren": 0, "head_status": [0] * 4, "tail_status": [0] * 4, "has_match": 1, "last_p": 0, "n_runs": 0,
            "first_run_start": 0, "first_run_len": 0, "last_run_start": 0, "last_run_len": 0}
    a = dict(base, last_p=9_500, n_runs=2, first_run_start=100, first_run_len=5, last_run_start=9_000, last_run_len=1_000)
    b = dict(base, has_match=0, n_runs=1, first_run_start=10_000, first_run_len=10_000, last_run_start=10_000, last_run_len=10_000)
    c = dict(base, last_p=29_000, n_runs=2, first_run_start=20_000, first_run_len=7, last_run_start=29_999, last_run_len=1)
    plan = sharding.plan_carries([a, b, c], ranges, 30_000)
    assert plan[0] == {"prev_p": 0, "skip_first_run": 0, "extra_last_len": 10_007, "prev_run_start": 0, "last_run_reaches_end": 0, "reserved": 0}
    assert plan[1]["skip_first_run"] == 1 and plan[1]["prev_p"] == 9_500
    assert plan[2] == {"prev_p": 9_500, "skip_first_run": 1, "extra_last_len": 0, "prev_run_start": 9_000, "last_run_reaches_end": 1, "reserved": 0}
    # a border window: fail, bad, fail, fail | fail -> abort
    a2 = dict(a, tail_status=[3, 1, 3, 3]); b2 = dict(b, head_status=[3, 0, 0, 0])
    assert sharding.plan_carries([a2, b2, c], ranges, 30_000) is None
    b3 = dict(b, head_status=[1, 0, 0, 0])                                 # the window must END with a failed segment
    assert sharding.plan_carries([a2, b3, c], ranges, 30_000) is not None
