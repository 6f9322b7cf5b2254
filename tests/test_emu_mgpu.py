"""Host logic of the C++ multi-GPU layer (csrc/sccg_mgpu.cuh) on the CPU: the ranks are host threads of one process, the
transport is the in-process hub of the emulator build (the product build uses NCCL; GPU tests cover that).  Checks: LPT
assignment, whole-genome gather of the encoded streams to rank 0, one pair sharded by segment range (byte-identical to the
unsharded file, incl. the fallbacks), decompression by output range."""
import ctypes as C
import threading

import pytest

import oracle_lib as ol
import sccg_b200
from cases import rnd
from emu_lib import EMU_LIB, emu_context
from sccg_genome_compression_b200 import sharding, synth


BACKEND = {"lib": EMU_LIB, "devices": None}     # test_gpu_mgpu.py points this at the product library / real devices


def run_ranks(world, body):
    """runs body(rank, mg, ctx) on `world` threads, one context + communicator each; returns the results by rank"""
    lib = BACKEND["lib"]
    if lib == EMU_LIB:
        emu_context().close()                  # builds the emulator library
    uid = sccg_b200.mgpu_unique_id(lib)
    out, errs = [None] * world, []

    def worker(r):
        ctx = mg = None
        try:
            ctx = sccg_b200.Context(BACKEND["devices"][r] if BACKEND["devices"] else 0, lib_path=lib)
            mg = sccg_b200.Mgpu(ctx, uid, r, world)
            out[r] = body(r, mg, ctx)
        except BaseException as e:          # noqa: BLE001
            errs.append((r, e))
        finally:
            if mg: mg.close()
            if ctx: ctx.close()
    assert BACKEND["devices"] is None or world <= len(BACKEND["devices"]), "more ranks than devices"
    th = [threading.Thread(target=worker, args=(r,), daemon=True) for r in range(world)]
    for t in th: t.start()
    for t in th: t.join(timeout=300)
    assert not any(t.is_alive() for t in th), "a rank is stuck"
    assert not errs, errs
    return out


def test_assign_matches_python_lpt():
    if BACKEND["lib"] == EMU_LIB:
        emu_context().close()
    for world in (1, 2, 3, 8):
        for lengths in (synth.HG19_LENGTHS, [5, 5, 5, 5], [10], [], [3, 9, 9, 1, 7, 7, 2]):
            owner = sccg_b200.mgpu_assign(list(lengths), world, BACKEND["lib"])
            expect = sharding.assign_chromosomes(list(lengths), world)
            for r, idxs in enumerate(expect):
                assert all(owner[i] == r for i in idxs)


@pytest.mark.parametrize("world", [1, 2, 3])
def test_whole_genome_gather(world):
    """7 small pairs (local and global mode), LPT over the ranks, streams gathered to rank 0 == the oracle's files"""
    pairs = []
    for i in range(7):
        if i == 3:
            r, t = synth.global_gap_pair(26_000, 24_000, synth.seed_for(1, 40 + i))
        else:
            r, t = synth.local_pair(12_000 + 3_000 * i, synth.seed_for(2, 40 + i))
        pairs.append((r.tobytes(), t.tobytes(), b">pair %d" % i))
    owner = sccg_b200.mgpu_assign([len(p[1]) for p in pairs], world, BACKEND["lib"])

    def body(rank, mg, ctx):
        for rep in range(2):                                  # the communicator is reusable
            for i, (r, t, h) in enumerate(pairs):
                if owner[i] == rank:
                    mg.compress_item(i, r, t, h)
            got = mg.gather()
        # the device-side variant: same streams, left in rank 0's device memory
        for i, (r, t, h) in enumerate(pairs):
            if owner[i] == rank:
                mg.compress_item(i, r, t, h)
        dev = mg.gather_device()
        if rank == 0:
            ptr, where = dev
            assert {i: ctx.download(ptr + o, n) for i, (o, n) in where.items()} == got
        return got
    res = run_ranks(world, body)
    assert all(x is None for x in res[1:])
    assert sorted(res[0]) == list(range(7))
    for i, (r, t, h) in enumerate(pairs):
        rc, exp, mode = ol.orc_compress(r, t, h)
        assert rc == 0 and res[0][i] == exp, i


def _sharded(world, ref, tgt, header):
    def body(rank, mg, ctx):
        return mg.compress_sharded(ref, tgt, header)
    res = run_ranks(world, body)
    assert all(x[0] is None for x in res[1:])
    return res[0]


@pytest.mark.parametrize("world", [2, 3])
def test_compress_sharded_local_identical_to_unsharded(world):
    ref, tgt = synth.local_pair(61_000, synth.seed_for(2, 55))
    ref, tgt = ref.tobytes(), tgt.tobytes()
    tgt = tgt + rnd(2_345, "leftover")                                             # leftover target segments (:476-481)
    rc, exp, mode = ol.orc_compress(ref, tgt, b">sharded")
    assert rc == 0 and mode == 0
    got, gmode, sharded = _sharded(world, ref, tgt, b">sharded")
    assert sharded and gmode == 0 and got == exp


@pytest.mark.parametrize("world", [2, 3])
def test_compress_sharded_runs_across_borders(world):
    """lowercase runs that cross (or fill) shard borders, a run reaching the end, no header"""
    n = 48_000
    ref = rnd(n, "bord")
    t = bytearray(ref)
    cuts = [n * (r + 1) // world // 1000 * 1000 for r in range(world - 1)]
    for c in cuts:
        t[c - 150:c + 2300] = bytes(t[c - 150:c + 2300]).lower()                    # crosses the border
    t[n - 500:] = bytes(t[n - 500:]).lower()                                        # reaches the end
    if world == 3:
        a, b = n // 3 // 1000 * 1000, 2 * n // 3 // 1000 * 1000
        t[a - 10:b + 10] = bytes(t[a - 10:b + 10]).lower()                          # covers the whole middle shard
    tgt = bytes(t)
    rc, exp, mode = ol.orc_compress(ref, tgt, b"")
    assert rc == 0 and mode == 0
    got, gmode, sharded = _sharded(world, ref, tgt, b"")
    assert sharded and got == exp


def test_compress_sharded_fallbacks():
    # too small to shard
    ref, tgt = synth.local_pair(9_000, synth.seed_for(2, 56))
    rc, exp, mode = ol.orc_compress(ref.tobytes(), tgt.tobytes(), b">tiny")
    got, gmode, sharded = _sharded(2, ref.tobytes(), tgt.tobytes(), b">tiny")
    assert not sharded and (gmode, got) == (mode, exp)
    # T2 abort inside a shard -> global mode on rank 0
    ref, tgt = synth.global_gap_pair(52_000, 48_000, synth.seed_for(1, 56))
    rc, exp, mode = ol.orc_compress(ref.tobytes(), tgt.tobytes(), b">gap")
    got, gmode, sharded = _sharded(2, ref.tobytes(), tgt.tobytes(), b">gap")
    assert mode == 1 and not sharded and (gmode, got) == (mode, exp)
    # '(' in the target -> text-level delta on rank 0
    ref, tgt = synth.local_pair(40_000, synth.seed_for(2, 57))
    t = bytearray(tgt.tobytes()); t[20_500:20_503] = b"(7,"
    rc, exp, mode = ol.orc_compress(ref.tobytes(), bytes(t), b">paren")
    if rc == 0:
        got, gmode, sharded = _sharded(2, ref.tobytes(), bytes(t), b">paren")
        assert not sharded and (gmode, got) == (mode, exp)


def test_decompress_sharded_pieces(world=3):
    ref, tgt = synth.local_pair(150_000, synth.seed_for(2, 58))
    ref, tgt = ref.tobytes(), tgt.tobytes()
    rc, inter, mode = ol.orc_compress(ref, tgt, b">parts")
    rc, exp = ol.orc_decompress(ref, inter)

    def body(rank, mg, ctx):
        buf = C.create_string_buffer(len(exp) + 64)
        off, n, total = mg.decompress_sharded(ref, inter, C.cast(buf, C.c_void_p), len(exp) + 64)
        return off, buf.raw[:n], total
    import os
    os.environ["SCCG_PIPE_CHUNK"] = "20000"
    try:
        res = run_ranks(world, body)
    finally:
        del os.environ["SCCG_PIPE_CHUNK"]
    image = bytearray(len(exp))
    for off, piece, total in res:
        assert total == len(exp)
        image[off:off + len(piece)] = piece
    assert bytes(image) == exp
