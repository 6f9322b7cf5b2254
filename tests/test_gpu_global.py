"""GPU parity tests of global mode (B200, `-m gpu`): index build + banded parse through the C ABI
versus the C oracle, function level and whole-file level."""
import random

import pytest

import oracle_lib as ol
from cases import rnd
from test_emu_compress import _mutated_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import sccg_b200
    c = sccg_b200.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("seed", range(16))
def test_match_sequences_global_vs_oracle(ctx, seed):
    alphabet = [b"ACGT", b"ACGT", b"AC", b"ACGT"][seed % 4]
    n = 40000 if alphabet == b"ACGT" else 4000
    ref, tgt = _mutated_pair(("glob", seed), n, alphabet, snp=0.01, indel=0.002)
    r = random.Random(seed)
    if seed % 2:
        cut = sorted(r.sample(range(len(tgt)), 4))
        parts = [tgt[:cut[0]], tgt[cut[0]:cut[1]], tgt[cut[1]:cut[2]], tgt[cut[2]:cut[3]], tgt[cut[3]:]]
        r.shuffle(parts)
        tgt = b"".join(parts)
    if seed % 3 == 0:
        tgt = rnd(r.randint(1, 3000), ("pre", seed)) + tgt
    exp = [(x.p, x.l, x.lit) for x in ol.orc_match_sequences(ref, tgt, 14, 100, True, 0)]
    got = [(x.p, x.l, x.lit) for x in ctx.match_sequences(ref, tgt, 14, 100, True, 0)]
    assert got == exp


@pytest.mark.parametrize("k,m", [(8, 0), (10, 5), (16, 120), (12, 37)])
def test_match_sequences_global_other_parameters(ctx, k, m):
    ref, tgt = _mutated_pair(("gp", k, m), 20000, b"ACG", snp=0.02, indel=0.003)
    exp = [(x.p, x.l, x.lit) for x in ol.orc_match_sequences(ref, tgt, k, m, True, 11)]
    got = [(x.p, x.l, x.lit) for x in ctx.match_sequences(ref, tgt, k, m, True, 11)]
    assert got == exp


def test_global_degenerate_inputs(ctx):
    ref = rnd(3000, "gd")
    for tgt in (b"", ref[:5], ref[:14], rnd(500, "unrelated"), ref, b"A" * 400):
        for rr in (ref, b"", ref[:10], b"A" * 300):
            exp = [(x.p, x.l, x.lit) for x in ol.orc_match_sequences(rr, tgt, 14, 100, True, 0)]
            got = [(x.p, x.l, x.lit) for x in ctx.match_sequences(rr, tgt, 14, 100, True, 0)]
            assert got == exp, (len(rr), len(tgt))


@pytest.mark.parametrize("shape,n", [("gap", 6_000_000), ("divergent", 4_000_000)])
def test_compress_global_synthetic_vs_oracle(ctx, shape, n):
    """BASELINE configs[0] / configs[2] shapes at a size the oracle finishes in seconds; plus decode round trip"""
    from sccg_genome_compression_b200 import synth
    if shape == "gap":
        ref, tgt = synth.global_gap_pair(int(n * 1.08), n, synth.seed_for(1, 5))
    else:
        ref, tgt = synth.divergent_pair(n, synth.seed_for(3, 5))
    ref, tgt = ref.tobytes(), tgt.tobytes()
    rc, exp, mode = ol.orc_compress(ref, tgt, b">global synthetic")
    assert rc == 0 and mode == 1
    got, gmode = ctx.compress(ref, tgt, b">global synthetic")
    assert gmode == 1 and got == exp
    back = ctx.decompress(ref, got)
    rc, oback = ol.orc_decompress(ref, exp)
    assert rc == 0 and back == oback


def _compress_device(ctx, ref: bytes, tgt: bytes, header: bytes):
    import torch
    dr = torch.frombuffer(bytearray(ref + bytes(64)), dtype=torch.uint8).cuda()
    dt = torch.frombuffer(bytearray(tgt + bytes(64)), dtype=torch.uint8).cuda()
    ptr, n, mode = ctx.compress_device(dr.data_ptr(), len(ref), dt.data_ptr(), len(tgt), header)
    return ctx.download(ptr, n), mode, ctx.profile()["launches"]


@pytest.mark.parametrize("shape", ["gap", "divergent", "shifted_tail", "leftover_only", "small_shift"])
def test_compress_device_abort_probe(ctx, shape):
    """device-resident entry point on pairs of different length: the abort probe over the last segments runs first
    (DESIGN 4.7c).  Aborting or not, the file is the oracle's."""
    from sccg_genome_compression_b200 import synth
    if shape == "gap":
        ref, tgt = synth.global_gap_pair(3_240_000, 3_000_000, synth.seed_for(1, 9)); ref, tgt = ref.tobytes(), tgt.tobytes()
    elif shape == "divergent":
        ref, tgt = synth.divergent_pair(2_000_000, synth.seed_for(3, 9)); ref, tgt = ref.tobytes(), tgt.tobytes()
    else:
        ref, t = synth.local_pair(2_000_000, synth.seed_for(2, 9)); ref, t = ref.tobytes(), t.tobytes()
        if shape == "shifted_tail":
            tgt = t[:600_000] + t[603_000:]                               # a 3 kb deletion: the rest is three segments off
        elif shape == "leftover_only":
            tgt = t + rnd(5_000, "extra")                                 # no shift, leftover target segments
        else:
            tgt = t[:900_000] + b"ACGTTGCAAC" * 3 + t[900_000:] + rnd(1_200, "tail")    # 30 symbols off: still matches inside the segments
    rc, exp, mode = ol.orc_compress(ref, tgt, b">probe")
    assert rc == 0
    got, gmode, _ = _compress_device(ctx, ref, tgt, b">probe")
    assert (gmode, got) == (mode, exp)
    assert mode == (0 if shape in ("leftover_only", "small_shift") else 1)


@pytest.mark.parametrize("bits", [8, 20, 24])
def test_global_index_bucket_table_widths(ctx, bits, monkeypatch):
    """offset table over the sorted index at forced widths (24 = one bucket per hash value)"""
    from sccg_genome_compression_b200 import synth
    monkeypatch.setenv("SCCG_GP_BUCKET_BITS", str(bits))
    ref, tgt = synth.divergent_pair(1_500_000, synth.seed_for(3, 11))
    ref, tgt = ref.tobytes(), tgt.tobytes()
    rc, exp, mode = ol.orc_compress(ref, tgt, b">bkt")
    got, gmode = ctx.compress(ref, tgt, b">bkt")
    assert rc == 0 and (gmode, got) == (mode, exp)


@pytest.mark.parametrize("seed", range(6))
def test_global_lost_scan(ctx, seed, monkeypatch):
    """lost state + gp_lost_scan_k (see test_emu_compress.lost_scan_pair), default chunk size and a long unrelated stretch"""
    from test_emu_compress import lost_scan_pair
    ref, tgt = lost_scan_pair(seed, junk=400_000)
    exp = [(x.p, x.l, x.lit) for x in ol.orc_match_sequences(ref, tgt, 14, 100, True, 0)]
    got = [(x.p, x.l, x.lit) for x in ctx.match_sequences(ref, tgt, 14, 100, True, 0)]
    assert got == exp
    assert ctx.profile()["spec_rounds"] >= 2


@pytest.mark.parametrize("batch0", [1, 3, 40])
@pytest.mark.parametrize("shape", ["gap", "divergent"])
def test_global_two_batch_speculation(ctx, shape, batch0, monkeypatch):
    """tiny first speculation batch: the front reaches the second batch while it is still running (GP_WAIT), or cancels it when
    the parse gets lost early; either way the file is the oracle's"""
    from sccg_genome_compression_b200 import synth
    monkeypatch.setenv("SCCG_GP_BATCH0", str(batch0))
    if shape == "gap":
        ref, tgt = synth.global_gap_pair(3_240_000, 3_000_000, synth.seed_for(1, 13))
    else:
        ref, tgt = synth.divergent_pair(2_500_000, synth.seed_for(3, 13))
    ref, tgt = ref.tobytes(), tgt.tobytes()
    rc, exp, mode = ol.orc_compress(ref, tgt, b">two batches")
    assert rc == 0 and mode == 1
    for rep in range(3):                                     # timing varies from run to run: the result must not
        got, gmode = ctx.compress(ref, tgt, b">two batches")
        assert (gmode, got) == (mode, exp)



@pytest.mark.parametrize("stride", [1, 4, 8])
@pytest.mark.parametrize("shape", ["gap", "divergent"])
def test_global_index_modes(ctx, shape, stride, monkeypatch):
    """index of every reference k-mer (stride 1) against the sampled index references of a million k-mers and more get (8):
    same file, and the asked-for mode is the one that produced it"""
    from sccg_genome_compression_b200 import synth
    monkeypatch.setenv("SCCG_GP_STRIDE", str(stride))
    if shape == "gap":
        ref, tgt = synth.global_gap_pair(2_160_000, 2_000_000, synth.seed_for(1, 17))
    else:
        ref, tgt = synth.divergent_pair(1_800_000, synth.seed_for(3, 17))
    ref, tgt = ref.tobytes(), tgt.tobytes()
    rc, exp, mode = ol.orc_compress(ref, tgt, b">index modes")
    assert rc == 0 and mode == 1
    got, gmode = ctx.compress(ref, tgt, b">index modes")
    assert (gmode, got) == (mode, exp)
    assert ctx.profile()["index_stride"] == stride


@pytest.mark.parametrize("shape", ["first_kmer_absent", "first_kmers_repeat", "long_unmatched_start", "first_kmer_everywhere", "p0_fallthrough", "first_kmer_twice"])
def test_global_sampled_index_falls_back(ctx, shape, monkeypatch):
    """lookups the sampled mode cannot serve (tests/test_emu_compress.sampled_fallback_pair) repeat the parse with the full index"""
    from test_emu_compress import sampled_fallback_pair
    monkeypatch.setenv("SCCG_GP_STRIDE", "8")
    ref, tgt, expect_stride = sampled_fallback_pair(shape)
    exp = [(x.p, x.l, x.lit) for x in ol.orc_match_sequences(ref, tgt, 14, 100, True, 0)]
    got = [(x.p, x.l, x.lit) for x in ctx.match_sequences(ref, tgt, 14, 100, True, 0)]
    assert got == exp
    assert ctx.profile()["index_stride"] == expect_stride
