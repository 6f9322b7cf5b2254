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
