"""Pins the C oracle against the compiled UNMODIFIED reference (oracle/_ref, built by
oracle/Makefile from /root/reference) at function level on seeded random inputs.  Runs wherever
oracle/_ref exists (build container, and the GPU box because the built files travel); skipped
otherwise -- the committed goldens (test_oracle_golden.py) cover that situation."""
import random
import tempfile
from pathlib import Path

import pytest

import oracle_lib as ol
from cases import rnd

pytestmark = pytest.mark.skipif(not ol.have_reference(), reason="oracle/_ref not built (needs /root/reference)")


def _pair(seed, n, alphabet=b"ACGT", snp=0.01, indel=0.0):
    r = random.Random(repr(seed))
    ref = bytes(r.choice(alphabet) for _ in range(n))
    t = bytearray()
    for c in ref:
        x = r.random()
        if x < snp:
            t.append(r.choice(alphabet))
        elif x < snp + indel / 2:
            continue
        elif x < snp + indel:
            t.append(c); t.append(r.choice(alphabet))
        else:
            t.append(c)
    return ref, bytes(t)


@pytest.mark.parametrize("seed", range(12))
def test_match_sequences_local(seed):
    alphabet = [b"ACGT", b"AC", b"ACGTN", b"A"][seed % 4]
    ref, tgt = _pair(("loc", seed), random.Random(seed).randint(5, 1000), alphabet, snp=0.02, indel=0.004)
    for k in (14, 10):
        assert ol.orc_match_sequences(ref, tgt, k, 0, False, 7000) == ol.ref_match_sequences(ref, tgt, k, 0, False, 7000)


@pytest.mark.parametrize("seed", range(10))
def test_match_sequences_global(seed):
    alphabet = [b"ACGT", b"ACGT", b"AC"][seed % 3]
    n = 20000 if alphabet == b"ACGT" else 3000
    ref, tgt = _pair(("glob", seed), n, alphabet, snp=0.01, indel=0.002)
    r = random.Random(seed)
    if seed % 2:                       # rearrange blocks so the parse gets lost and re-syncs by chance
        cut = sorted(r.sample(range(len(tgt)), 4))
        parts = [tgt[:cut[0]], tgt[cut[0]:cut[1]], tgt[cut[1]:cut[2]], tgt[cut[2]:cut[3]], tgt[cut[3]:]]
        r.shuffle(parts)
        tgt = b"".join(parts)
    assert ol.orc_match_sequences(ref, tgt, 14, 100, True, 0) == ol.ref_match_sequences(ref, tgt, 14, 100, True, 0)


@pytest.mark.parametrize("seed", range(6))
def test_compress_and_reconstruct_end_to_end(seed):
    r = random.Random(repr(("e2e", seed)))
    n = r.randint(3000, 40000)
    ref, tgt = _pair(("e2e", seed), n, b"ACGT", snp=0.003, indel=0.0005 if seed % 2 else 0.0)
    t = bytearray(tgt)
    for _ in range(8):                 # lowercase + N runs
        a = r.randrange(len(t)); b = min(len(t), a + r.choice([1, 2, 30, 700]))
        t[a:b] = bytes(t[a:b]).lower()
        a = r.randrange(len(t)); b = min(len(t), a + r.choice([1, 5, 120]))
        t[a:b] = b"N" * (b - a)
    tgt = bytes(t)
    res = ol.ref_roundtrip(ref, tgt, b">e2e %d" % seed)
    rc, text, mode = ol.orc_compress(ref, tgt, b">e2e %d" % seed)
    assert rc == 0 and text == res["intermediate"]
    rc, out = ol.orc_decompress(ref, text)
    assert res["rc_decompress"] == 0 and rc == 0 and out == res["reconstructed"]
    # function-level reconstruct_genome as well
    header, low, nline, body = ol.split_intermediate(text)
    prepared = ol.prepare_reference(ref, nline)
    rc_r, out_r, _ = ol.ref_reconstruct(prepared, body, b"" if nline == b"," else nline, low)
    rc_o, out_o = ol.orc_reconstruct(prepared, body, b"" if nline == b"," else nline, low)
    assert rc_r == 0 and rc_o == 0 and out_r == out_o


@pytest.mark.parametrize("text", [
    b">h\n\n,\n(0,500)(7,(503,497)(1000,1000)",
    b"\n,\n(10,5)AC(4,6)(100,1)",
    b">chr1 (x,1)\n(5,2)\n,\n(7,7)(9,9)",
    b">h\n(3,4)9,\n(1,2)\n(100,5)(90,5)ACGT(2000000000,7)(-5,3)",
    b">h\n\n,\nACGT",
    b">h\n\n,\n(12)(3,4)(5",
])
def test_delta_encode_text(text):
    with tempfile.TemporaryDirectory() as d:
        p = Path(d) / "f.txt"
        p.write_bytes(text)
        assert ol.ref().sccg_ref_delta_encode(str(p).encode()) == 0
        expect = p.read_bytes()
    rc, out = ol.orc_delta_encode(text)
    assert rc == 0 and out == expect


def test_reconstruct_malformed_inputs_agree():
    ref = rnd(500, "mal")
    for enc, nidx, low in [
        (b"(0,100)ACGT(50,20)", b"", b"(3,4)10,200"),
        (b"(0,100)ACGT(50,20)", b"2,(5,3)(100,2)", b""),
        (b"AC(10,-3)GT", b"", b""),                      # negative length: substr count wraps -> rest of reference
        (b"(0,10)x5y", b"", b"0,1,1,"),
        (b"(5,5)", b"", b"(0,2),(4,2)7"),
    ]:
        rc_r, out_r, _ = ol.ref_reconstruct(ref, enc, nidx, low)
        rc_o, out_o = ol.orc_reconstruct(ref, enc, nidx, low)
        assert (rc_r, out_r) == (rc_o, out_o), (enc, nidx, low)
    for enc in [b"(a,5)", b"(5,)", b"(99999999999,5)"]:   # stoi throws
        rc_r, _, _ = ol.ref_reconstruct(ref, enc, b"", b"")
        rc_o, _ = ol.orc_reconstruct(ref, enc, b"", b"")
        assert rc_r == 1 and rc_o == 1


@pytest.mark.parametrize("seed", range(24))
def test_grammar_symbols_in_target_cli_level(seed):
    """literal '(' ')' ',' digits in the target (SURVEY N2 / experiment J): the oracle's text-level delta_encode against
    the reference program itself, incl. the stoi failure (exit 1, compressed_genome.txt left un-rewritten)"""
    from test_emu_compress import grammar_pair
    ref, tgt = grammar_pair(seed, make_global=(seed % 4 >= 2))
    header = b">gram (alt) 1,2"
    rc, exp, _ = ol.orc_compress(ref, tgt, header)
    with tempfile.TemporaryDirectory() as d:
        d = Path(d)
        ol.write_fasta(d / "r.fa", ref, b">ref")
        ol.write_fasta(d / "t.fa", tgt, header)
        rc_ref, inter = ol.ref_compress_cli(d / "r.fa", d / "t.fa", d / "out")
    assert (rc != 0) == (rc_ref != 0)
    assert inter == exp


@pytest.mark.parametrize("shape", ["shifted_tail", "leftover_only", "reference_insertion", "small_shift"])
def test_length_mismatch_shapes_end_to_end(shape):
    """pairs of different length, > 160 segment pairs: the shapes for which the GPU path probes the last segments for a T2
    abort before the bulk launch.  The oracle's verdict (local vs global) and bytes are the reference's."""
    n = 200_000
    ref = rnd(n, "probe")
    r = random.Random("probe" + shape)
    t = bytearray(ref)
    for p in r.sample(range(n), 150):
        t[p] = r.choice(b"ACGT")
    if shape == "shifted_tail":
        tgt = bytes(t[:60_000]) + bytes(t[63_000:])
    elif shape == "leftover_only":
        tgt = bytes(t) + rnd(5_000, "extra")
    elif shape == "reference_insertion":
        tgt = bytes(t); ref = ref[:20_000] + rnd(2_500, "ins") + ref[20_000:]
    else:
        tgt = bytes(t[:90_000]) + b"ACGTTGCAAC" * 3 + bytes(t[90_000:]) + rnd(1_200, "tail")
    res = ol.ref_roundtrip(ref, tgt, b">probe")
    rc, text, mode = ol.orc_compress(ref, tgt, b">probe")
    assert rc == 0 and text == res["intermediate"]
    assert mode == (1 if shape in ("shifted_tail", "reference_insertion") else 0)
    rc, out = ol.orc_decompress(ref, text)
    assert res["rc_decompress"] == 0 and rc == 0 and out == res["reconstructed"]
