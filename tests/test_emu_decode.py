"""Kernel-logic tests of the record-decode path on the CPU (SIMT emulator build) versus the golden
outputs of the compiled reference and the C oracle."""
import base64
import random
import zlib

import pytest

import oracle_lib as ol
import sccg_b200
from cases import cases, fasta_cases, rnd
from emu_lib import emu_context

CASES = cases()


@pytest.fixture(scope="module")
def ctx():
    c = emu_context()
    yield c
    c.close()


def unpack(s):
    return zlib.decompress(base64.b64decode(s))


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_decompress_matches_golden(ctx, case, golden):
    g = golden["cases"][case.name]
    inter = unpack(g["intermediate_z"])
    if g["rc_decompress"] != 0:
        with pytest.raises(sccg_b200.SccgError) as e:
            ctx.decompress(case.ref, inter)
        assert e.value.code in (sccg_b200.SCCG_E_FORMAT, sccg_b200.SCCG_E_BOUNDS)
        return
    got = ctx.decompress(case.ref, inter)
    assert got == unpack(g["reconstructed_z"])


def test_fasta_level_goldens(ctx, golden):
    for fc in fasta_cases():
        g = golden["fasta_cases"][fc.name]
        ref = ol.orc_parse_reference_fasta(fc.ref_file)
        got = ctx.decompress(ref, unpack(g["intermediate_z"]))
        assert got == unpack(g["reconstructed_z"]), fc.name


def make_record_stream(seed):
    """hand-built record streams: random tokens / literals / N runs / lowercase runs"""
    r = random.Random(seed)
    ref = rnd(r.randint(50, 5000), ("rr", seed))
    body = bytearray(); prev = 0; length = 0
    for _ in range(r.randint(0, 300)):
        if r.random() < 0.5:
            lit = rnd(r.randint(1, 40), ("lit", seed, len(body)), b"ACGTNRYK")
            body += lit; length += len(lit)
        else:
            p = r.randrange(len(ref)); l = r.randint(0, min(700, len(ref) - p))
            body += b"(%d,%d)" % (p - prev, l); prev = p; length += l

    def runs(total, maxlen):
        out, pos, prev_s, items = bytearray(), 0, 0, []
        while pos < total and r.random() < 0.9 and len(items) < 40:
            s = pos + r.randint(0, max(1, total // 10)); l = r.choice([1, 1, 2, r.randint(1, maxlen)])
            if s + l > total:
                break
            items.append((s, l)); pos = s + l + 1
        for idx, (s, l) in enumerate(items):
            last = idx == len(items) - 1
            if l == 1:
                out += b"%d" % (s - prev_s) + (b"" if last and r.random() < 0.5 else b",")
            else:
                out += b"(%d,%d)" % (s - prev_s, l) + (b"," if r.random() < 0.2 else b"")
            prev_s = s
        return bytes(out), sum(l for _, l in items)
    nlist, nsum = runs(length, 50) if seed % 2 else (b"", 0)
    low, _ = runs(length + nsum, 300)
    return ref, bytes(body), nlist, low


@pytest.mark.parametrize("seed", range(40))
def test_reconstruct_random_vs_oracle(ctx, seed):
    ref, body, nlist, low = make_record_stream(seed)
    rc, exp = ol.orc_reconstruct(ref, body, nlist, low)
    if rc != 0:
        with pytest.raises(sccg_b200.SccgError):
            ctx.reconstruct(ref, body, nlist, low)
    else:
        assert ctx.reconstruct(ref, body, nlist, low) == exp


def test_reconstruct_errors(ctx):
    ref = rnd(500, "e")
    with pytest.raises(sccg_b200.SccgError) as e:
        ctx.reconstruct(ref, b"(400,200)", b"", b"")                   # decompression.cpp:223-229
    assert e.value.code == sccg_b200.SCCG_E_BOUNDS
    for enc in (b"(a,5)", b"(5,)", b"(99999999999,5)", b"AC(7,(503,497)"):   # stoi throws in the reference
        with pytest.raises(sccg_b200.SccgError) as e:
            ctx.reconstruct(ref, enc, b"", b"")
        assert e.value.code == sccg_b200.SCCG_E_FORMAT
    for nlist in (b"1000,", b"(6,5)"):                                       # N run past the end: UB in the reference
        with pytest.raises(sccg_b200.SccgError) as e:
            ctx.reconstruct(ref, b"ACGT", nlist, b"")
        assert e.value.code == sccg_b200.SCCG_E_FORMAT
    assert ctx.reconstruct(ref, b"ACGT", b"4,", b"") == b"ACGTN\n"              # trailing N is fine
    assert ctx.reconstruct(ref, b"", b"", b"") == b"\n"
    assert ctx.reconstruct(ref, b"ACGT", b"", b"1,") == b"AcGT\n"
    assert ctx.reconstruct(ref, b"ACGT", b"(1,2)", b"(0,3)") == b"annCGT\n"      # a lowercase run covers re-inserted N


@pytest.mark.parametrize("shape", ["local", "gap", "divergent", "divergent_big"])
def test_decompress_many_tiles_vs_oracle(ctx, shape):
    """multi-tile gather (search windows per CTA) on record streams produced by the oracle compressor"""
    from sccg_genome_compression_b200 import synth
    if shape == "local":
        ref, tgt = synth.local_pair(300_000, synth.seed_for(2, 21))
    elif shape == "gap":
        ref, tgt = synth.global_gap_pair(200_000, 180_000, synth.seed_for(1, 21))
    elif shape == "divergent_big":          # > 100 scan tiles: several look-back rounds of the single-pass scan
        ref, tgt = synth.divergent_pair(600_000, synth.seed_for(3, 22))
    else:
        ref, tgt = synth.divergent_pair(150_000, synth.seed_for(3, 21))
    ref, tgt = ref.tobytes(), tgt.tobytes()
    tgt = tgt[:5000] + b"N" * 7 + tgt[5007:90_000] + b"N" * 4300 + tgt[94_300:]      # N runs crossing tile borders
    rc, inter, mode = ol.orc_compress(ref, tgt, b">tiles")
    assert rc == 0
    rc, exp = ol.orc_decompress(ref, inter)
    assert rc == 0
    assert ctx.decompress(ref, inter) == exp


@pytest.mark.parametrize("chunk", [4096, 20000, 1 << 20])
def test_decompress_pipelined_small_chunks(ctx, chunk, monkeypatch):
    """the pipelined host path (reference uploaded / upper-cased chunk by chunk, text gathered and sent home chunk by
    chunk): every output chunk may only touch reference chunks that were prepared for it (poisoned otherwise)"""
    from sccg_genome_compression_b200 import synth
    monkeypatch.setenv("SCCG_PIPE_CHUNK", str(chunk))
    monkeypatch.setenv("SCCG_PIPE_POISON", "1")
    for shape in ("local", "gap"):
        if shape == "local":
            ref, tgt = synth.local_pair(260_000, synth.seed_for(2, 31))
            ref, tgt = ref.tobytes(), tgt.tobytes()
            tgt = tgt[:100_000] + tgt[100_000:100_900][::-1] + tgt[100_900:]          # a segment that falls back to literals
        else:
            ref, tgt = synth.global_gap_pair(150_000, 140_000, synth.seed_for(1, 31))
            ref, tgt = ref.tobytes(), tgt.tobytes()
        rc, inter, mode = ol.orc_compress(ref, tgt, b">pipelined")
        assert rc == 0
        rc, exp = ol.orc_decompress(ref, inter)
        assert rc == 0
        assert ctx.decompress(ref, inter) == exp
    # tokens that point far ahead / far back in the reference (hand-built stream, local-mode N line)
    ref = rnd(120_000, "pipe")
    inter = b">x\n\n,\n(110000,5000)(-110000,30000)ACGT(60000,100)(-50000,20000)"
    rc, exp = ol.orc_decompress(ref, inter)
    assert rc == 0 and ctx.decompress(ref, inter) == exp


@pytest.mark.parametrize("n_parts,chunk", [(2, 4096), (3, 4096), (5, 20000), (8, 1 << 20)])
def test_decompress_parts_concatenate_to_the_whole(ctx, n_parts, chunk, monkeypatch):
    """output-range sharding: the pieces of sccg_decompress_part, placed at their offsets, are the whole image; a piece only
    has the reference chunks it needs (everything else is poisoned)"""
    from sccg_genome_compression_b200 import synth
    monkeypatch.setenv("SCCG_PIPE_CHUNK", str(chunk))
    monkeypatch.setenv("SCCG_PIPE_POISON", "1")
    cases = []
    ref, tgt = synth.local_pair(150_000, synth.seed_for(2, 81))
    cases.append((ref.tobytes(), tgt.tobytes()))
    ref, tgt = synth.global_gap_pair(90_000, 80_000, synth.seed_for(1, 81))
    cases.append((ref.tobytes(), tgt.tobytes()))
    cases.append((rnd(3000, "tinyp"), rnd(3000, "tinyp")[:2100]))         # fewer chunks than parts: some pieces are empty
    for ref, tgt in cases:
        rc, inter, mode = ol.orc_compress(ref, tgt, b">parts")
        assert rc == 0
        rc, exp = ol.orc_decompress(ref, inter)
        assert rc == 0
        image = bytearray(len(exp))
        covered = 0
        for p in range(n_parts):
            off, piece, total = ctx.decompress_part(ref, inter, p, n_parts)
            assert total == len(exp)
            image[off:off + len(piece)] = piece
            covered += len(piece)
        assert covered == len(exp) and bytes(image) == exp


def test_far_out_of_range_tokens_fail_cleanly(ctx):
    import robustness_cases
    robustness_cases.check_far_out_of_range_tokens(ctx)


def test_long_header_line(ctx):
    import robustness_cases
    robustness_cases.check_long_header_line(ctx)


def test_stale_shard_write_is_rejected(ctx):
    import robustness_cases
    robustness_cases.check_stale_shard_write(ctx)


@pytest.mark.parametrize("chunk", [None, 20000])
def test_streaming_decompress_and_fasta_into(ctx, chunk):
    import robustness_cases
    robustness_cases.check_streaming_and_fasta_into(ctx, chunk)


def test_decoder_tolerance(ctx):
    import robustness_cases
    robustness_cases.check_decoder_tolerance(ctx)
