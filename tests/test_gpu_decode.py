"""GPU parity tests of the record-decode path (B200, `-m gpu`) through the C ABI: golden outputs of the
compiled reference, the C oracle on seeded record streams, and size-independent round-trip
properties at larger sizes."""
import base64
import zlib

import pytest

import oracle_lib as ol
import sccg_b200
from cases import cases, fasta_cases, rnd
from test_emu_decode import make_record_stream

pytestmark = pytest.mark.gpu
CASES = cases()


@pytest.fixture(scope="module")
def ctx():
    c = sccg_b200.Context(0)
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
    assert ctx.decompress(case.ref, inter) == unpack(g["reconstructed_z"])


def test_fasta_level_goldens(ctx, golden):
    for fc in fasta_cases():
        g = golden["fasta_cases"][fc.name]
        ref = ol.orc_parse_reference_fasta(fc.ref_file)
        assert ctx.decompress(ref, unpack(g["intermediate_z"])) == unpack(g["reconstructed_z"]), fc.name


@pytest.mark.parametrize("seed", range(20))
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
        ctx.reconstruct(ref, b"(400,200)", b"", b"")
    assert e.value.code == sccg_b200.SCCG_E_BOUNDS
    for enc in (b"(a,5)", b"(5,)", b"(99999999999,5)", b"AC(7,(503,497)"):
        with pytest.raises(sccg_b200.SccgError) as e:
            ctx.reconstruct(ref, enc, b"", b"")
        assert e.value.code == sccg_b200.SCCG_E_FORMAT
    assert ctx.reconstruct(ref, b"", b"", b"") == b"\n"


@pytest.mark.parametrize("n", [5_000_000, 40_000_000])
def test_roundtrip_local_synthetic(ctx, n):
    """compress -> decompress reproduces the 50-column FASTA body of the target (lossless envelope), and
    the GPU decoder agrees with the oracle decoder on the same record stream"""
    from sccg_genome_compression_b200 import synth
    ref, tgt = synth.local_pair(n, synth.seed_for(2, 11))
    ref, tgt = ref.tobytes(), tgt.tobytes()
    inter, mode = ctx.compress(ref, tgt, b">rt")
    assert mode == 0
    back = ctx.decompress(ref, inter)
    expect = b">rt\n" + b"\n".join(tgt[i:i + 50] for i in range(0, len(tgt), 50)) + b"\n"
    assert back == expect
    if n <= 5_000_000:
        rc, orc = ol.orc_decompress(ref, inter)
        assert rc == 0 and orc == back


@pytest.mark.parametrize("chunk", [4096, 65536, 1 << 20])
def test_decompress_pipelined_small_chunks(ctx, chunk, monkeypatch):
    """pipelined host path on the GPU: many small reference / output chunks, un-prepared reference chunks are poisoned"""
    from sccg_genome_compression_b200 import synth
    monkeypatch.setenv("SCCG_PIPE_CHUNK", str(chunk))
    monkeypatch.setenv("SCCG_PIPE_POISON", "1")
    for shape in ("local", "gap"):
        if shape == "local":
            ref, tgt = synth.local_pair(3_000_000, synth.seed_for(2, 31))
        else:
            ref, tgt = synth.global_gap_pair(1_500_000, 1_400_000, synth.seed_for(1, 31))
        ref, tgt = ref.tobytes(), tgt.tobytes()
        rc, inter, mode = ol.orc_compress(ref, tgt, b">pipelined")
        assert rc == 0
        rc, exp = ol.orc_decompress(ref, inter)
        assert rc == 0
        assert ctx.decompress(ref, inter) == exp
    ref = rnd(120_000, "pipe")
    inter = b">x\n\n,\n(110000,5000)(-110000,30000)ACGT(60000,100)(-50000,20000)"
    rc, exp = ol.orc_decompress(ref, inter)
    assert rc == 0 and ctx.decompress(ref, inter) == exp


@pytest.mark.parametrize("n_parts,chunk", [(2, 65536), (8, 1 << 20), (3, 4096)])
def test_decompress_parts_concatenate_to_the_whole(ctx, n_parts, chunk, monkeypatch):
    """output-range sharding on the GPU: pieces at their offsets = the whole image; un-needed reference chunks stay poisoned"""
    from sccg_genome_compression_b200 import synth
    monkeypatch.setenv("SCCG_PIPE_CHUNK", str(chunk))
    monkeypatch.setenv("SCCG_PIPE_POISON", "1")
    cases = []
    ref, tgt = synth.local_pair(4_000_000, synth.seed_for(2, 81))
    cases.append((ref.tobytes(), tgt.tobytes()))
    ref, tgt = synth.global_gap_pair(1_200_000, 1_100_000, synth.seed_for(1, 81))
    cases.append((ref.tobytes(), tgt.tobytes()))
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
