"""GPU parity tests of the FASTA ingest (csrc/sccg_fasta.cuh) through the C ABI: raw file images in."""
import base64
import random
import zlib

import pytest

import oracle_lib as ol
from cases import fasta_cases
from test_emu_fasta import random_fasta, unpack
from cases import rnd

pytestmark = pytest.mark.gpu
FASTA = fasta_cases()


@pytest.fixture(scope="module")
def ctx():
    import sccg_b200
    c = sccg_b200.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("fc", FASTA, ids=[c.name for c in FASTA])
def test_fasta_images_match_reference_programs(ctx, fc, golden):
    g = golden["fasta_cases"][fc.name]
    inter, mode = ctx.compress_fasta(fc.ref_file, fc.tgt_file)
    assert inter == unpack(g["intermediate_z"])
    assert ctx.decompress_fasta(fc.ref_file, inter) == unpack(g["reconstructed_z"])


@pytest.mark.parametrize("seed", range(12))
def test_fasta_ingest_random_images(ctx, seed):
    r = random.Random(repr(("fasta", seed)))
    n = r.choice([0, 1, 15, 16, 17, 1000, 5000, 9000])
    ref_seq = rnd(n, ("fr", seed), b"ACGTNacgtn")
    tgt_seq = bytearray(ref_seq)
    for _ in range(n // 300):
        tgt_seq[r.randrange(n)] = r.choice(b"ACGT")
    ref_file = random_fasta(r, ref_seq, headers=r.choice([0, 1, 3]), first_header=r.random() < 0.8)
    tgt_file = random_fasta(r, bytes(tgt_seq), headers=r.choice([0, 0, 2]), first_header=r.random() < 0.8)
    ref = ol.orc_parse_reference_fasta(ref_file)
    tgt, header = ol.orc_parse_target_fasta(tgt_file)
    rc, exp, mode = ol.orc_compress(ref, tgt, header)
    if rc != 0 or len(tgt) == 0:
        pytest.skip("degenerate image")
    got, gmode = ctx.compress_fasta(ref_file, tgt_file)
    assert (gmode, got) == (mode, exp)
    rc, back = ol.orc_decompress(ref, exp)
    if rc == 0 and len(ol.split_intermediate(exp + b"\n\n\n")[3]) > 0:
        assert ctx.decompress_fasta(ref_file, exp) == back


def test_fasta_ingest_large_roundtrip(ctx):
    """20 Mbp pair as 50-column FASTA images: compress from the images, decompress from the reference image, and the
    result must be the target file itself (the lossless envelope of SURVEY N2)"""
    from sccg_genome_compression_b200 import synth
    import numpy as np
    ref, tgt = synth.local_pair(20_000_000, synth.seed_for(2, 51))

    def image(seq, header):
        n = seq.size
        full = n // 50 * 50
        body = np.empty((full // 50, 51), dtype=np.uint8)
        body[:, :50] = seq[:full].reshape(-1, 50); body[:, 50] = 10
        tail = seq[full:].tobytes()
        return header + b"\n" + body.tobytes() + (tail + b"\n" if tail else b"")
    ref_file, tgt_file = image(ref, b">chrR synthetic"), image(tgt, b">chrT synthetic target")
    inter, mode = ctx.compress_fasta(ref_file, tgt_file)
    rc, exp, emode = ol.orc_compress(ref.tobytes(), tgt.tobytes(), b">chrT synthetic target")
    assert rc == 0 and (mode, inter) == (emode, exp)
    assert ctx.decompress_fasta(ref_file, inter) == tgt_file
