"""GPU parity at the FULL sizes of the BASELINE.json configurations (B200, `-m gpu`), through the C ABI:

  cfg 1  chr19-shaped gap pair      63,811,651 / 59,128,983 symbols   -> local attempt aborts, global mode
  cfg 2  chr1-sized local pair      249,250,621 symbols                -> local segment matching
  cfg 3  chr21-shaped divergent     48,129,895 symbols                 -> global mode, lookup heavy

Each compares bytes and mode of sccg_compress with the C oracle (`ol.orc_compress`, compression.cpp:320-582) and the
decompressed image with `ol.orc_decompress` (decompression.cpp:117-279), and checks that the image is the target's FASTA
text.  The oracle needs roughly 10 s + 10 s per configuration on one host core."""
import numpy as np
import pytest

import oracle_lib as ol
import sccg_b200
from sccg_genome_compression_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = sccg_b200.Context(0)
    yield c
    c.close()


def _fasta_image(header: bytes, tgt: np.ndarray) -> bytes:
    n = tgt.size
    full = n // 50 * 50
    body = np.empty((full // 50, 51), dtype=np.uint8)
    body[:, :50] = tgt[:full].reshape(-1, 50)
    body[:, 50] = 10
    tail = tgt[full:].tobytes()
    text = body.tobytes()
    if tail:
        text += tail + b"\n"
    return header + b"\n" + text


def _check_full(ctx, ref: np.ndarray, tgt: np.ndarray, header: bytes, want_mode: int):
    rb, tb = ref.tobytes(), tgt.tobytes()
    rc, exp, emode = ol.orc_compress(rb, tb, header)
    assert rc == 0 and emode == want_mode
    got, gmode = ctx.compress(ref, tgt, header)
    assert gmode == emode
    assert len(got) == len(exp)
    assert got == exp
    # decode: GPU image == oracle image == the target's own FASTA text (lossless envelope)
    back = ctx.decompress(rb, got)
    rc, oback = ol.orc_decompress(rb, exp)
    assert rc == 0 and back == oback
    assert back == _fasta_image(header, tgt)
    return len(got)


def test_cfg2_chr1_local_full_size(ctx):
    ref, tgt = synth.local_pair(synth.CHR1_LEN, synth.seed_for(2, 0))
    assert ref.size == 249_250_621
    _check_full(ctx, ref, tgt, b">chr1 synthetic hg19-vs-hg18 shape", 0)


def test_cfg1_chr19_gap_full_size(ctx):
    ref, tgt = synth.global_gap_pair(63_811_651, 59_128_983, synth.seed_for(1, 0))
    _check_full(ctx, ref, tgt, b">chr19 synthetic gap pair", 1)


def test_cfg3_chr21_divergent_full_size(ctx):
    ref, tgt = synth.divergent_pair(48_129_895, synth.seed_for(3, 0))
    _check_full(ctx, ref, tgt, b">chr21 synthetic divergent pair", 1)
