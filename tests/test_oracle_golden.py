"""The C oracle (oracle/sccg_oracle.c) against the committed outputs of the compiled reference
(tests/golden/reference_outputs.json, made by tests/golden/make_golden.py).  CPU only."""
import base64
import hashlib
import zlib

import pytest

import oracle_lib as ol
from cases import cases, fasta_cases


def unpack(s: str) -> bytes:
    return zlib.decompress(base64.b64decode(s))


CASES = cases()
FASTA = fasta_cases()


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_oracle_compress_matches_reference(case, golden):
    g = golden["cases"][case.name]
    assert hashlib.sha256(case.ref).hexdigest() == g["ref_sha256"], "case generator drifted from the goldens"
    assert hashlib.sha256(case.tgt).hexdigest() == g["tgt_sha256"], "case generator drifted from the goldens"
    rc, text, mode = ol.orc_compress(case.ref, case.tgt, case.header)
    # rc != 0 <=> the reference died in delta_encode's stoi (exit 1) and left the un-rewritten file behind
    assert (rc == 0) == (g["rc_compress"] == 0)
    assert text == unpack(g["intermediate_z"])
    assert mode == g["mode"]
    if case.expect_mode is not None:
        assert mode == case.expect_mode
    assert g["roundtrip"] == case.lossless


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_oracle_decompress_matches_reference(case, golden):
    g = golden["cases"][case.name]
    inter = unpack(g["intermediate_z"])
    if g["rc_decompress"] != 0:
        if len(ol.split_intermediate(inter + b"\n\n\n")[3]) == 0:
            pytest.skip("reference fails in decompress_genome (empty body line), before the hot path")
        rc, _ = ol.orc_decompress(case.ref, inter)
        assert rc != 0
        return
    rc, out = ol.orc_decompress(case.ref, inter)
    assert rc == 0
    assert out == unpack(g["reconstructed_z"])


@pytest.mark.parametrize("fc", FASTA, ids=[c.name for c in FASTA])
def test_oracle_fasta_level(fc, golden):
    g = golden["fasta_cases"][fc.name]
    ref = ol.orc_parse_reference_fasta(fc.ref_file)
    tgt, header = ol.orc_parse_target_fasta(fc.tgt_file)
    rc, text, _ = ol.orc_compress(ref, tgt, header)
    assert rc == 0 and text == unpack(g["intermediate_z"])
    rc, out = ol.orc_decompress(ref, text)
    assert rc == 0 and out == unpack(g["reconstructed_z"])
    assert (out == fc.tgt_file) == g["roundtrip"] == fc.lossless


def test_delta_encode_text_semantics():
    # survey N1 / experiment J: '(' inside literals poisons the chain exactly like the reference
    rc, out = ol.orc_delta_encode(b">h\n\n,\n(0,500)(7,(503,497)(993,1000)")
    assert rc == 0 and out == b">h\n\n,\n(0,500)(7,(503,497)(986,1000)"
    rc, out = ol.orc_delta_encode(b"\n,\n(10,5)AC(4,6)(100,1)")
    assert rc == 0 and out == b"\n,\n(10,5)AC(-6,6)(96,1)"
    # header line content is skipped even if it contains parentheses
    rc, out = ol.orc_delta_encode(b">chr1 (x,1)\n(5,2)\n,\n(7,7)(9,9)")
    assert rc == 0 and out == b">chr1 (x,1)\n(5,2)\n,\n(7,7)(2,9)"
    rc, out = ol.orc_delta_encode(b">h\n\n,\n(,5)")
    assert rc == 2
