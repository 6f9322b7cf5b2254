"""Process-level drop-in check (B200, `-m gpu`): our `compress` / `decompress` executables with the
reference's argv against the golden files produced by the reference executables, and cross-decoding
(the reference's decompress consumes our record file where oracle/_ref is available)."""
import base64
import os
import shutil
import subprocess
import zlib
from pathlib import Path

import pytest

import oracle_lib as ol
from cases import fasta_cases

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "sccg-genome-compression_b200" / "bin"


def unpack(s):
    return zlib.decompress(base64.b64decode(s))


@pytest.fixture(scope="module")
def env(tmp_path_factory):
    """PATH with a stand-in `7z` (the image has none): the copy shim of oracle/7z_shim.sh"""
    d = tmp_path_factory.mktemp("shim")
    shutil.copy(ROOT / "oracle" / "7z_shim.sh", d / "7z")
    os.chmod(d / "7z", 0o755)
    e = dict(os.environ)
    e["PATH"] = str(d) + os.pathsep + e.get("PATH", "")
    import __graft_entry__ as g
    g.build_library(); g.build_cli()
    return e


@pytest.mark.parametrize("fc", fasta_cases(), ids=[c.name for c in fasta_cases()])
def test_cli_roundtrip_matches_reference_files(env, fc, golden, tmp_path):
    g = golden["fasta_cases"][fc.name]
    (tmp_path / "ref.fa").write_bytes(fc.ref_file)
    (tmp_path / "tgt.fa").write_bytes(fc.tgt_file)
    r = subprocess.run([str(BIN / "compress"), str(tmp_path / "ref.fa"), str(tmp_path / "tgt.fa"), str(tmp_path / "out")],
                       env=env, capture_output=True)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "out" / "compressed_genome.txt").read_bytes() == unpack(g["intermediate_z"])
    assert (tmp_path / "out" / "compressed_genome.txt.7z").exists()
    r = subprocess.run([str(BIN / "decompress"), str(tmp_path / "out" / "compressed_genome.txt.7z"), str(tmp_path / "ref.fa"),
                        str(tmp_path / "dec")], env=env, capture_output=True)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "dec" / "reconstructed_genome.fa").read_bytes() == unpack(g["reconstructed_z"])
    if ol.have_reference():       # the reference's own decompress.cpp consumes the GPU output unchanged
        rc, recon, err = ol.ref_decompress_cli(tmp_path / "out" / "compressed_genome.txt.7z", tmp_path / "ref.fa", tmp_path / "dec_ref")
        assert rc == 0 and recon == unpack(g["reconstructed_z"])


def test_cli_usage_and_errors(env, tmp_path):
    r = subprocess.run([str(BIN / "compress"), "a", "b"], env=env, capture_output=True)
    assert r.returncode == 1 and b"Usage:" in r.stderr
    r = subprocess.run([str(BIN / "decompress")], env=env, capture_output=True)
    assert r.returncode == 1 and b"Usage:" in r.stderr
    r = subprocess.run([str(BIN / "compress"), str(tmp_path / "missing.fa"), str(tmp_path / "missing2.fa"), str(tmp_path / "o")], env=env, capture_output=True)
    assert r.returncode == 1 and b"Error opening reference file" in r.stderr
    # bounds error of the decoder (decompression.cpp:223-229) -> exit 1
    (tmp_path / "ref.fa").write_bytes(b">r\nACGTACGTACGTACGTACGT\n")
    (tmp_path / "bad.txt.7z").write_bytes(b">h\n\n,\n(5,100)")
    r = subprocess.run([str(BIN / "decompress"), str(tmp_path / "bad.txt.7z"), str(tmp_path / "ref.fa"), str(tmp_path / "d")], env=env, capture_output=True)
    assert r.returncode == 1 and b"exceeds reference genome size" in r.stderr


def test_cli_batch_mode(env, golden, tmp_path):
    """additive `--batch <list>`: many pairs in one process (one CUDA context), same files as the one-pair invocations"""
    fcs = fasta_cases()[:4]
    clist, dlist = [], []
    for i, fc in enumerate(fcs):
        d = tmp_path / f"p{i}"; d.mkdir()
        (d / "ref.fa").write_bytes(fc.ref_file); (d / "tgt.fa").write_bytes(fc.tgt_file)
        clist.append(f"{d / 'ref.fa'} {d / 'tgt.fa'} {d / 'out'}")
        dlist.append(f"{d / 'out' / 'compressed_genome.txt.7z'} {d / 'ref.fa'} {d / 'dec'}")
    (tmp_path / "c.txt").write_text("\n".join(clist) + "\n\n")
    (tmp_path / "d.txt").write_text("\n".join(dlist) + "\n")
    r = subprocess.run([str(BIN / "compress"), "--batch", str(tmp_path / "c.txt")], env=env, capture_output=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(BIN / "decompress"), "--batch", str(tmp_path / "d.txt")], env=env, capture_output=True)
    assert r.returncode == 0, r.stderr
    for i, fc in enumerate(fcs):
        g = golden["fasta_cases"][fc.name]
        assert (tmp_path / f"p{i}" / "out" / "compressed_genome.txt").read_bytes() == unpack(g["intermediate_z"])
        assert (tmp_path / f"p{i}" / "dec" / "reconstructed_genome.fa").read_bytes() == unpack(g["reconstructed_z"])
