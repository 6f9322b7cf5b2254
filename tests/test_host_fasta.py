"""Host-side FASTA reader of the C++ CLIs (sccg-genome-compression_b200/host/fasta_io.hpp) against the
oracle's restatement of read_genomes_from_files (compression.cpp:181-220), which is pinned to the
compiled reference by the goldens."""
import subprocess
from pathlib import Path

import pytest

import oracle_lib as ol
from cases import fasta_cases

HERE = Path(__file__).resolve().parent


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    exe = tmp_path_factory.mktemp("probe") / "fasta_probe"
    subprocess.check_call(["/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++", "-O1", "-std=c++17",
                           str(HERE / "host_fasta_probe.cpp"), "-o", str(exe)])
    return exe


@pytest.mark.parametrize("fc", fasta_cases(), ids=[c.name for c in fasta_cases()])
def test_host_reader_matches_reference_semantics(probe, fc, tmp_path):
    (tmp_path / "r.fa").write_bytes(fc.ref_file)
    (tmp_path / "t.fa").write_bytes(fc.tgt_file)
    out = subprocess.run([str(probe), "reference", str(tmp_path / "r.fa")], capture_output=True, check=True).stdout
    assert out == b"\n" + ol.orc_parse_reference_fasta(fc.ref_file)
    out = subprocess.run([str(probe), "target", str(tmp_path / "t.fa")], capture_output=True, check=True).stdout
    seq, header = ol.orc_parse_target_fasta(fc.tgt_file)
    assert out == header + b"\n" + seq
