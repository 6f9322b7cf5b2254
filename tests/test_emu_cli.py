"""Host logic of the two command-line programs on the CPU: compress.cpp / decompress.cpp linked against the SIMT-emulator
build of the library (tests/emu), so argv handling, file layout, the batch driver (LPT packing, two workers per "GPU",
page-locked buffers) and the streaming writer are checked in the GPU-less build container.  The product executables
(sccg-genome-compression_b200/bin, linked against the CUDA library) are checked by test_gpu_cli.py on the B200."""
import base64
import os
import shutil
import subprocess
import zlib
from pathlib import Path

import pytest

import oracle_lib as ol
from cases import fasta_cases
from emu_lib import EMU_DIR, emu_context

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "sccg-genome-compression_b200" / "host"
BIN = EMU_DIR / "bin"


def unpack(s):
    return zlib.decompress(base64.b64decode(s))


@pytest.fixture(scope="module")
def env(tmp_path_factory):
    emu_context().close()                                    # builds tests/emu/libsccg_b200_emu.so
    BIN.mkdir(exist_ok=True)
    for name in ("compress", "decompress"):
        out = BIN / name
        srcs = [HOST / f"{name}.cpp", HOST / "batch.hpp", ROOT / "include" / "sccg.h", EMU_DIR / "libsccg_b200_emu.so"]
        if not out.exists() or any(s.stat().st_mtime > out.stat().st_mtime for s in srcs):
            tmp = BIN / f"{name}.tmp.{os.getpid()}"         # built next to the target and renamed: parallel test workers never run a half-written program
            subprocess.check_call(["/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++", "-O1", "-std=c++17", "-pthread", f"-I{ROOT / 'include'}",
                                   str(HOST / f"{name}.cpp"), "-o", str(tmp), f"-L{EMU_DIR}", "-lsccg_b200_emu", f"-Wl,-rpath,{EMU_DIR}"])
            os.replace(tmp, out)
    d = tmp_path_factory.mktemp("shim")
    shutil.copy(ROOT / "oracle" / "7z_shim.sh", d / "7z")
    os.chmod(d / "7z", 0o755)
    e = dict(os.environ)
    e["PATH"] = str(d) + os.pathsep + e.get("PATH", "")
    e["SCCG_EMU_DEVICES"] = "3"
    e["SCCG_EMU_THREADS"] = "2"
    return e


def test_single_pair_goldens(env, golden, tmp_path):
    for fc in fasta_cases()[:3]:
        g = golden["fasta_cases"][fc.name]
        d = tmp_path / fc.name; d.mkdir()
        (d / "ref.fa").write_bytes(fc.ref_file); (d / "tgt.fa").write_bytes(fc.tgt_file)
        r = subprocess.run([str(BIN / "compress"), str(d / "ref.fa"), str(d / "tgt.fa"), str(d / "out")], env=env, capture_output=True)
        assert r.returncode == 0, r.stderr
        assert (d / "out" / "compressed_genome.txt").read_bytes() == unpack(g["intermediate_z"])
        r = subprocess.run([str(BIN / "decompress"), str(d / "out" / "compressed_genome.txt.7z"), str(d / "ref.fa"), str(d / "dec")], env=env, capture_output=True)
        assert r.returncode == 0, r.stderr
        assert (d / "dec" / "reconstructed_genome.fa").read_bytes() == unpack(g["reconstructed_z"])
        assert b"Time taken to decompress" in r.stdout


@pytest.mark.parametrize("gpus", [1, 3])
def test_batch_driver(env, gpus, tmp_path):
    """9 pairs of different sizes (one of them global mode, one malformed archive) over 1 / 3 devices: every output file equals
    the oracle's, a failing pair does not stop the others, exit code 1 iff a pair failed"""
    from sccg_genome_compression_b200 import synth
    clist, dlist, expect = [], [], []
    for i in range(9):
        if i == 4:
            ref, tgt = synth.global_gap_pair(40_000, 36_000, synth.seed_for(1, 70 + i))
        else:
            ref, tgt = synth.local_pair(20_000 + 7_000 * i, synth.seed_for(2, 70 + i))
        ref, tgt = ref.tobytes(), tgt.tobytes()
        header = b">pair %d" % i
        d = tmp_path / f"p{i}"; d.mkdir()
        ol.write_fasta(d / "ref.fa", ref, b">ref")
        ol.write_fasta(d / "tgt.fa", tgt, header)
        rc, inter, mode = ol.orc_compress(ref, tgt, header)
        rc2, recon = ol.orc_decompress(ref, inter)
        assert rc == 0 and rc2 == 0
        expect.append((inter, recon))
        clist.append(f"{d / 'ref.fa'} {d / 'tgt.fa'} {d / 'out'}")
        dlist.append(f"{d / 'out' / 'compressed_genome.txt.7z'} {d / 'ref.fa'} {d / 'dec'}")
    (tmp_path / "c.txt").write_text("\n".join(clist) + "\n\n")
    r = subprocess.run([str(BIN / "compress"), "--batch", str(tmp_path / "c.txt"), "--gpus", str(gpus)], env=env, capture_output=True)
    assert r.returncode == 0, r.stderr + r.stdout
    assert r.stdout.count(b": ok") == 9 and (b"on %d GPU(s)" % gpus) in r.stdout
    for i, (inter, recon) in enumerate(expect):
        assert (tmp_path / f"p{i}" / "out" / "compressed_genome.txt").read_bytes() == inter
    # a malformed archive in the middle of the list
    bad = tmp_path / "bad.txt.7z"; bad.write_bytes(b">h\n\n,\n(5,100000)")
    dlist.insert(3, f"{bad} {tmp_path / 'p0' / 'ref.fa'} {tmp_path / 'bad_dec'}")
    (tmp_path / "d.txt").write_text("\n".join(dlist) + "\n")
    r = subprocess.run([str(BIN / "decompress"), "--batch", str(tmp_path / "d.txt"), "--gpus", str(gpus)], env=env, capture_output=True)
    assert r.returncode == 1 and r.stdout.count(b": ok") == 9 and r.stdout.count(b"FAILED") == 1
    assert b"exceeds reference genome size" in r.stderr
    assert not (tmp_path / "bad_dec" / "reconstructed_genome.fa").exists()
    for i, (inter, recon) in enumerate(expect):
        assert (tmp_path / f"p{i}" / "dec" / "reconstructed_genome.fa").read_bytes() == recon
