#!/usr/bin/env python
"""Process-level comparison: our `compress` / `decompress` executables against the reference's (oracle/_ref, `7z` = copy
shim) on one FASTA pair written to a temporary directory; wall clock of the whole process, outputs compared byte for byte.
usage: cli_bench.py [bp]"""
import json, os, shutil, subprocess, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import oracle_lib as ol
import sccg_b200  # noqa: F401
from sccg_genome_compression_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000
ref, tgt = synth.local_pair(n, synth.seed_for(2, 5))


def image(seq, header):
    full = seq.size // 50 * 50
    body = np.empty((full // 50, 51), dtype=np.uint8)
    body[:, :50] = seq[:full].reshape(-1, 50); body[:, 50] = 10
    tail = seq[full:].tobytes()
    return header + b"\n" + body.tobytes() + (tail + b"\n" if tail else b"")


with tempfile.TemporaryDirectory() as d:
    d = Path(d)
    (d / "ref.fa").write_bytes(image(ref, b">chrR")); (d / "tgt.fa").write_bytes(image(tgt, b">chrT synthetic"))
    shim = d / "bin"; shim.mkdir(); shutil.copy(ROOT / "oracle" / "7z_shim.sh", shim / "7z"); os.chmod(shim / "7z", 0o755)
    env = dict(os.environ); env["PATH"] = str(shim) + os.pathsep + env["PATH"]
    out = {"bp": n}
    for name, bindir in (("ours", ROOT / "sccg-genome-compression_b200" / "bin"), ("reference", ol.REF_DIR)):
        for rep in range(2 if name == "ours" else 1):                  # ours twice: the first run pays the CUDA context / page cache warm-up
            env["SCCG_TIMING"] = "1"                                   # ours: context / read / gpu / write breakdown on stderr
            t0 = time.perf_counter()
            r = subprocess.run([str(bindir / "compress"), str(d / "ref.fa"), str(d / "tgt.fa"), str(d / name)], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
            t1 = time.perf_counter()
            r2 = subprocess.run([str(bindir / "decompress"), str(d / name / "compressed_genome.txt.7z"), str(d / "ref.fa"), str(d / (name + "_dec"))], env=env,
                                stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
            t2 = time.perf_counter()
            assert r.returncode == 0 and r2.returncode == 0, (r.stderr[-300:], r2.stderr[-300:])
        out[name] = {"compress_s": round(t1 - t0, 3), "decompress_s": round(t2 - t1, 3)}
        if name == "ours":
            out[name]["compress_breakdown"] = [l for l in r.stderr.decode().splitlines() if l.startswith("timing:")]
            out[name]["decompress_breakdown"] = [l for l in r2.stderr.decode().splitlines() if l.startswith("timing:")]
    same_c = (d / "ours" / "compressed_genome.txt").read_bytes() == (d / "reference" / "compressed_genome.txt").read_bytes()
    same_d = (d / "ours_dec" / "reconstructed_genome.fa").read_bytes() == (d / "reference_dec" / "reconstructed_genome.fa").read_bytes()
    out["compressed_genome_txt_identical"] = same_c; out["reconstructed_fa_identical"] = same_d
    out["roundtrip_is_target_file"] = (d / "ours_dec" / "reconstructed_genome.fa").read_bytes() == (d / "tgt.fa").read_bytes()
    # batch mode: the same pair three times in one process (one CUDA context)
    ours = ROOT / "sccg-genome-compression_b200" / "bin"
    (d / "c.txt").write_text("".join(f"{d / 'ref.fa'} {d / 'tgt.fa'} {d / ('b%d' % i)}\n" for i in range(3)))
    (d / "d.txt").write_text("".join(f"{d / ('b%d' % i) / 'compressed_genome.txt.7z'} {d / 'ref.fa'} {d / ('bd%d' % i)}\n" for i in range(3)))
    t0 = time.perf_counter(); r = subprocess.run([str(ours / "compress"), "--batch", str(d / "c.txt")], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE); t1 = time.perf_counter()
    r2 = subprocess.run([str(ours / "decompress"), "--batch", str(d / "d.txt")], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE); t2 = time.perf_counter()
    assert r.returncode == 0 and r2.returncode == 0, (r.stderr[-300:], r2.stderr[-300:])
    out["ours_batch_of_3"] = {"compress_s": round(t1 - t0, 3), "decompress_s": round(t2 - t1, 3),
                              "identical": (d / "bd2" / "reconstructed_genome.fa").read_bytes() == (d / "tgt.fa").read_bytes()}
    print(json.dumps(out))
