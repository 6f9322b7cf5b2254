#!/usr/bin/env python
"""The 24-pair hg19-vs-hg18-shaped genome through the PRODUCT command-line programs: `compress --batch <list> --gpus N` and
`decompress --batch <list> --gpus N` (one process each, two workers per GPU), files on the local disk, `7z` = copy shim.
Every compressed_genome.txt is compared with the C oracle's file and every reconstructed_genome.fa with its target file;
the reference executables run the same pairs one process per pair (its own way of doing a genome) on a reduced-size subset
for the wall-clock comparison.
usage: genome_cli.py [--scale S] [--gpus N] [--ref-pairs K]      (scale 1.0 = the named configuration: 3.1 Gbp, 6.3 GB of FASTA)"""
import argparse, concurrent.futures as cf, ctypes, json, os, shutil, subprocess, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import oracle_lib as ol
import sccg_b200  # noqa: F401
from sccg_genome_compression_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.1)
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--ref-pairs", type=int, default=4, help="pairs (the smallest ones) also run through the reference executables")
ap.add_argument("--dir", default=None)
args = ap.parse_args()


def image(seq, header):
    full = seq.size // 50 * 50
    body = np.empty((full // 50, 51), dtype=np.uint8)
    body[:, :50] = seq[:full].reshape(-1, 50); body[:, 50] = 10
    tail = seq[full:].tobytes()
    return header + b"\n" + body.tobytes() + (tail + b"\n" if tail else b"")


lengths = [max(20_000, int(n * args.scale)) for n in synth.HG19_LENGTHS]
work = Path(args.dir) if args.dir else Path(tempfile.mkdtemp(prefix="sccg_genome_"))
work.mkdir(parents=True, exist_ok=True)
shim = work / "bin"; shim.mkdir(exist_ok=True); shutil.copy(ROOT / "oracle" / "7z_shim.sh", shim / "7z"); os.chmod(shim / "7z", 0o755)
env = dict(os.environ); env["PATH"] = str(shim) + os.pathsep + env["PATH"]
pairs = []
t0 = time.perf_counter()
for i, n in enumerate(lengths):
    ref, tgt = synth.local_pair(n, synth.seed_for(2, i))
    d = work / f"chr{i + 1}"; d.mkdir(exist_ok=True)
    header = b">chr%d synthetic hg19-vs-hg18 shape" % (i + 1)
    (d / "ref.fa").write_bytes(image(ref, b">chr%d reference" % (i + 1)))
    (d / "tgt.fa").write_bytes(image(tgt, header))
    pairs.append((d, ref, tgt, header))
gen_s = time.perf_counter() - t0
(work / "c.txt").write_text("".join(f"{d / 'ref.fa'} {d / 'tgt.fa'} {d / 'out'}\n" for d, *_ in pairs))
(work / "d.txt").write_text("".join(f"{d / 'out' / 'compressed_genome.txt.7z'} {d / 'ref.fa'} {d / 'dec'}\n" for d, *_ in pairs))
ours = ROOT / "sccg-genome-compression_b200" / "bin"
out = {"pairs": len(pairs), "target_bp": int(sum(lengths)), "scale": args.scale, "gpus": args.gpus, "fasta_bytes": int(sum((d / 'tgt.fa').stat().st_size + (d / 'ref.fa').stat().st_size for d, *_ in pairs)),
       "generate_and_write_s": round(gen_s, 2)}
for rep in range(2):                                          # the second run has the files in the page cache and is the one reported
    t0 = time.perf_counter()
    r = subprocess.run([str(ours / "compress"), "--batch", str(work / "c.txt"), "--gpus", str(args.gpus)], env=env, capture_output=True)
    t1 = time.perf_counter()
    assert r.returncode == 0, (r.stderr[-400:], r.stdout[-400:])
    r2 = subprocess.run([str(ours / "decompress"), "--batch", str(work / "d.txt"), "--gpus", str(args.gpus)], env=env, capture_output=True)
    t2 = time.perf_counter()
    assert r2.returncode == 0, (r2.stderr[-400:], r2.stdout[-400:])
out["ours_batch"] = {"compress_s": round(t1 - t0, 3), "decompress_s": round(t2 - t1, 3), "compress_mbp_s": round(sum(lengths) / (t1 - t0) / 1e6, 1),
                     "decompress_gbp_s": round(sum(lengths) / (t2 - t1) / 1e9, 3), "log_tail": r.stdout.decode().splitlines()[-3:], "dlog_tail": r2.stdout.decode().splitlines()[-3:]}
# ---- verification: every record file against the C oracle (threads: ctypes releases the GIL), every FASTA against its target file
lib = ol.oracle()


def check(p):
    d, ref, tgt, header = p
    o = ctypes.c_void_p(); n = ctypes.c_long(); mode = ctypes.c_int()
    rc = lib.orc_compress(ctypes.cast(ref.ctypes.data, ctypes.c_char_p), ref.size, ctypes.cast(tgt.ctypes.data, ctypes.c_char_p), tgt.size, header, len(header),
                          ctypes.byref(o), ctypes.byref(n), ctypes.byref(mode))
    exp = ctypes.string_at(o, n.value); lib.orc_free(o)
    return rc == 0 and (d / "out" / "compressed_genome.txt").read_bytes() == exp and (d / "dec" / "reconstructed_genome.fa").read_bytes() == (d / "tgt.fa").read_bytes()


with cf.ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
    out["all_files_identical_to_oracle_and_targets"] = all(ex.map(check, pairs))
# ---- the reference's way: one process per pair, on the smallest pairs
if ol.have_reference() and args.ref_pairs > 0:
    small = sorted(range(len(pairs)), key=lambda i: lengths[i])[:args.ref_pairs]
    t0 = time.perf_counter(); same = True
    for i in small:
        d = pairs[i][0]
        r = subprocess.run([str(ol.REF_DIR / "compress"), str(d / "ref.fa"), str(d / "tgt.fa"), str(d / "ref_out")], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert r.returncode == 0
    t1 = time.perf_counter()
    for i in small:
        d = pairs[i][0]
        r = subprocess.run([str(ol.REF_DIR / "decompress"), str(d / "ref_out" / "compressed_genome.txt.7z"), str(d / "ref.fa"), str(d / "ref_dec")], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert r.returncode == 0
        same = same and (d / "ref_out" / "compressed_genome.txt").read_bytes() == (d / "out" / "compressed_genome.txt").read_bytes() \
            and (d / "ref_dec" / "reconstructed_genome.fa").read_bytes() == (d / "dec" / "reconstructed_genome.fa").read_bytes()
    t2 = time.perf_counter()
    bp = sum(lengths[i] for i in small)
    out["reference_one_process_per_pair"] = {"pairs": len(small), "bp": int(bp), "compress_s": round(t1 - t0, 2), "decompress_s": round(t2 - t1, 2),
                                             "compress_mbp_s": round(bp / (t1 - t0) / 1e6, 2), "decompress_gbp_s": round(bp / (t2 - t1) / 1e9, 4), "files_identical_to_ours": same}
print(json.dumps(out))
if not args.dir:
    shutil.rmtree(work, ignore_errors=True)
