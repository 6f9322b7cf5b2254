"""Kernel-logic tests of the compression path on the CPU: the CUDA sources compiled against the SIMT
emulator (tests/emu) versus the C oracle and the golden outputs of the compiled reference.  These
do not replace the `-m gpu` parity tests (tests/test_gpu_*.py); they keep the kernels honest in the
GPU-less build container."""
import base64
import os
import random
import zlib

import pytest

import oracle_lib as ol
from cases import cases, rnd
from emu_lib import emu_context

CASES = cases()


@pytest.fixture(scope="module")
def ctx():
    c = emu_context()
    yield c
    c.close()


@pytest.fixture(params=[0, 1, 2], ids=["asc", "desc", "mixed"])
def sweep_order(request):
    os.environ["SCCG_EMU_ORDER"] = str(request.param)
    yield request.param
    os.environ.pop("SCCG_EMU_ORDER", None)


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_compress_matches_golden(ctx, case, golden):
    g = golden["cases"][case.name]
    expect = zlib.decompress(base64.b64decode(g["intermediate_z"]))
    if g["rc_compress"] != 0:
        # the reference died in delta_encode's stoi (literal '(' in the target): same failure, same file left behind
        import sccg_b200
        with pytest.raises(sccg_b200.SccgError) as ei:
            ctx.compress(case.ref, case.tgt, case.header)
        assert ei.value.code == sccg_b200.SCCG_E_STOI and ei.value.partial == expect
        return
    got, mode = ctx.compress(case.ref, case.tgt, case.header)
    assert mode == g["mode"]
    assert got == expect


def _mutated_pair(seed, n, alphabet=b"ACGT", snp=0.01, indel=0.0):
    r = random.Random(repr(seed))
    ref = bytes(r.choice(alphabet) for _ in range(n))
    t = bytearray()
    for c in ref:
        x = r.random()
        if x < snp:
            t.append(r.choice(alphabet))
        elif x < snp + indel / 2:
            continue
        elif x < snp + indel:
            t.append(c); t.append(r.choice(alphabet))
        else:
            t.append(c)
    return ref, bytes(t)


@pytest.mark.parametrize("seed", range(16))
def test_match_sequences_local_vs_oracle(ctx, seed, sweep_order):
    alphabet = [b"ACGT", b"AC", b"ACGTN", b"A", b"ACGTacgtn"][seed % 5]
    n = random.Random(seed).randint(1, 1000)
    ref, tgt = _mutated_pair(("loc", seed), n, alphabet, snp=0.03, indel=0.004)
    ref, tgt = ref.upper()[:1000], tgt.upper()[:1000]
    for k in (14, 10):
        exp = [(r.p, r.l, r.lit) for r in ol.orc_match_sequences(ref, tgt, k, 0, False, 5000)]
        got = [(r.p, r.l, r.lit) for r in ctx.match_sequences(ref, tgt, k, 0, False, 5000)]
        assert got == exp


@pytest.mark.parametrize("seed", range(6))
def test_compress_local_random(ctx, seed):
    r = random.Random(repr(("cl", seed)))
    n = r.randint(2000, 30000)
    ref, tgt = _mutated_pair(("cl", seed), n, b"ACGT", snp=[0.001, 0.01, 0.05][seed % 3])
    t = bytearray(tgt)
    for _ in range(10):
        a = r.randrange(len(t)); b = min(len(t), a + r.choice([1, 2, 30, 700, 2500]))
        t[a:b] = bytes(t[a:b]).lower()
        a = r.randrange(len(t)); b = min(len(t), a + r.choice([1, 5, 120, 1500]))
        t[a:b] = b"N" * (b - a)
        if seed % 2:
            ref = ref[:a] + b"N" * (b - a) + ref[b:]
    rc, exp, mode = ol.orc_compress(ref, bytes(t), b">rnd")
    assert rc == 0
    if mode != 0:
        pytest.skip("went global")
    got, gmode = ctx.compress(ref, bytes(t), b">rnd")
    assert gmode == 0 and got == exp


GRAMMAR_PIECES = [b"(", b")", b",", b"(5,", b"(12,3)", b"()", b"(-7,2)", b"(+3,1)", b"(99999999999,1)", b"(2147483647,9)", b"((", b"))", b"(,)",
                  b"(3", b"7)", b"(x,1)", b"0123456789", b"(1,2,3)", b"(-,", b"(2147483648,1)", b"(-2147483648,1)"]


SAFE_PIECES = [b")", b",", b"(5,", b"(12,3)", b"()", b"(-7,2)", b"(+3,1)", b"0123456789", b"(1,2,3)", b"(2147483647,9)", b"(-2147483648,1)", b"))", b"7)"]


def grammar_pair(seed, n=9000, pieces=6, make_global=False):
    """a near-identical pair whose target carries literal grammar symbols of the record stream (experiment J)"""
    r = random.Random(repr(("gram", seed)))
    ref = rnd(n, ("gramref", seed))
    t = bytearray(mutate_some(ref, r, n // 700))
    if make_global:
        t[n // 4:n // 4] = rnd(5200, ("gramins", seed))            # a 5 kb insertion: T2 abort -> global parse
    for _ in range(pieces):
        piece = r.choice(GRAMMAR_PIECES if seed % 2 else SAFE_PIECES)     # even seeds: stoi never throws
        at = r.randrange(0, len(t) - len(piece))
        t[at:at + len(piece)] = piece
    if seed % 3 == 0:
        t += r.choice(SAFE_PIECES) + b"ACGT"                      # leftover literals past the last reference segment
    return ref, bytes(t)


def mutate_some(seq, r, count):
    b = bytearray(seq)
    for p in r.sample(range(len(b)), count):
        b[p] = r.choice(b"ACGT")
    return bytes(b)


def check_compress_like_oracle(ctx, ref, tgt, header=b">gram (alt) 1,2"):
    import sccg_b200
    rc, exp, mode = ol.orc_compress(ref, tgt, header)
    if rc != 0:                                                     # the reference throws from stoi: file left un-rewritten, exit 1
        with pytest.raises(sccg_b200.SccgError) as ei:
            ctx.compress(ref, tgt, header)
        assert ei.value.code == sccg_b200.SCCG_E_STOI and ei.value.partial == exp
        return "stoi"
    got, gmode = ctx.compress(ref, tgt, header)
    assert (gmode, got) == (mode, exp)
    return "ok"


@pytest.mark.parametrize("seed", range(24))
def test_text_level_delta_fuzz(ctx, seed):
    """literal '(' ')' ',' and digits in the target: the text-level delta pass must reproduce delta_encode (:262-292) byte for byte"""
    ref, tgt = grammar_pair(seed, make_global=(seed % 4 >= 2))
    check_compress_like_oracle(ctx, ref, tgt)


def test_text_level_delta_long_literal_runs(ctx):
    # '(' far away from the next ')' (wide searches / copies of the warp), target much longer than the reference
    ref = rnd(3000, "tl")
    tgt = ref[:1500] + b"(" + ref[1501:] + rnd(70000, "tl2") + b")" + rnd(5000, "tl3") + b"(12,3)" + rnd(300, "tl4")
    assert check_compress_like_oracle(ctx, ref, tgt) in ("ok", "stoi")


def diag_fuzz_pair(r):
    """near-identical segments (the diagonal-hypothesis path of seg_match_k) with planted traps: repeats inside the
    reference segment, mutated windows that occur elsewhere, SNPs next to segment borders, clustered SNPs"""
    alphabet = r.choice([b"ACGT", b"ACGT", b"AC", b"ACG", b"ACGTN"])
    n = r.choice([1000, 2000, 3000, 2500, 1014, 1999, 14, 15, 40])
    ref = bytearray(r.choice(alphabet) for _ in range(n))
    for _ in range(r.randint(0, 3)):                                   # self-repeats of 14..60 symbols inside one segment
        if n < 200:
            break
        seg = r.randrange(0, (n + 999) // 1000) * 1000
        hi = min(n, seg + 1000)
        ln = r.randint(14, 60)
        if hi - seg < 2 * ln + 2:
            continue
        a = r.randrange(seg, hi - ln); b = r.randrange(seg, hi - ln)
        ref[b:b + ln] = ref[a:a + ln]
    tgt = bytearray(ref)
    for _ in range(r.choice([0, 1, 1, 2, 3, 5, 12, 30])):              # substitutions, some clustered, some at borders
        p = r.choice([r.randrange(n), r.randrange(n), min(n - 1, r.choice([0, 1, 13, 14, 985, 986, 987, 999, 1000, 1001, 1013]))])
        tgt[p] = r.choice(alphabet)
        if r.random() < 0.3 and p + 3 < n:
            tgt[p + r.randint(1, 3)] = r.choice(alphabet)
    if r.random() < 0.3 and n >= 200:                                  # a mutated window that occurs elsewhere in the reference segment
        seg = r.randrange(0, (n + 999) // 1000) * 1000
        hi = min(n, seg + 1000)
        if hi - seg > 120:
            a = r.randrange(seg, hi - 50); b = r.randrange(seg, hi - 50)
            tgt[b:b + 20] = ref[a:a + 20]
    if r.random() < 0.2:
        tgt = tgt[:r.randrange(max(1, n - 30), n + 1)]                  # shorter last target segment (Lr != Lt)
    return bytes(ref), bytes(tgt)


@pytest.mark.parametrize("seed", range(10))
def test_diag_hypothesis_fuzz(ctx, seed):
    r = random.Random(repr(("dv", seed)))
    for it in range(60):
        ref, tgt = diag_fuzz_pair(r)
        rc, exp, mode = ol.orc_compress(ref, tgt, b">dv")
        assert rc == 0
        got, gmode = ctx.compress(ref, tgt, b">dv")
        assert (gmode, got) == (mode, exp), (seed, it, ref, tgt)


def crowded_bucket_pair(r: random.Random) -> tuple[bytes, bytes]:
    """one segment pair whose k-mer index has a crowded bucket: a long run of one symbol (the border of an N block), a
    short-period repeat or several copies of a stretch, with substitutions / small indels in and around it"""
    n = r.choice([1000, 1000, 1000, r.randint(200, 999)])
    ref = bytearray(r.choice(b"ACGT") for _ in range(n))
    kind = r.randrange(4)
    if kind == 0:                                   # run at the start / end / in the middle
        ln = r.randint(20, min(700, n - 20)); at = r.choice([0, n - ln, r.randint(0, n - ln)])
        ref[at:at + ln] = bytes([r.choice(b"NNNA")]) * ln
    elif kind == 1:                                 # short-period repeat
        unit = bytes(r.choice(b"ACGT") for _ in range(r.randint(2, 5)))
        ln = r.randint(40, min(600, n - 20)); at = r.randint(0, n - ln)
        ref[at:at + ln] = (unit * (ln // len(unit) + 1))[:ln]
    elif kind == 2:                                 # two runs of the same symbol
        for _ in range(2):
            ln = r.randint(20, 200); at = r.randint(0, n - ln)
            ref[at:at + ln] = b"N" * ln
    else:                                           # many copies of one 40-mer
        w = bytes(r.choice(b"ACGT") for _ in range(40))
        for _ in range(r.randint(3, 12)):
            at = r.randint(0, n - 40); ref[at:at + 40] = w
    tgt = bytearray(ref)
    for _ in range(r.randint(0, 6)):
        x = r.randrange(len(tgt)); tgt[x] = r.choice(b"ACGTN")
    for _ in range(r.randint(0, 2)):                # runs of one symbol grow / shrink in the target (closed-form candidate fold of lm_fold_runs)
        x = r.randrange(len(tgt)); d = r.randint(1, 30)
        if r.random() < 0.5:
            tgt[x:x] = bytes([tgt[x]]) * d
        else:
            del tgt[x:x + d]
    if r.random() < 0.15:
        tgt = bytearray(b"N" * r.randint(10, 60)) + tgt
    if r.random() < 0.3:
        x = r.randrange(len(tgt)); d = r.randint(1, 6)
        tgt[x:x] = bytes(r.choice(b"ACGTN") for _ in range(d)); del tgt[-d:]
    if r.random() < 0.2:
        tgt = tgt[r.randint(1, 30):] + bytearray(r.choice(b"ACGT") for _ in range(10))
    return bytes(ref), bytes(tgt[:1000])


def runs_of_one_symbol_pair(r: random.Random) -> tuple[bytes, bytes]:
    """segment pairs around long runs of one symbol (N blocks, poly-A): runs at the start / end / inside of the reference, runs
    that grow or shrink in the target, targets that begin inside a run, several runs of the same symbol -- the closed-form
    candidate fold of lm_fold_runs next to the chains (runs shorter than 15 symbols stay there)"""
    n = r.choice([1000, 1000, r.randint(30, 999)])
    alpha = r.choice([b"ACGT", b"ACGT", b"AN", b"ACGTN"])
    ref = bytearray(r.choice(alpha) for _ in range(n))
    for _ in range(r.randint(1, 4)):
        ln = r.randint(8, min(400, n)); at = r.choice([0, n - ln, r.randint(0, n - ln)])
        ref[at:at + ln] = bytes([r.choice(b"NNA")]) * ln
    tgt = bytearray(ref)
    for _ in range(r.randint(0, 5)):
        x = r.randrange(len(tgt)); tgt[x] = r.choice(b"ACGTN")
    for _ in range(r.randint(0, 3)):
        x = r.randrange(len(tgt)); d = r.randint(1, 30)
        if r.random() < 0.5:
            tgt[x:x] = bytes([tgt[x]]) * d
        else:
            del tgt[x:x + d]
    if r.random() < 0.3:
        tgt = tgt[r.randint(0, 50):]
    if r.random() < 0.2:
        tgt = bytearray(b"N" * r.randint(10, 60)) + tgt
    return bytes(ref), bytes(tgt[:1000])


@pytest.mark.parametrize("seed", range(8))
def test_runs_of_one_symbol_vs_oracle(ctx, seed):
    r = random.Random(repr(("runs", seed)))
    for it in range(150):
        ref, tgt = runs_of_one_symbol_pair(r)
        for k in (14, 10):
            exp = [(x.p, x.l, x.lit) for x in ol.orc_match_sequences(ref, tgt, k, 0, False, 0)]
            got = [(x.p, x.l, x.lit) for x in ctx.match_sequences(ref, tgt, k, 0, False, 0)]
            assert got == exp, (seed, it, k, ref, tgt)


@pytest.mark.parametrize("seed", range(12))
def test_crowded_buckets_vs_oracle(ctx, seed):
    """low-complexity segments: dozens to hundreds of candidates per looked-up k-mer, ties between them (the tie-break of
    compression.cpp:114-130 decides), the diagonal candidate folded ahead of its turn and the rest pruned by length"""
    r = random.Random(repr(("crowd", seed)))
    for it in range(40):
        ref, tgt = crowded_bucket_pair(r)
        for k in (14, 10):
            exp = [(x.p, x.l, x.lit) for x in ol.orc_match_sequences(ref, tgt, k, 0, False, 0)]
            got = [(x.p, x.l, x.lit) for x in ctx.match_sequences(ref, tgt, k, 0, False, 0)]
            assert got == exp, (seed, it, k, ref, tgt)


@pytest.mark.parametrize("chunk", [4096, 50000])
def test_compress_chunked_upload(ctx, chunk, monkeypatch):
    """host entry point: the reference arrives in chunks and the segment matcher is launched once per chunk"""
    from sccg_genome_compression_b200 import synth
    monkeypatch.setenv("SCCG_PIPE_CHUNK", str(chunk))
    for n, cut in ((123_456, 0), (90_000, 17_000), (64_000, -9_000)):
        ref, tgt = synth.local_pair(n, synth.seed_for(2, 41))
        ref, tgt = ref.tobytes(), tgt.tobytes()
        if cut > 0:
            tgt = tgt[:-cut]                        # reference longer than the target
        elif cut < 0:
            ref = ref[:cut]                         # target longer than the reference: leftover literals
        rc, exp, mode = ol.orc_compress(ref, tgt, b">chunks")
        assert rc == 0
        got, gmode = ctx.compress(ref, tgt, b">chunks")
        assert (gmode, got) == (mode, exp)


def test_many_segments_scan_paths(ctx):
    # > 2048 segments so that the device-wide scan takes its multi-tile path
    ref = rnd(3_000_000 // 4, "big") * 4
    tgt = bytearray(ref)
    r = random.Random("bigm")
    for p in r.sample(range(len(tgt)), 3000):
        tgt[p] = r.choice(b"ACGT")
    tgt = bytes(tgt[:2_999_500]).replace(b"ACGTAC", b"acgtac", 50)
    rc, exp, mode = ol.orc_compress(ref, tgt, b">big")
    got, gmode = ctx.compress(ref, tgt, b">big")
    assert (gmode, got) == (mode, exp)


@pytest.mark.parametrize("seed", range(12))
def test_match_sequences_global_vs_oracle(ctx, seed, sweep_order):
    """banded global parse incl. lost state, chance re-syncs, repeats (2-letter alphabet) and the pn2 == 0 quirk region"""
    alphabet = [b"ACGT", b"ACGT", b"AC", b"ACGT"][seed % 4]
    n = 12000 if alphabet == b"ACGT" else 2500
    ref, tgt = _mutated_pair(("glob", seed), n, alphabet, snp=0.01, indel=0.002)
    r = random.Random(seed)
    if seed % 2:
        cut = sorted(r.sample(range(len(tgt)), 4))
        parts = [tgt[:cut[0]], tgt[cut[0]:cut[1]], tgt[cut[1]:cut[2]], tgt[cut[2]:cut[3]], tgt[cut[3]:]]
        r.shuffle(parts)
        tgt = b"".join(parts)
    if seed % 3 == 0:
        tgt = rnd(r.randint(1, 700), ("pre", seed)) + tgt        # unrelated prefix: long first-match search
    exp = [(x.p, x.l, x.lit) for x in ol.orc_match_sequences(ref, tgt, 14, 100, True, 0)]
    got = [(x.p, x.l, x.lit) for x in ctx.match_sequences(ref, tgt, 14, 100, True, 0)]
    assert got == exp


@pytest.mark.parametrize("k,m", [(8, 0), (10, 5), (16, 120), (12, 37)])
def test_match_sequences_global_other_parameters(ctx, k, m):
    ref, tgt = _mutated_pair(("gp", k, m), 6000, b"ACG", snp=0.02, indel=0.003)
    exp = [(x.p, x.l, x.lit) for x in ol.orc_match_sequences(ref, tgt, k, m, True, 11)]
    got = [(x.p, x.l, x.lit) for x in ctx.match_sequences(ref, tgt, k, m, True, 11)]
    assert got == exp


def test_global_degenerate_inputs(ctx):
    ref = rnd(3000, "gd")
    for tgt in (b"", ref[:5], ref[:14], rnd(500, "unrelated"), ref, b"A" * 400):
        for rr in (ref, b"", ref[:10], b"A" * 300):
            exp = [(x.p, x.l, x.lit) for x in ol.orc_match_sequences(rr, tgt, 14, 100, True, 0)]
            got = [(x.p, x.l, x.lit) for x in ctx.match_sequences(rr, tgt, 14, 100, True, 0)]
            assert got == exp, (len(rr), len(tgt))


@pytest.mark.parametrize("seed", range(4))
def test_compress_global_random(ctx, seed):
    from sccg_genome_compression_b200 import synth
    if seed < 2:
        ref, tgt = synth.global_gap_pair(150_000 + 7 * seed, 140_000, synth.seed_for(1, seed))
    else:
        ref, tgt = synth.divergent_pair(120_000, synth.seed_for(3, seed))
    ref, tgt = ref.tobytes(), tgt.tobytes()
    rc, exp, mode = ol.orc_compress(ref, tgt, b">g")
    got, gmode = ctx.compress(ref, tgt, b">g")
    assert mode == 1 and (gmode, got) == (mode, exp)
    assert ctx.decompress(ref, got) == ol.orc_decompress(ref, exp)[1]


@pytest.mark.parametrize("chunk", [64, 300, 1024])
@pytest.mark.parametrize("seed", range(8))
def test_global_speculative_parse_small_chunks(ctx, seed, chunk):
    """tiny speculation chunks: splices, failed guesses, lost re-speculation rounds and long matches skipping chunks"""
    os.environ["SCCG_GP_CHUNK"] = str(chunk)
    try:
        alphabet = [b"ACGT", b"ACGT", b"AC", b"ACGT"][seed % 4]
        n = 9000 if alphabet == b"ACGT" else 2500
        ref, tgt = _mutated_pair(("spec", seed), n, alphabet, snp=[0.002, 0.02][seed % 2], indel=0.002)
        r = random.Random(seed)
        if seed % 2:
            cut = sorted(r.sample(range(len(tgt)), 4))
            parts = [tgt[:cut[0]], tgt[cut[0]:cut[1]], tgt[cut[1]:cut[2]], tgt[cut[2]:cut[3]], tgt[cut[3]:]]
            r.shuffle(parts)
            tgt = b"".join(parts)
        if seed % 3 == 0:
            tgt = tgt[:len(tgt) // 2] + rnd(3000, ("junk", seed)) + tgt[len(tgt) // 2:]     # long unmatched stretch: lost state
        exp = [(x.p, x.l, x.lit) for x in ol.orc_match_sequences(ref, tgt, 14, 100, True, 0)]
        got = [(x.p, x.l, x.lit) for x in ctx.match_sequences(ref, tgt, 14, 100, True, 0)]
        assert got == exp
        prof = ctx.profile()
        assert prof["spec_rounds"] >= 1
    finally:
        os.environ.pop("SCCG_GP_CHUNK", None)


def lost_scan_pair(seed, junk=40_000):
    """a pair whose global parse gets LOST (a long unrelated stretch) and is picked up again by planted hits: k-mers of the
    window around the last match end that reappear far ahead -- in range (resumes the parse there) or followed by more of
    the same diagonal (a real continuation)"""
    r = random.Random(repr(("lost", seed)))
    ref = rnd(30_000, ("lostref", seed))
    a_len = r.randint(2_000, 6_000)
    tgt = bytearray(ref[:a_len])                                # parse ends this block with e = a_len - 1
    j = bytearray(rnd(junk, ("lostjunk", seed)))
    for _ in range(r.randint(1, 4)):                            # chance-like hits: a k-mer from the window of e, somewhere in the junk
        w = a_len - 1 + r.randint(-100, 100)
        at = r.randrange(0, junk - 200)
        ln = r.choice([14, 14, 15, 40])
        j[at:at + ln] = ref[w:w + ln]
    tgt += j
    if seed % 2:
        tgt += ref[a_len + r.randint(-50, 50):a_len + 9_000]    # the true continuation, in range of e after all that junk
    tgt += rnd(r.randint(0, 3000), ("losttail", seed))
    return ref, bytes(tgt)


@pytest.mark.parametrize("grid", [1, 3, 0])
@pytest.mark.parametrize("seed", range(6))
def test_global_lost_scan(ctx, seed, grid, monkeypatch):
    """the front gets lost, gp_lost_scan_k finds the next position with a candidate in the window of e (tiles interleaved
    over 1 / 3 / the default number of CTAs), the front resumes there: identical to the sequential parse"""
    monkeypatch.setenv("SCCG_GP_CHUNK", "2048")
    if grid:
        monkeypatch.setenv("SCCG_GP_SCAN_GRID", str(grid))
    ref, tgt = lost_scan_pair(seed)
    exp = [(x.p, x.l, x.lit) for x in ol.orc_match_sequences(ref, tgt, 14, 100, True, 0)]
    got = [(x.p, x.l, x.lit) for x in ctx.match_sequences(ref, tgt, 14, 100, True, 0)]
    assert got == exp
    assert ctx.profile()["spec_rounds"] >= 2                    # at least one lost scan happened


def _compress_device_emu(ctx, ref: bytes, tgt: bytes, header: bytes):
    """device-resident entry point under the emulator (device memory is host memory there)"""
    import numpy as np
    r = np.frombuffer(ref + bytes(64), dtype=np.uint8).copy(); t = np.frombuffer(tgt + bytes(64), dtype=np.uint8).copy()
    ptr, n, mode = ctx.compress_device(r.ctypes.data, len(ref), t.ctypes.data, len(tgt), header)
    return ctx.download(ptr, n), mode


@pytest.mark.parametrize("shape", ["shifted_tail", "leftover_only", "reference_insertion", "small_shift", "tail_of_n"])
def test_compress_device_abort_probe(ctx, shape):
    """sequences of different length: a probe launch over the last segments may raise the T2 abort before the main launch
    (device-resident path only).  With and without an abort the file must be the oracle's."""
    n = 200_000
    ref = rnd(n, "probe")
    r = random.Random("probe" + shape)
    t = bytearray(ref)
    for p in r.sample(range(n), 150):
        t[p] = r.choice(b"ACGT")
    if shape == "shifted_tail":
        tgt = bytes(t[:60_000]) + bytes(t[63_000:])                    # deletion: everything after it is 3 segments off
    elif shape == "leftover_only":
        tgt = bytes(t) + rnd(5_000, "extra")                           # no shift, leftover target segments
    elif shape == "reference_insertion":
        tgt = bytes(t); ref = ref[:20_000] + rnd(2_500, "ins") + ref[20_000:]
    elif shape == "small_shift":
        tgt = bytes(t[:90_000]) + b"ACGTTGCAAC" * 3 + bytes(t[90_000:]) + rnd(1_200, "tail")   # 30 symbols off: still matches inside the segments
    else:
        tgt = bytes(t[:150_000]); ref = ref[:150_000] + b"N" * 50_000  # the probe range is all N on the reference side
    rc, exp, mode = ol.orc_compress(ref, tgt, b">probe")
    assert rc == 0
    got, gmode = _compress_device_emu(ctx, ref, tgt, b">probe")
    assert (gmode, got) == (mode, exp)
    assert mode == (1 if shape in ("shifted_tail", "reference_insertion") else 0)


@pytest.mark.parametrize("two_phase", [1, 0])
@pytest.mark.parametrize("seed", range(6))
def test_compress_device_two_phase_matcher(ctx, seed, two_phase, monkeypatch):
    """device-resident pairs: the bulk launch of the matcher queues the segments that need the generic path, a second launch
    works the queue off (pairs of 64 K segments and more; forced here for small ones, and off) -- same file either way, the oracle's"""
    monkeypatch.setenv("SCCG_LM_TWO_PHASE_MIN", "0" if two_phase else "2000000000")
    r = random.Random(repr(("2ph", seed)))
    if seed % 3 == 0:
        ref, tgt = _mutated_pair(("2ph", seed), 40_000, b"ACGT", snp=0.004, indel=0.0)          # equal lengths: identical / diagonal / generic mix
        t = bytearray(tgt)
        for _ in range(12):                                                                       # compensated indels: generic segments
            x = r.randrange(1000, len(t) - 1000); d = r.randint(1, 8)
            t[x:x] = bytes(r.choice(b"ACGT") for _ in range(d)); del t[x + 150:x + 150 + d]
        tgt = bytes(t)
    elif seed % 3 == 1:
        pairs = [crowded_bucket_pair(r) for _ in range(30)]
        pairs = [(a, b) for a, b in pairs if len(a) == 1000 and len(b) == 1000]
        ref = b"".join(a for a, _ in pairs) + rnd(700, ("2ph-tail", seed)); tgt = b"".join(b for _, b in pairs) + rnd(350, ("2ph-tail2", seed))
    else:
        ref, tgt = _mutated_pair(("2ph", seed), 25_000, b"ACGTacgtN", snp=0.01, indel=0.0005)    # drifting diagonals, some failing segments
    rc, exp, mode = ol.orc_compress(ref, tgt, b">two phase")
    assert rc == 0
    got, gmode = _compress_device_emu(ctx, ref, tgt, b">two phase")
    assert (gmode, got) == (mode, exp)


@pytest.mark.parametrize("bits", [4, 16, 24])
@pytest.mark.parametrize("seed", [1, 2, 3, 6])
def test_global_index_bucket_table_widths(ctx, seed, bits, monkeypatch):
    """the offset table over the sorted k-mer index at forced widths: 24 bits = one bucket per hash value (no search through
    the keys, the width large references get), 4 bits = long binary searches inside a bucket"""
    monkeypatch.setenv("SCCG_GP_BUCKET_BITS", str(bits))
    monkeypatch.setenv("SCCG_GP_CHUNK", "300")
    alphabet = [b"ACGT", b"AC"][seed % 2]                      # AC: many equal k-mers, long candidate runs
    n = 6000 if alphabet == b"ACGT" else 2500
    ref, tgt = _mutated_pair(("bkt", seed), n, alphabet, snp=0.01, indel=0.002)
    r = random.Random(seed)
    cut = sorted(r.sample(range(len(tgt)), 2))
    tgt = tgt[cut[1]:] + tgt[cut[0]:cut[1]] + tgt[:cut[0]]     # rearranged: unrestricted lookups after every block boundary
    exp = [(x.p, x.l, x.lit) for x in ol.orc_match_sequences(ref, tgt, 14, 100, True, 0)]
    got = [(x.p, x.l, x.lit) for x in ctx.match_sequences(ref, tgt, 14, 100, True, 0)]
    assert got == exp


@pytest.mark.parametrize("stride", [2, 8, 16])
@pytest.mark.parametrize("seed", range(8))
def test_global_sampled_index(ctx, seed, stride, monkeypatch):
    """sampled reference index (every stride-th position; what references of a million k-mers and more get): the diagonal
    guesses come from the sampled index, the first step from the brute-force occurrence list -- same records as the
    sequential parse, and the sampled mode is the one that produced them"""
    monkeypatch.setenv("SCCG_GP_STRIDE", str(stride))
    monkeypatch.setenv("SCCG_GP_CHUNK", ["300", "1024"][seed % 2])
    alphabet = [b"ACGT", b"ACGT", b"ACGT", b"AC"][seed % 4]
    n = 9000 if alphabet == b"ACGT" else 2500
    ref, tgt = _mutated_pair(("samp", seed), n, alphabet, snp=[0.002, 0.02][seed % 2], indel=0.002)
    r = random.Random(seed)
    if seed % 3 == 1:
        cut = sorted(r.sample(range(100, len(tgt)), 2))
        tgt = tgt[:cut[0]] + tgt[cut[1]:] + tgt[cut[0]:cut[1]]                   # rearranged behind a common beginning
    if seed % 3 == 2:
        tgt = tgt[:len(tgt) // 2] + rnd(3000, ("sjunk", seed)) + tgt[len(tgt) // 2:]
    tgt = ref[300:340] + tgt                                                    # the first k-mer occurs in the reference
    exp = [(x.p, x.l, x.lit) for x in ol.orc_match_sequences(ref, tgt, 14, 100, True, 0)]
    got = [(x.p, x.l, x.lit) for x in ctx.match_sequences(ref, tgt, 14, 100, True, 0)]
    assert got == exp
    if alphabet == b"ACGT":
        assert ctx.profile()["index_stride"] == stride


SAMPLED_FALLBACK_SHAPES = ["first_kmer_absent", "first_kmers_repeat", "long_unmatched_start", "first_kmer_everywhere", "p0_fallthrough", "first_kmer_twice"]


def sampled_fallback_pair(shape):
    """pairs around the lookups the sampled index mode cannot serve; returns (ref, tgt, stride that must have produced the result
    when 8 was asked for)"""
    ref = rnd(12_000, "fbref")
    body = bytearray(ref[200:9000])
    for p in random.Random("fb").sample(range(len(body)), 40):
        body[p] = ord("A")
    if shape == "first_kmer_absent":
        return ref, b"T" * 30 + bytes(body), 8                                  # found by the scan of the first 256 target positions
    if shape == "first_kmers_repeat":
        # equal k-mers at several of the first positions (the scan must report the smallest), several occurrences in the reference
        ref = ref[:3000] + b"ACGTACGTACGTACGTACGTAC" + ref[3000:7000] + b"ACGTACGTACGTACGTACGTAC" + ref[7000:]
        return ref, b"TTTTTTT" + b"ACGTACGTACGTACGTACGTAC" + bytes(body), 8
    if shape == "long_unmatched_start":
        return ref, rnd(400, "fbjunk") + bytes(body), 1                         # nothing in the first 256 positions: full index
    if shape == "first_kmer_everywhere":
        return b"AC" * 6000 + ref, b"AC" * 20 + bytes(body), 1                 # 6,000 occurrences of the first k-mer > list capacity
    if shape == "p0_fallthrough":
        # first match ends near the start of the reference, the next k-mer of the target is the reference's first one:
        # candidate p = 0 is in range and chosen -> `pn2 != 0` fails -> unrestricted best over all candidates
        return ref, ref[20:60] + ref[0:30] + bytes(body), 1
    ref = ref[:5000] + ref[100:160] + ref[5000:]                                # two occurrences: both in the list, sampled mode stays
    return ref, ref[100:160] + bytes(body), 8


@pytest.mark.parametrize("chunk", [64, 300])
@pytest.mark.parametrize("shape", SAMPLED_FALLBACK_SHAPES)
def test_global_sampled_index_falls_back(ctx, shape, chunk, monkeypatch):
    """the first step of the parse in sampled mode (occurrences of the first target k-mer that has any, by brute force) and
    the lookups the sampled mode cannot serve, which repeat the parse with the full index: nothing matches in the first 256
    target positions, more occurrences than the list holds, and the `pn2 == 0` fall-through of compression.cpp:134 (the
    chosen in-range candidate is reference position 0)"""
    monkeypatch.setenv("SCCG_GP_STRIDE", "8")
    monkeypatch.setenv("SCCG_GP_CHUNK", str(chunk))
    ref, tgt, expect_stride = sampled_fallback_pair(shape)
    exp = [(x.p, x.l, x.lit) for x in ol.orc_match_sequences(ref, tgt, 14, 100, True, 0)]
    got = [(x.p, x.l, x.lit) for x in ctx.match_sequences(ref, tgt, 14, 100, True, 0)]
    assert got == exp
    assert ctx.profile()["index_stride"] == expect_stride


def test_resident_reference_round_trips(ctx, monkeypatch):
    """sccg_reference_set + *_resident calls: the bytes of sccg_compress / sccg_decompress for every target, the reference
    uploaded once; target chunks smaller than the sequences (matcher launched per chunk, leftovers past the last launch)"""
    from sccg_genome_compression_b200 import synth
    monkeypatch.setenv("SCCG_PIPE_CHUNK", "8192")
    ref, t0 = synth.local_pair(60_000, synth.seed_for(2, 77))
    ref, t0 = ref.tobytes(), t0.tobytes()
    g_ref, g_tgt = synth.global_gap_pair(64_000, 60_000, synth.seed_for(1, 77))
    targets = [t0, t0[:41_234], t0 + rnd(7_777, "left"), t0[:20_000] + t0[23_000:], b"", t0[:999], g_tgt.tobytes()]
    with pytest.raises(Exception):
        ctx.compress_resident(t0, b">none")                     # no resident reference yet
    ctx.set_reference(ref)
    for i, tgt in enumerate(targets):
        hdr = b">resident %d" % i
        rc, exp, mode = ol.orc_compress(ref, tgt, hdr)
        got, gmode = ctx.compress_resident(tgt, hdr)
        assert (gmode, got) == (mode, exp), i
        assert got == ctx.compress(ref, tgt, hdr)[0]
        if tgt:
            assert ctx.decompress_resident(got) == ol.orc_decompress(ref, exp)[1], i
    # a second reference replaces the first (smaller and larger than the allocation)
    for r2 in (g_ref.tobytes(), ref[:30_000], ref + ref[:10_000]):
        ctx.set_reference(r2)
        tgt = g_tgt.tobytes() if r2 is not ref else t0
        rc, exp, mode = ol.orc_compress(r2, tgt, b">second")
        got, gmode = ctx.compress_resident(tgt, b">second")
        assert (gmode, got) == (mode, exp)
        assert ctx.decompress_resident(got) == ol.orc_decompress(r2, exp)[1]
    # caller-owned buffers, as bench.py passes them: ctypes views in, result written through a raw pointer
    import ctypes
    import numpy as np
    r_np = np.frombuffer(ref, dtype=np.uint8).copy(); t_np = np.frombuffer(t0, dtype=np.uint8).copy()
    out_np = np.zeros(2 * len(t0) + 4096, dtype=np.uint8)
    ctx.set_reference((ctypes.c_char * r_np.size).from_address(r_np.ctypes.data))
    n, gmode = ctx.compress_resident((ctypes.c_char * t_np.size).from_address(t_np.ctypes.data), b">ptr", out_np.ctypes.data, out_np.size)
    rc, exp, mode = ol.orc_compress(ref, t0, b">ptr")
    assert (gmode, bytes(out_np[:n])) == (mode, exp)
    back = np.zeros(2 * len(t0) + 4096, dtype=np.uint8)
    m = ctx.decompress_resident(exp, back.ctypes.data, back.size)
    assert bytes(back[:m]) == ol.orc_decompress(ref, exp)[1]
    ctx.clear_reference()
    with pytest.raises(Exception):
        ctx.decompress_resident(got)


def test_output_buffer_guess_too_small(monkeypatch):
    import robustness_cases
    robustness_cases.check_output_guess(emu_context, monkeypatch)


def test_lowercase_line_shapes(sweep_order):
    import robustness_cases
    c = emu_context()
    try:
        robustness_cases.check_lowercase_line_shapes(c)
    finally:
        c.close()
