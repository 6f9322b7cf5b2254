"""Checks shared by the emulator (CPU) and the GPU test files: malformed / hostile inputs must fail cleanly and leave
the context usable (round-1 advisor findings)."""
import ctypes as C

import pytest

import oracle_lib as ol
import sccg_b200
from cases import rnd


def check_far_out_of_range_tokens(ctx):
    """tokens whose absolute position / end lies far outside the reference (decompression.cpp:223-229: ERROR + exit 1, or
    substr's out_of_range): error code, no out-of-bounds read, and a valid call on the same context afterwards"""
    ref = rnd(500, "far")
    good = b">ok\n\n,\n(10,100)ACGT(20,50)"
    rc, exp = ol.orc_decompress(ref, good)
    assert rc == 0
    for enc, codes in [
        (b"(2000000000,1000)", (sccg_b200.SCCG_E_BOUNDS,)),                       # far past the end
        (b"(2000000000,1000000000)", (sccg_b200.SCCG_E_FORMAT, sccg_b200.SCCG_E_BOUNDS)),   # abs + len wraps negative in int
        (b"(-7,3)", (sccg_b200.SCCG_E_FORMAT,)),                                  # negative absolute position: substr throws
        (b"(100,10)(-2000000000,5)", (sccg_b200.SCCG_E_FORMAT,)),
        (b"AC(400,200)GT", (sccg_b200.SCCG_E_BOUNDS,)),
    ]:
        rc_o, _ = ol.orc_reconstruct(ref, enc, b"", b"")
        assert rc_o != 0, enc                                                       # the reference (via the oracle) exits 1 too
        with pytest.raises(sccg_b200.SccgError) as e:
            ctx.reconstruct(ref, enc, b"", b"")
        assert e.value.code in codes, (enc, e.value.code)
        with pytest.raises(sccg_b200.SccgError):
            ctx.decompress(ref, b">h\n\n,\n" + enc)
        assert ctx.decompress(ref, good) == exp                                     # the context survived


def check_long_header_line(ctx):
    """a header line longer than the library's staging area (2 MiB+): decompress must reproduce it, not overrun a buffer"""
    ref = rnd(3000, "hdr")
    header = b">" + b"x" * (2 * 1024 * 1024 + 123)
    inter = header + b"\n(5,10)\n,\n(0,1000)ACGT(1000,500)"
    rc, exp = ol.orc_decompress(ref, inter)
    assert rc == 0 and exp.startswith(header + b"\n")
    assert ctx.decompress(ref, inter) == exp
    buf = C.create_string_buffer(len(exp) + 64)
    n = ctx.decompress_into(ref, inter, C.cast(buf, C.c_void_p), len(exp) + 64)
    assert buf.raw[:n] == exp


def check_stale_shard_write(ctx):
    """sccg_shard_write after another entry point reused the context must fail (SCCG_E_ARG), not emit garbage"""
    from sccg_genome_compression_b200 import synth
    ref, tgt = synth.local_pair(40_000, synth.seed_for(2, 77))
    ref, tgt = ref.tobytes(), tgt.tobytes()
    ctx.shard_match(ref[:20_000], tgt[:20_000], 0, False)
    ctx.compress(ref, tgt, b">other call")                                          # reuses / reallocates the shard's buffers
    with pytest.raises(sccg_b200.SccgError) as e:
        ctx.shard_write({"prev_p": 0, "skip_first_run": 0, "extra_last_len": 0, "prev_run_start": 0, "last_run_reaches_end": 0, "reserved": 0})
    assert e.value.code == sccg_b200.SCCG_E_ARG


def check_streaming_and_fasta_into(ctx, chunk_env=None):
    """sccg_decompress_stream / _fasta_stream deliver the image of sccg_decompress in order from the double buffer; errors come
    before the first piece; sccg_compress_fasta_into == sccg_compress_fasta"""
    import os
    from sccg_genome_compression_b200 import synth
    old = os.environ.get("SCCG_PIPE_CHUNK")
    if chunk_env:
        os.environ["SCCG_PIPE_CHUNK"] = str(chunk_env)
    try:
        for shape in ("local", "gap"):
            if shape == "local":
                ref, tgt = synth.local_pair(230_000, synth.seed_for(2, 95))
            else:
                ref, tgt = synth.global_gap_pair(160_000, 150_000, synth.seed_for(1, 95))
            ref, tgt = ref.tobytes(), tgt.tobytes()
            rc, inter, mode = ol.orc_compress(ref, tgt, b">stream me")
            rc, exp = ol.orc_decompress(ref, inter)
            assert rc == 0
            pieces = ctx.decompress_stream(ref, inter)
            assert [o for o, _ in pieces] == sorted(o for o, _ in pieces) and pieces[0][0] == 0
            assert b"".join(p for _, p in pieces) == exp
            if chunk_env:
                assert len(pieces) > 2                                           # both buffers were reused
            # FASTA file images
            ref_fa = b">ref\n" + b"\n".join(ref[i:i + 60] for i in range(0, len(ref), 60)) + b"\n"
            tgt_fa = b">stream me\n" + b"\n".join(tgt[i:i + 50] for i in range(0, len(tgt), 50)) + b"\n"
            pieces = ctx.decompress_stream(ref_fa, inter, fasta=True)
            assert b"".join(p for _, p in pieces) == exp
            buf = C.create_string_buffer(len(inter) + 4096)
            n, m = ctx.compress_fasta_into(ref_fa, tgt_fa, C.cast(buf, C.c_void_p).value, len(inter) + 4096)
            assert (m, buf.raw[:n]) == (mode, inter)
        ref = rnd(500, "far")
        for bad in (b">h\n\n,\n(400,200)", b">h\n\n,\n(a,5)"):
            with pytest.raises(sccg_b200.SccgError):
                ctx.decompress_stream(ref, bad)
    finally:
        if chunk_env:
            if old is None:
                os.environ.pop("SCCG_PIPE_CHUNK", None)
            else:
                os.environ["SCCG_PIPE_CHUNK"] = old


def check_decoder_tolerance(ctx):
    """streams the reference accepts although no compressor writes them (decompression.cpp:126-207 expands the run lists into
    positions and SORTS them; an out-of-range lowercase position is only a warning, :255-262): same bytes as the oracle.
    Streams on which the reference runs into undefined behaviour or throws are errors in both.  Two pinned divergences
    (INTEGRATION.md section 5): a negative token length (`substr` wraps to "rest of the reference" there, SCCG_E_FORMAT here),
    and run-list TEXT outside the grammar that the reference's find / stoi splitter happens to digest (stoi ignores what follows
    a number: "38,1)3,36," reads as 38, 1, 36 there; SCCG_E_FORMAT here)."""
    import random
    ref = rnd(6000, "tol")
    body = b"(0,1000)ACGT(1000,2000)TT(2100,1500)"
    cases = [
        (body, b"", b"(50,10)(-45,3),(12,20)(30,100)"),                    # lowercase runs out of order and overlapping
        (body, b"", b"(4400,50)(-4400,10)4600,"),                          # ... and past the end of the sequence (warning only)
        (body, b"(100,2)(-95,3),50", b"(60,10)(-58,3)"),                   # N runs out of order (distinct positions)
        (body, b"(4000,3)(-3990,2)", b"(10,4000)"),
        (b"(0,100)ACGT(50,20)", b"2,(5,3)(100,2)", b"(3,4)10,200"),        # optional commas, singles
        (body, b"(10,5)(2,5)", b""),                                       # overlapping N runs: out-of-range read in the reference
        (body, b"", b"(50,10)(-60,3)"),                                    # negative lowercase position: undefined behaviour
        (body, b"(5000,10)", b""),                                         # N run past the end
    ]
    r = random.Random("tolerance")
    runs = []
    pos = 0
    for _ in range(3000):                                                  # many runs, shuffled: several tiles and scan blocks
        pos += r.randint(1, 3); ln = r.randint(1, 2); runs.append((pos, ln)); pos += ln
    big_body = b"(0,6000)" + b"(-100,6000)" * 2
    for which in ("low", "n"):
        rr = runs[:]; r.shuffle(rr)
        if which == "low":
            rr += [(x + 1, 3) for x, _ in rr[:200]]                        # overlaps (lowercase only)
        text = bytearray(); prev = 0
        for st, ln in rr:
            text += (b"%d," % (st - prev)) if ln == 1 and r.random() < 0.7 else b"(%d,%d)" % (st - prev, ln)
            prev = st
        cases.append((big_body, bytes(text) if which == "n" else b"", bytes(text) if which == "low" else b""))
    for enc, nidx, low in cases:
        rc_o, out_o = ol.orc_reconstruct(ref, enc, nidx, low)
        if rc_o != 0:
            with pytest.raises(sccg_b200.SccgError):
                ctx.reconstruct(ref, enc, nidx, low)
            continue
        assert ctx.reconstruct(ref, enc, nidx, low) == out_o, (enc[:40], nidx[:40], low[:40])
        inter = b">tolerant\n" + low + b"\n" + (nidx if nidx else b"") + b"\n" + enc
        if nidx:                                                            # the file-level entry points (pipelined, streaming) take the same road
            rc2, exp2 = ol.orc_decompress(ref, inter)
            assert rc2 == 0 and ctx.decompress(ref, inter) == exp2
            assert b"".join(p for _, p in ctx.decompress_stream(ref, inter)) == exp2
    # pinned divergence
    rc_o, out_o = ol.orc_reconstruct(ref, b"AC(10,-3)GT", b"", b"")
    assert rc_o == 0 and out_o.startswith(b"AC" + ref[10:48].upper())
    with pytest.raises(sccg_b200.SccgError) as e:
        ctx.reconstruct(ref, b"AC(10,-3)GT", b"", b"")
    assert e.value.code == sccg_b200.SCCG_E_FORMAT, str(e.value)
    for low in (b"38,1)3,36,", b"(20,17)(41,1)(3,17)(23(,1)", b"12,(41,3,(36,1)4,", b"5,,7,"):
        rc_o, _ = ol.orc_reconstruct(ref, body, b"", low)
        assert rc_o == 0, low                                                # stoi stops at the first non-digit, find(')') / find(',') skip ahead
        with pytest.raises(sccg_b200.SccgError) as e:
            ctx.reconstruct(ref, body, b"", low)
        assert e.value.code == sccg_b200.SCCG_E_FORMAT, (low, str(e.value))
    assert ctx.reconstruct(ref, body, b"", b"(50,10)") == ol.orc_reconstruct(ref, body, b"", b"(50,10)")[1]      # the context survived


def check_output_guess(make_ctx, monkeypatch):
    """compress_device assembles the local-mode image into a buffer whose size is a GUESS made before the matcher has finished
    (no host round trip in between); the device refuses a guess that is too small and the host assembles again.  Forced
    here with a guess of 0 bytes on fresh contexts: plain local pair with a leftover tail, a target with literal '('
    (text-level delta pass), a pair that falls back to global mode; then the same context again with the normal guess."""
    import random
    from sccg_genome_compression_b200 import synth
    ref, tgt = synth.local_pair(60_000, synth.seed_for(2, 41))
    ref, tgt = ref.tobytes(), tgt.tobytes()
    r = random.Random(41)
    paren = bytearray(tgt)
    for piece in (b"(12,3)", b"(", b")(7,", b"(5,"):
        at = r.randrange(0, len(paren) - 16); paren[at:at + len(piece)] = piece
    gref, gtgt = synth.global_gap_pair(90_000, 80_000, synth.seed_for(1, 41))
    pairs = [(ref, tgt + rnd(2500, "tail")), (ref, bytes(paren) + b"(ACGT"), (gref.tobytes(), gtgt.tobytes()), (ref[:999], tgt[:700]), (b"", tgt[:100])]
    for rf, tg in pairs:
        rc, exp, mode = ol.orc_compress(rf, tg, b">guess")
        for guess in ("0", "17", None):
            ctx = make_ctx()
            try:
                if guess is None: monkeypatch.delenv("SCCG_OUT_GUESS", raising=False)
                else: monkeypatch.setenv("SCCG_OUT_GUESS", guess)
                for _ in range(2):                                                   # second call: the buffer has grown, the env guess still applies
                    if rc != 0:
                        with pytest.raises(sccg_b200.SccgError) as e:
                            ctx.compress(rf, tg, b">guess")
                        assert e.value.code == sccg_b200.SCCG_E_STOI and e.value.partial == exp
                    else:
                        assert ctx.compress(rf, tg, b">guess") == (exp, mode)
                monkeypatch.delenv("SCCG_OUT_GUESS", raising=False)
                if rc == 0:
                    assert ctx.compress(rf, tg, b">guess") == (exp, mode)
            finally:
                ctx.close()


def check_lowercase_line_shapes(ctx):
    """the lowercase-run line (compression.cpp:341-367) on shapes that stress tile and word borders of the run-list kernels: runs
    that span several tiles, tiles without any run between two runs, one-symbol runs on 64-symbol word and tile borders, a run
    that ends with the sequence, no run at all / everything lowercase"""
    import random
    r = random.Random(20261018)
    T = 131072
    def build(n, runs):
        t = bytearray(rnd(n, ("lowline", n)))
        for a, b in runs:
            t[a:b] = bytes(t[a:b]).lower()
        return bytes(t)
    shapes = []
    n = 5 * T + 777
    shapes.append((n, [(100, 4 * T + 5)]))                                        # one run over four tile borders
    shapes.append((n, [(T - 1, T), (T, T + 1)]))                                  # (adjacent: really one run of two over the border)
    shapes.append((n, [(T - 1, T), (T + 1, T + 2), (2 * T - 1, 2 * T + 1), (3 * T + 63, 3 * T + 64), (3 * T + 64, 3 * T + 66)]))
    shapes.append((n, [(5, 6), (4 * T + 700, 4 * T + 701)]))                      # three empty tiles between two one-symbol runs
    shapes.append((n, [(0, 1), (n - 1, n)]))                                      # first and last symbol
    shapes.append((n, [(0, n)]))                                                  # everything
    shapes.append((n, []))                                                        # nothing
    shapes.append((n, [(2 * T + 64 * k, 2 * T + 64 * k + 63) for k in range(40)]))  # a run in every chunk, ending one before the chunk border
    shapes.append((n, [(63 + 64 * k, 65 + 64 * k) for k in range(0, 600, 3)]))    # runs across chunk borders
    shapes.append((n, [(n - 3, n)]))                                              # ends with the sequence (longer than one symbol)
    shapes.append((T, [(T - 1, T)]))                                              # exactly one tile, last symbol
    shapes.append((3 * T, [(T, 2 * T)]))                                          # tile-aligned run
    for _ in range(6):                                                            # random mixtures, many tiles
        n = r.randrange(2 * T, 5 * T)
        runs, p = [], 0
        while p < n:
            p += r.choice([1, 2, 7, 64, 500, 20000, 140000])
            ln = r.choice([1, 1, 2, 3, 63, 64, 65, 1000, 17000, 135000, 270000])
            if p < n: runs.append((p, min(n, p + ln)))
            p += ln
        shapes.append((n, runs))
    for n, runs in shapes:
        tgt = build(n, runs)
        ref = tgt.upper()
        rc, exp, mode = ol.orc_compress(ref, tgt, b">low")
        assert rc == 0 and mode == 0
        assert ctx.compress(ref, tgt, b">low") == (exp, mode), (n, runs[:4])
