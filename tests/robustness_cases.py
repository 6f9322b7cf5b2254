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
