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
