"""Deterministic parity cases for the SCCG hot path.

Each case is (name, reference symbols, target symbols, header) at the level of
read_genomes_from_files' output (compression.cpp:181-220): raw, case-preserved, newline-free
symbols.  FASTA_CASES are file-level inputs for the FASTA reader / CLI tests.
"""
from __future__ import annotations

import random
from dataclasses import dataclass


@dataclass
class Case:
    name: str
    ref: bytes
    tgt: bytes
    header: bytes = b">tgt case"
    expect_mode: int | None = None      # 0 local, 1 global, None = whatever the reference does
    lossless: bool = True               # False: the reference itself does not round-trip this input


def rnd(n: int, seed, alphabet: bytes = b"ACGT") -> bytes:
    r = random.Random(seed if isinstance(seed, (int, str, bytes)) else repr(seed))
    return bytes(r.choice(alphabet) for _ in range(n))


def mutate(seq: bytes, positions) -> bytes:
    b = bytearray(seq)
    for p in positions:
        b[p] = {65: 67, 67: 71, 71: 84, 84: 65}.get(b[p], 65)   # A->C->G->T->A
    return bytes(b)


def lower_at(seq: bytes, ranges) -> bytes:
    b = bytearray(seq)
    for lo, hi in ranges:
        b[lo:hi] = bytes(b[lo:hi]).lower()
    return bytes(b)


def put(seq: bytes, pos: int, piece: bytes) -> bytes:
    return seq[:pos] + piece + seq[pos + len(piece):]


def _differs_from(c: int) -> int:
    return 67 if c == 65 else 65


def cases() -> list[Case]:
    out: list[Case] = []
    R3 = rnd(3000, "R3")
    out.append(Case("identical_3000", R3, R3, expect_mode=0))

    # A: SNPs (incl. adjacent-segment literals 2999/3000), target-only N runs, lowercase single/run/at-end
    R5 = rnd(5000, "A")
    t = mutate(R5, [100, 1500, 1507, 2999, 3000])
    t = put(t, 2200, b"N" * 60)
    t = put(t, 4000, b"N")
    t = lower_at(t, [(10, 20), (50, 51), (4990, 5000)])
    out.append(Case("A_snps_N_lowercase", R5, t, expect_mode=0))

    # B: dropped segment (all-N / poly-A target segment vs ordinary reference segment)
    R4 = rnd(4000, "B")
    out.append(Case("B1_dropped_allN_segment", R4, put(R4, 1000, b"N" * 1000), expect_mode=0, lossless=False))
    out.append(Case("B2_dropped_polyA_segment", R4, put(R4, 1000, b"A" * 1000), expect_mode=0, lossless=False))

    # C: target longer / shorter than the reference
    out.append(Case("C1_target_longer", R3, R3 + rnd(1234, "C1"), expect_mode=0))
    out.append(Case("C2_target_shorter", R3, R3[:2500], expect_mode=0))

    # D: tails
    out.append(Case("D1_tail5_dropped", R3, R3[:2005], expect_mode=0, lossless=False))
    out.append(Case("D2_tail12_k2_pass", R3, R3[:2012], expect_mode=0))
    out.append(Case("D2b_tail14", R3, R3[:2014], expect_mode=0))
    out.append(Case("D2c_tail10", R3, R3[:2010], expect_mode=0))
    out.append(Case("D2d_tail9_dropped", R3, R3[:2009], expect_mode=0, lossless=False))
    out.append(Case("D3_ref_tail5", R3[:2005], R3, expect_mode=0, lossless=False))
    out.append(Case("D5_ref_tail9", R3[:2009], R3, expect_mode=0, lossless=False))
    out.append(Case("D4_ref_tail13", R3[:2013], R3, expect_mode=0))
    out.append(Case("D6_ref_tail10", R3[:2010], R3, expect_mode=0))

    # E: global trigger; insertion is absorbed, deletions > 100 lose the parse
    R20 = rnd(20000, "E")
    out.append(Case("E1_global_insertion3000", R20, R20[:2500] + rnd(3000, "E1") + R20[2500:], expect_mode=1))
    out.append(Case("E2_global_deletion3000", R20, R20[:2500] + R20[5500:], expect_mode=1))
    e3 = R20[:2500] + rnd(2500, "E3") + R20[2500:9000] + R20[9060:14000] + R20[14150:]
    out.append(Case("E3_global_ins_del60_del150", R20, e3, expect_mode=1))

    # F: tie-breaks inside one segment
    X = rnd(30, "X30"); Y = rnd(40, "Y40")
    r = rnd(1000, "F1r")
    r1 = put(put(r, 0, X), 130, X)
    r1 = put(r1, 30, bytes([_differs_from(r1[160])]))          # make the two copies extend equally (30)
    tF = X + rnd(970, "F1t")
    tF = put(tF, 30, bytes([84 if r1[30] != 84 and r1[160] != 84 else 71]))
    out.append(Case("F1_tie_p0_displaced", r1, tF, expect_mode=0))
    r2 = put(put(r, 1, X), 131, X)
    out.append(Case("F2_tie_p1_kept", r2, tF, expect_mode=0))
    def isolate(t: bytes, rr: bytes, at: int, copies) -> bytes:
        """make the symbols just before / after the X copy in the target differ from every reference context"""
        before = {rr[c - 1] for c in copies}; after = {rr[c + 30] for c in copies}
        t = put(t, at - 1, bytes([next(c for c in b"ACGT" if c not in before)]))
        return put(t, at + 30, bytes([next(c for c in b"ACGT" if c not in after)]))
    r3 = put(put(put(put(r, 500, Y), 100, X), 560, X), 700, X)
    t3 = Y + rnd(8, "F3gap") + X + rnd(922, "F3t")
    out.append(Case("F3_tie_nearest_prev_end", r3, isolate(t3, r3, 48, (100, 560, 700)), expect_mode=0))
    r4 = put(put(put(r, 480, Y), 400, X), 638, X)
    out.append(Case("F4_tie_equidistant_negative_delta", r4, isolate(t3, r4, 48, (400, 638)), expect_mode=0))

    # G5: IUPAC symbols, one of them lowercase
    g5 = put(put(R3[:2000], 300, b"R"), 301, b"y")
    out.append(Case("G5_iupac", put(put(R3[:2000], 300, b"R"), 301, b"Y"), g5, expect_mode=0))

    # H: global with N line: single, run, trailing run; lowercase block
    body = rnd(12000, "H")
    rH = b"N" * 1500 + body[:6000] + b"N" * 700 + body[6000:]
    tH = body[:3000] + b"N" + body[3000:5000] + b"N" * 120 + body[5000:9000].lower() + body[9000:] + b"NN"
    out.append(Case("H_global_N_line", rH, tH, expect_mode=1))

    # I: pn2 == 0 fall-through to an out-of-range candidate
    Xi = rnd(30, "Xi"); Yi = rnd(20, "Yi"); Zi = rnd(30, "Zi")
    rI = Xi + Yi + rnd(4950, "I1") + Xi + Zi + rnd(15000, "I2")
    tI = Yi + Xi + Zi + rnd(3000, "I3") + rI[6000:]
    out.append(Case("I_global_pn2_zero_jump", rI, tI, expect_mode=1))

    # J: grammar symbols in the target corrupt the body (compress side is still pinned)
    out.append(Case("J_grammar_symbols", R3[:2000], put(R3[:2000], 500, b"(7,"), expect_mode=0, lossless=False))
    # J2..J6: what the reference's TEXT-level delta_encode (compression.cpp:262-292) does with literal parentheses
    #   J2 a lone '(' pairs with the next token's ')' -> stoi("(501") throws: exit 1, file left un-rewritten (rc_compress = 1)
    #   J3 "()" has no comma -> skipped (:274-277)      J4 '(' with no ')' after it ends the loop (:269-270)
    #   J5 a literal look-alike "(12x,5)" is rewritten and poisons the chain      J6 the same in global mode
    out.append(Case("J2_lone_paren_stoi_throws", R3[:2000], put(R3[:2000], 500, b"("), expect_mode=0, lossless=False))
    out.append(Case("J3_empty_parens", R3[:2500], put(R3[:2500], 1500, b"()"), expect_mode=0, lossless=False))
    out.append(Case("J4_unclosed_paren_in_tail", R3[:2000], R3[:2000] + b"ACGT(ACGTACGT", expect_mode=0, lossless=False))
    out.append(Case("J5_token_lookalike", R3[:3000], put(R3[:3000], 1200, b"(12x,5)"), expect_mode=0, lossless=False))
    rJ = rnd(20000, "J6")
    out.append(Case("J6_global_literal_paren", rJ, put(rJ[:2500] + rnd(3000, "J6ins") + rJ[2500:], 3000, b"(7,"), expect_mode=1, lossless=False))

    # K: co-located N runs compress to plain tokens
    k = rnd(1000, "K1") + b"N" * 1000 + rnd(300, "K2") + b"N" * 400 + rnd(300, "K3")
    out.append(Case("K_colocated_N", k, k, expect_mode=0))

    # T2 counter: fail, bad, fail, fail, bad (cnt=5, no abort), good -> stays local
    R8 = rnd(8000, "T2")
    t = bytearray(R8)
    def fail(i): t[i * 1000:(i + 1) * 1000] = rnd(1000, f"T2f{i}")
    def bad(i):                                                # >50 % literals but one 14-mer hit
        seg = bytearray(rnd(1000, f"T2b{i}")); seg[100:130] = R8[i * 1000 + 100:i * 1000 + 130]
        t[i * 1000:(i + 1) * 1000] = seg
    fail(0); bad(1); fail(2); fail(3); bad(4)
    out.append(Case("T2_counter_no_abort", R8, bytes(t), expect_mode=0, lossless=False))
    fail(5)
    out.append(Case("T2_counter_abort_at_6th", R8, bytes(t), expect_mode=1))
    # an all-N target segment resets the counter
    t2 = bytearray(R8)
    for i in (0, 1, 2, 3):
        t2[i * 1000:(i + 1) * 1000] = rnd(1000, f"T2g{i}")
    t2[4000:5000] = b"N" * 1000
    t2[5000:6000] = rnd(1000, "T2g5")
    out.append(Case("T2_counter_reset_by_allN", R8, bytes(t2), expect_mode=0, lossless=False))

    # low-complexity: bucket skew (H7)
    out.append(Case("polyA_both", b"A" * 2500, b"A" * 2500, expect_mode=0))
    out.append(Case("allN_both", b"N" * 3000, b"N" * 3000, expect_mode=0))
    out.append(Case("dinucleotide_repeat", b"AC" * 1200, b"AC" * 700 + b"CA" * 500, expect_mode=0))
    two = rnd(4000, "two", b"AC")
    out.append(Case("two_letter_alphabet", two, mutate(two, range(7, 4000, 97))))

    # lowercase everywhere / alternating single lowercase
    out.append(Case("all_lowercase", R3, R3.lower(), expect_mode=0))
    alt = bytearray(R3[:1200])
    for i in range(0, 1200, 2):
        alt[i] = alt[i] + 32
    out.append(Case("alternating_case", R3[:1200], bytes(alt), expect_mode=0))
    out.append(Case("lowercase_reference", R3.lower(), R3, expect_mode=0))

    # degenerate sizes
    out.append(Case("tiny_target_9", R3, R3[:9], expect_mode=0, lossless=False))
    out.append(Case("tiny_target_14", R3, R3[:14], expect_mode=0))
    out.append(Case("empty_reference", b"", R3[:1500], expect_mode=0))
    out.append(Case("no_header", R3, mutate(R3, [5, 1999]), header=b"", expect_mode=0, lossless=False))

    # denser random divergence: 2 % substitutions + compensated indels (stays local)
    rr = random.Random("dense")
    base = rnd(30000, "dense")
    t = bytearray(base)
    for p in rr.sample(range(30000), 600):
        t[p] = rr.choice(b"ACGT")
    t = bytes(t)
    for s in range(2000, 29000, 3000):
        d = rr.randint(1, 8)
        t = t[:s] + rnd(d, f"ins{s}") + t[s:s + 150] + t[s + 150 + d:]
    out.append(Case("dense_2pct_compensated_indels", base, lower_at(t, [(100, 700), (15000, 15001), (29990, 30000)])))

    # global with many SNPs and several small indels (all deletions <= 100)
    base = rnd(60000, "gl")
    t = bytearray(base[:1000] + rnd(6000, "gl_ins") + base[1000:])
    for p in rr.sample(range(8000, len(t)), 300):
        t[p] = rr.choice(b"ACGT")
    t = bytes(t)
    t = t[:20000] + t[20040:30000] + rnd(77, "gl2") + t[30000:45000] + t[45099:]
    t = put(t, 33000, b"N" * 333)
    out.append(Case("global_snps_small_indels", b"N" * 250 + base[:30000] + b"NNN" + base[30000:], t, expect_mode=1))
    return out


@dataclass
class FastaCase:
    name: str
    ref_file: bytes
    tgt_file: bytes
    lossless: bool = True


def fasta_cases() -> list[FastaCase]:
    def wrap(seq: bytes, w=50, nl=b"\n"):
        return nl.join(seq[i:i + w] for i in range(0, len(seq), w)) + nl
    R = rnd(2000, "fa")
    T = lower_at(mutate(R, [77, 1200]), [(300, 340)])
    out = [
        FastaCase("plain", b">ref\n" + wrap(R), b">tgt plain\n" + wrap(T)),
        FastaCase("G1_no_header", b">ref\n" + wrap(R), wrap(T), lossless=False),
        FastaCase("G2_two_records", b">ref\n" + wrap(R), b">rec1\n" + wrap(T[:1000]) + b">REC2\n" + wrap(T[1000:1995] + b"ggacc"), lossless=False),
        FastaCase("G3_crlf", b">ref\r\n" + wrap(R, nl=b"\r\n"), b">tgt crlf\r\n" + wrap(T, nl=b"\r\n"), lossless=False),
        FastaCase("G4_width60", b">ref\n" + wrap(R, 60), b">tgt w60\n" + wrap(T, 60), lossless=False),
        FastaCase("G6_multi_record_reference", b">r1\n" + wrap(R[:1000]) + b">r2\n" + wrap(R[1000:]), b">tgt\n" + wrap(T)),
        FastaCase("no_trailing_newline", b">ref\n" + wrap(R)[:-1], b">tgt\n" + wrap(T)[:-1], lossless=False),
        FastaCase("blank_lines_and_spaces", b">ref\n\n" + wrap(R[:1000]) + b"\n  \n" + wrap(R[1000:]), b"\n>tgt\n" + wrap(T[:500]) + b" \t\n" + wrap(T[500:]), lossless=False),
    ]
    return out
