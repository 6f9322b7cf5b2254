"""Deterministic synthetic genome pairs for the BASELINE.json configurations (the bundled
genome/hg18|hg19 chr19 FASTA pair is absent from the reference checkout and there is no network).

All generators return (reference symbols, target symbols) as numpy uint8 arrays at the level of
read_genomes_from_files' output (compression.cpp:181-220): raw, case-preserved, newline-free.
numpy default_rng(seed); seeds follow SURVEY.md section 8d: 0x5CC60000 + config*256 + chromosome.
"""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
HG19_LENGTHS = [249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663, 146364022,
                141213431, 135534747, 135006516, 133851895, 115169878, 107349540, 102531392, 90354753,
                81195210, 78077248, 59128983, 63025520, 48129895, 51304566, 155270560, 59373566]
CHR1_LEN = HG19_LENGTHS[0]


def seed_for(config: int, chrom: int = 0) -> int:
    return 0x5CC60000 + config * 256 + chrom


def random_bases(rng: np.random.Generator, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint8)
    step = 1 << 26
    for a in range(0, n, step):
        b = min(n, a + step)
        out[a:b] = ACGT[rng.integers(0, 4, size=b - a, dtype=np.uint8)]
    return out


def substitute(rng: np.random.Generator, seq: np.ndarray, rate: float) -> None:
    """in place: `rate` of the positions get a different base (A->C->G->T->A shifted by 1..3)"""
    n = seq.size
    k = int(n * rate)
    if k == 0:
        return
    pos = rng.integers(0, n, size=k)
    lut = np.zeros(256, dtype=np.uint8)
    lut[ACGT] = np.arange(4, dtype=np.uint8)
    cur = lut[seq[pos]]
    seq[pos] = np.where(np.isin(seq[pos], ACGT), ACGT[(cur + rng.integers(1, 4, size=k, dtype=np.uint8)) % 4], seq[pos])


def lowercase_runs(rng: np.random.Generator, seq: np.ndarray, frac: float = 0.5, lo: int = 10, hi: int = 10_000) -> None:
    """in place: about `frac` of the sequence in lowercase runs with log-uniform lengths lo..hi"""
    n = seq.size
    mean = (hi - lo) / np.log(hi / lo)
    count = int(n / mean * 1.2) + 16
    lens = np.exp(rng.uniform(np.log(lo), np.log(hi), size=count)).astype(np.int64)
    gaps = np.exp(rng.uniform(np.log(lo), np.log(hi), size=count)) * ((1.0 - frac) / max(frac, 1e-9))
    gaps = np.maximum(gaps.astype(np.int64), 1)
    pieces = np.empty(2 * count, dtype=np.int64)
    pieces[0::2] = gaps
    pieces[1::2] = lens
    ends = np.cumsum(pieces)
    m = int(np.searchsorted(ends, n)) + 1
    ends = np.minimum(ends[:m], n)
    starts = np.concatenate(([0], ends[:-1]))
    for k in range(1, m, 2):                       # odd pieces are lowercase runs
        a, b = int(starts[k]), int(ends[k])
        if b > a:
            seq[a:b] |= 0x20


def compensated_indels(rng: np.random.Generator, seq: np.ndarray, count: int, max_d: int = 10, max_gap: int = 200) -> np.ndarray:
    """+d insertion at s paired with a -d deletion <= max_gap later: length and segment alignment preserved"""
    n = seq.size
    if count == 0:
        return seq
    sites = np.sort(rng.integers(1000, n - 2000, size=count))
    sites = sites[np.concatenate(([True], np.diff(sites) > 1000))]
    # assembly gaps carry no indels: keep sites whose neighbourhood is free of N
    sites = sites[(seq[sites] != ord("N")) & (seq[sites + 300] != ord("N")) & (seq[np.maximum(sites - 300, 0)] != ord("N"))]
    parts, cur = [], 0
    for s in sites.tolist():
        d = int(rng.integers(1, max_d + 1))
        gap = int(rng.integers(20, max_gap + 1))
        parts.append(seq[cur:s])
        parts.append(ACGT[rng.integers(0, 4, size=d)])
        parts.append(seq[s:s + gap])
        cur = s + gap + d
    parts.append(seq[cur:])
    out = np.concatenate(parts)
    assert out.size == n
    return out


def local_pair(n: int, seed: int, snp: float = 0.001, indels_per_mbp: float = 20.0, lowercase: float = 0.5,
               n_block: bool = True) -> tuple[np.ndarray, np.ndarray]:
    """BASELINE config 2 shape (chr1-sized: n = 249,250,621): stays on the local segment-matching path."""
    rng = np.random.default_rng(seed)
    ref = random_bases(rng, n)
    if n_block and n > 100_000:
        edge = min(10_000, n // 50)
        ref[:edge] = ord("N"); ref[n - edge:] = ord("N")
        blk = min(3_000_000, n // 80)
        a = n // 2
        ref[a:a + blk] = ord("N")                       # co-located gap (centromere-like)
    tgt = ref.copy()
    substitute(rng, tgt, snp)
    tgt = compensated_indels(rng, tgt, int(n / 1e6 * indels_per_mbp))
    for s in rng.integers(0, max(1, n - 300), size=max(1, n // 5_000_000)).tolist():   # a few target-only N runs of 1..200
        tgt[s:s + int(rng.integers(1, 201))] = ord("N")
    if lowercase > 0:
        lowercase_runs(rng, tgt, lowercase)
    return ref, tgt


def global_gap_pair(n_ref: int, n_tgt: int, seed: int, snp: float = 0.001, lowercase: float = 0.5) -> tuple[np.ndarray, np.ndarray]:
    """BASELINE config 1 shape (chr19: 63,811,651 vs 59,128,983): the target is the reference with its N gaps
    resized/removed + SNPs, no target-side deletion of non-N symbols -> local aborts, global succeeds."""
    rng = np.random.default_rng(seed)
    body = random_bases(rng, n_tgt)
    extra = n_ref - n_tgt
    cuts = np.sort(rng.integers(n_tgt // 20, n_tgt - n_tgt // 20, size=4))
    gap = [extra // 2, extra // 4, extra // 8, extra - extra // 2 - extra // 4 - extra // 8]
    parts, cur = [], 0
    for c, g in zip(cuts.tolist(), gap):
        parts += [body[cur:c], np.full(g, ord("N"), dtype=np.uint8)]
        cur = c
    parts.append(body[cur:])
    ref = np.concatenate(parts)
    tgt = body.copy()
    substitute(rng, tgt, snp)
    if lowercase > 0:
        lowercase_runs(rng, tgt, lowercase)
    return ref, tgt


def divergent_pair(n: int, seed: int, sub: float = 0.05, blocks: int = 20) -> tuple[np.ndarray, np.ndarray]:
    """BASELINE config 3 shape (chr21: 48,129,895): 5 % substitutions, permuted 0.1-1 Mbp blocks and an early
    5 kb insertion -> global path, lookup heavy, mostly literals."""
    rng = np.random.default_rng(seed)
    lead_n = min(10_000_000, n // 5)
    ref = random_bases(rng, n)
    ref[:lead_n] = ord("N")
    tgt = ref[lead_n:].copy()
    substitute(rng, tgt, sub)
    m = tgt.size
    ins_at = m // 100
    tgt = np.concatenate([tgt[:ins_at], random_bases(rng, 5000), tgt[ins_at:]])
    bl = np.sort(rng.integers(m // 10, m - m // 10, size=blocks))
    pieces = [tgt[:bl[0]]] + [tgt[bl[i]:bl[i + 1]] for i in range(blocks - 1)] + [tgt[bl[-1]:]]
    inner = pieces[1:-1]
    order = rng.permutation(len(inner))
    tgt = np.concatenate([pieces[0]] + [inner[i] for i in order] + [pieces[-1]])
    return ref, tgt
