#!/usr/bin/env python
"""Instruction mix of seg_match_k by PHASE of the algorithm: aggregates an `ncu --page source --csv` export by source line
(nvdisasm line info of the same build) and then by the line ranges of the phases in csrc/sccg_local.cuh (located by marker
comments / function names, so the table follows the source).  usage: ncu_phase_mix.py <source_page.csv> <nvdisasm --print-line-info output> [kernel: seg_match_defer_k | seg_match_queue_k | seg_match_kILi2]"""
import csv, re, sys
from collections import defaultdict
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
src_csv, disasm = sys.argv[1:3]
kname = sys.argv[3] if len(sys.argv) > 3 else "seg_match_defer_k"
lines = open(disasm).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('//--------------------- .text.') and kname in l)
cur = None; instrs = []
for l in lines[start + 1:]:
    if l.startswith('//--------------------- .'): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l): instrs.append(cur)
rows = list(csv.reader(open(src_csv)))
secs = []; c = None
for r in rows:
    if r and r[0] == 'Kernel Name': c = {'name': r[1], 'rows': []}; secs.append(c); continue
    if r and r[0] == 'Address': c['hdr'] = r; continue
    if c is not None and len(r) > 10: c['rows'].append(r)
s = next(x for x in secs if kname.split('ILi')[0] in x['name'])
iex = s['hdr'].index('Instructions Executed')
assert len(s['rows']) == len(instrs), (len(s['rows']), len(instrs))
local = (ROOT / "sccg-genome-compression_b200" / "csrc" / "sccg_local.cuh").read_text().split('\n')
def find(pat, after=0):
    return next(i + 1 for i in range(after, len(local)) if pat in local[i])
L = {
    "lcp": find("__device__ __forceinline__ int warp_lcp"), "build": find("__device__ __forceinline__ void lm_build_index"),
    "parse": find("__device__ __forceinline__ int lm_parse"), "diag": find("__device__ __forceinline__ int lm_diag_parse"),
    "d_int": find("// ---- 2. lane i owns the mismatch-free interval"), "d_tab": find("// ---- 3a. the chunks t[u .. u+8)"),
    "d_probe": find("// ---- 3b. probe"), "dvhelp": find("static const int DV_MAX_MISWORDS"), "runs": find("__device__ __forceinline__ int lm_runs"), "fold_runs": find("__device__ __forceinline__ void lm_fold_runs"),
    "fetch": find("__device__ __forceinline__ void lm_fetch"), "kernel": find("__device__ __forceinline__ void seg_match_body"),
    "loop": find("while (seg < n_iter) {"), "claim": find("// claim the next segment"), "decide": find("int nmatch = 0, covered = 0;"),
    "epilogue": find("// \"segment consists only of N\""), "bytes": find("__global__ void seg_bytes_k"),
}
def phase(k):
    if k is None: return "other (no line info)"
    f, n = k
    if f == "sccg_common.cuh": return "upper-casing / SWAR helpers (sccg_common.cuh)"
    if f != "sccg_local.cuh": return "intrinsics headers (shuffles, ballots, atomics)"
    if n < L["runs"]: return "generic path: extension (warp_lcp / diag_lcp)"
    if n < L["build"]: return "generic path: runs of one symbol (lm_runs)"
    if n < L["fold_runs"]: return "generic path: index build (lm_build_index), lane_lcp, candidate fold"
    if n < L["parse"]: return "generic path: runs of one symbol (lm_fold_runs)"
    if n < L["dvhelp"]: return "generic path: greedy parse (lm_parse)"
    if n < L["diag"]: return "diagonal path 3: chunk hash / slot helpers (dv_hash, dv_slot)"
    if n < L["d_int"]: return "diagonal path 1: mismatching symbols -> sorted list"
    if n < L["d_tab"]: return "diagonal path 2: intervals, looked-up windows, matches"
    if n < L["d_probe"]: return "diagonal path 3a: chunk table of the looked-up windows"
    if n < L["fetch"]: return "diagonal path 3b: probe of the 250 chunks of r + clean-up"
    if n < L["kernel"]: return "global loads of the segment pair (lm_fetch)"
    if n < L["claim"]: return "upper-case + diagonal compare in registers"
    if n < L["decide"]: return "segment claiming, abort poll, L2 prefetch"
    if n < L["epilogue"]: return "path selection, staging to shared memory, match copy-out"
    return "epilogue: all-N test, seginfo, T2 window check"
ex = defaultdict(int)
for i, r in enumerate(s['rows']): ex[phase(instrs[i])] += int(r[iex])
tot = sum(ex.values())
print(f"| phase | warp instructions | share | per segment pair (249,251) |\n|---|---|---|---|")
for k, v in sorted(ex.items(), key=lambda kv: -kv[1]): print(f"| {k} | {v:,} | {100 * v / tot:.1f} % | {v / 249251:.0f} |")
print(f"| **total** | {tot:,} | 100 % | {tot / 249251:.0f} |")
