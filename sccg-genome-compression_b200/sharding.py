"""Multi-GPU plumbing for whole-genome runs: the reference's unit of work is one
FASTA pair (one chromosome), pairs are independent, so they are LPT-packed onto the ranks (one process
per GPU) and only the encoded record streams are gathered back to rank 0.  torch.distributed is the
transport: NCCL between GPUs, gloo in the CPU tests.  No collective sits on a kernel's critical path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def assign_chromosomes(lengths: list[int], world_size: int) -> list[list[int]]:
    """Longest-processing-time packing of chromosome indices onto ranks; deterministic.
    Returns, per rank, the chromosome indices in the order they should be processed."""
    loads = [0] * world_size
    out: list[list[int]] = [[] for _ in range(world_size)]
    for idx in sorted(range(len(lengths)), key=lambda i: (-lengths[i], i)):
        r = min(range(world_size), key=lambda k: (loads[k], k))
        out[r].append(idx)
        loads[r] += lengths[idx]
    return out


# ------------------------------------------------------------------------------------------------
# one chromosome over several GPUs: segment-range shards of the local path
# ------------------------------------------------------------------------------------------------
SEG = 1000          # segment length L (compression.cpp:375)


def segment_ranges(ref_len: int, tgt_len: int, world_size: int) -> list[tuple[int, int]] | None:
    """Contiguous ranges [a, b) of segment-pair indices, one per rank (compression.cpp:385-392: n = min(#r, #t) pairs).
    None if the pair is too small to shard (every shard needs >= 8 pairs so that its border windows do not overlap)."""
    n_iter = min((ref_len + SEG - 1) // SEG, (tgt_len + SEG - 1) // SEG)
    if world_size < 2 or n_iter < 8 * world_size:
        return None
    base, rem = divmod(n_iter, world_size)
    out, a = [], 0
    for r in range(world_size):
        b = a + base + (1 if r < rem else 0)
        out.append((a, b))
        a = b
    return out


def shard_slices(ref, tgt, rng: tuple[int, int], is_last: bool):
    """the slices shard [a, b) needs; the last shard's target slice runs to the end of the target (leftover segments, :476-481)"""
    a, b = rng
    return ref[a * SEG:b * SEG], (tgt[a * SEG:] if is_last else tgt[a * SEG:b * SEG])


def plan_carries(infos: list[dict], ranges: list[tuple[int, int]], tgt_len: int) -> list[dict] | None:
    """From the border reports of all shards: the carry of every shard, or None when the pair has to take the unsharded
    path (T2 abort -> global mode, compression.cpp:462-473; '(' in the target -> text-level delta_encode).  Deterministic:
    every rank computes the same plan from the all-gathered reports."""
    world = len(infos)
    if any(i["abort_inside"] or i["has_paren"] for i in infos):
        return None
    # T2 windows that cross a shard border: the counter exceeds 4 at a failed, non-all-N segment whose 4 predecessors all
    # incremented it (:417-424, :454-462).  Every shard has >= 8 segments, so a window touches at most two shards.
    for r in range(1, world):
        seq = infos[r - 1]["tail_status"] + infos[r]["head_status"]
        for e in range(4, 8):
            if (seq[e] & 2) and all(seq[x] & 1 for x in range(e - 4, e + 1)):
                return None
    starts = [a * SEG for a, _ in ranges]
    ends = [b * SEG for _, b in ranges[:-1]] + [tgt_len]
    touches_start = [i["n_runs"] > 0 and i["first_run_start"] == starts[r] for r, i in enumerate(infos)]
    touches_end = [i["n_runs"] > 0 and i["last_run_start"] + i["last_run_len"] == ends[r] for r, i in enumerate(infos)]
    whole = [infos[r]["n_runs"] == 1 and touches_start[r] and touches_end[r] for r in range(world)]
    skip = [r > 0 and touches_start[r] and touches_end[r - 1] for r in range(world)]

    def cont_reaches_end(q: int) -> bool:          # a run that enters shard q at its first symbol runs to the end of the target
        while True:
            if not (touches_start[q] and whole[q]):
                return False
            if q == world - 1:
                return True
            q += 1

    carries = []
    for r in range(world):
        prev_p = 0
        for q in range(r - 1, -1, -1):
            if infos[q]["has_match"]:
                prev_p = infos[q]["last_p"]
                break
        extra = 0
        if touches_end[r]:
            q = r + 1
            while q < world and touches_start[q]:
                extra += infos[q]["first_run_len"]
                if not whole[q]:
                    break
                q += 1
        prev_run_start = 0
        for q in range(r - 1, -1, -1):
            if infos[q]["n_runs"] - (1 if skip[q] else 0) > 0:     # shard q starts at least one run of its own
                prev_run_start = infos[q]["last_run_start"]
                break
        reaches = touches_end[r] and (r == world - 1 or cont_reaches_end(r + 1))
        carries.append({"prev_p": prev_p, "skip_first_run": int(skip[r]), "extra_last_len": extra, "prev_run_start": prev_run_start,
                        "last_run_reaches_end": int(reaches), "reserved": 0})
    return carries


last_path = ""          # "sharded" / "unsharded": which way the most recent compress_sharded call went (tests, logs)


def compress_sharded(ctx, ref: bytes, tgt: bytes, header: bytes) -> tuple[bytes, int] | None:
    """compress_genome for ONE pair with the segment pairs spread over all ranks of the default process group (one GPU /
    sccg context per rank).  Every rank passes the same pair but only touches -- and only uploads -- its own slices.
    Returns (compressed_genome.txt image, mode) on rank 0 and None elsewhere.  Pairs that leave the local path (T2 abort,
    '(' in the target) or are too small to shard are compressed by rank 0 alone."""
    global last_path
    world, rank = dist.get_world_size(), dist.get_rank()
    ranges = segment_ranges(len(ref), len(tgt), world)
    carries = None
    if ranges is not None:
        r_slice, t_slice = shard_slices(ref, tgt, ranges[rank], rank == world - 1)
        info = ctx.shard_match(r_slice, t_slice, ranges[rank][0], rank == world - 1)
        infos = _all_gather_infos(info, world)                      # 20 integers per rank: the only exchange before the write
        carries = plan_carries(infos, ranges, len(tgt))
    if carries is None:                                             # unsharded fallback on rank 0
        last_path = "unsharded"
        return ctx.compress(ref, tgt, header) if rank == 0 else None
    last_path = "sharded"
    low, body = ctx.shard_write(carries[rank])
    parts = gather_streams({2 * rank: low, 2 * rank + 1: body}, dst=0)      # only the encoded streams travel
    if rank != 0:
        return None
    text = (header + b"\n" if header else b"") + b"".join(parts[2 * r] for r in range(world)) + b"\n,\n" + b"".join(parts[2 * r + 1] for r in range(world))
    return text, 0


def decompress_sharded(ctx, ref_raw, intermediate: bytes, out_path: str) -> int:
    """decompress for ONE pair with the output spread over all ranks by byte range: every rank produces its
    piece of reconstructed_genome.fa (sccg_decompress_part: only the reference chunks that piece copies from are uploaded)
    and writes it at its offset of `out_path`; nothing is gathered.  Returns the length of the whole file."""
    import os
    world, rank = dist.get_world_size(), dist.get_rank()
    off, piece, total = ctx.decompress_part(ref_raw, intermediate, rank, world)
    if rank == 0:
        with open(out_path, "wb") as f:
            f.truncate(total)
    dist.barrier()                                                  # the file exists with its final size
    fd = os.open(out_path, os.O_WRONLY)
    try:
        os.pwrite(fd, piece, off)
    finally:
        os.close(fd)
    dist.barrier()
    return total


_INFO_KEYS = ["n_segments", "abort_inside", "has_paren", "has_match", "last_p", "n_runs", "first_run_start", "first_run_len", "last_run_start", "last_run_len"]


def _all_gather_infos(info: dict, world: int) -> list[dict]:
    dev = _device_for_backend()
    mine = torch.tensor([info[k] for k in _INFO_KEYS] + list(info["head_status"]) + list(info["tail_status"]), dtype=torch.int64, device=dev)
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    out = []
    for v in allv:
        v = v.cpu().tolist()
        d = dict(zip(_INFO_KEYS, v[:len(_INFO_KEYS)]))
        d["head_status"] = v[len(_INFO_KEYS):len(_INFO_KEYS) + 4]; d["tail_status"] = v[len(_INFO_KEYS) + 4:len(_INFO_KEYS) + 8]
        out.append(d)
    return out


def _device_for_backend() -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def gather_streams(local: dict[int, bytes], dst: int = 0) -> dict[int, bytes] | None:
    """Gathers {chromosome index: encoded record stream} from every rank to `dst`.
    Sizes travel in one all_gather of int64 (per rank: count, then (index, length) pairs padded to the
    max count); payloads in one gather of uint8 tensors padded to the longest per-rank blob."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = _device_for_backend()
    items = sorted(local.items())
    count = torch.tensor([len(items)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count)
    max_items = max(int(c.item()) for c in counts)
    meta = torch.full((max(1, 2 * max_items),), -1, dtype=torch.int64, device=dev)
    for i, (idx, data) in enumerate(items):
        meta[2 * i] = idx
        meta[2 * i + 1] = len(data)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    totals = [int(m[1::2].clamp(min=0).sum().item()) for m in metas]
    pad = max(1, max(totals))
    blob = torch.zeros(pad, dtype=torch.uint8)
    cur = 0
    for _, data in items:
        blob[cur:cur + len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8) if data else blob[cur:cur]
        cur += len(data)
    blob = blob.to(dev)
    bufs = [torch.empty(pad, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == dst else None
    dist.gather(blob, bufs, dst=dst)
    if rank != dst:
        return None
    out: dict[int, bytes] = {}
    for r in range(world):
        data = bufs[r].cpu().numpy().tobytes()
        cur = 0
        m = metas[r].cpu().tolist()
        for i in range(int(counts[r].item())):
            idx, n = m[2 * i], m[2 * i + 1]
            out[idx] = data[cur:cur + n]
            cur += n
    return out
