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
