"""Utterance-level sharding over ranks (one process per GPU; no data-path collective).

The velocity net has no padding mask at inference (acoustic.py:288-318), so a
batch is exact only for equal-length items: utterances are bucketed by length, cut into batches of at
most ``batch`` equal-length items, and the batches are dealt to ranks longest-first onto the least
loaded rank (cost ~ B*N*(L_lin + 16384*N)).  The only collective is the final
reduction of (audio seconds, wall seconds) for the aggregate throughput.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple


def batch_cost(n_items: int, length: int, l_lin: float = 111.5e6) -> float:
    return n_items * length * (l_lin + 16384.0 * length)


def plan_batches(lengths: Sequence[int], batch: int) -> List[Tuple[int, List[int]]]:
    """-> [(length, [utterance indices])] with every batch holding equal-length utterances."""
    by_len: Dict[int, List[int]] = {}
    for i, n in enumerate(lengths):
        by_len.setdefault(int(n), []).append(i)
    out = []
    for n in sorted(by_len, reverse=True):
        idx = by_len[n]
        for k in range(0, len(idx), batch):
            out.append((n, idx[k:k + batch]))
    return out


def assign_batches(lengths: Sequence[int], world: int, batch: int) -> List[List[Tuple[int, List[int]]]]:
    """Greedy longest-processing-time assignment of equal-length batches to ``world`` ranks."""
    batches = plan_batches(lengths, batch)
    batches.sort(key=lambda b: batch_cost(len(b[1]), b[0]), reverse=True)
    load = [0.0] * world
    per_rank: List[List[Tuple[int, List[int]]]] = [[] for _ in range(world)]
    for b in batches:
        r = min(range(world), key=lambda i: (load[i], i))
        per_rank[r].append(b)
        load[r] += batch_cost(len(b[1]), b[0])
    return per_rank


def reduce_throughput(audio_seconds: float, wall_seconds: float, group=None, device="cpu") -> Tuple[float, float]:
    """SUM of audio seconds and MAX of wall seconds over ranks -> (total audio s, slowest rank's s)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return audio_seconds, wall_seconds
    a = torch.tensor([audio_seconds], dtype=torch.float64, device=device)
    w = torch.tensor([wall_seconds], dtype=torch.float64, device=device)
    dist.all_reduce(a, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(w, op=dist.ReduceOp.MAX, group=group)
    return float(a.item()), float(w.item())


def broadcast_blob(blob, src: int = 0, group=None, device="cpu"):
    """Optional init-time weight distribution: rank ``src`` packs, everyone receives the packed blob
    (NCCL over NVLink when ``device`` is a CUDA device; gloo on CPU).  ``blob``: uint8 torch tensor or None."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return blob
    n = torch.tensor([0 if blob is None else blob.numel()], dtype=torch.int64, device=device)
    dist.broadcast(n, src=src, group=group)
    if blob is None:
        blob = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    else:
        blob = blob.to(device)
    dist.broadcast(blob, src=src, group=group)
    return blob
