"""Sharding of one corpus over the ranks of a job: line-aligned byte ranges, and the one
exchange step of the path — an all-gather of per-shard {matches, newlines} from which every rank derives its
record, line-number and byte-offset bases.  Host logic only; the backend is whatever torch.distributed was
initialised with (NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def line_aligned_cuts(data: np.ndarray, world: int) -> list[int]:
    """world + 1 cut offsets: shard r is data[cuts[r]:cuts[r+1]]; every interior cut follows a newline, so a line
    belongs to exactly one shard and no halo is needed."""
    n = int(data.size)
    cuts = [0]
    for r in range(1, world):
        target = max(cuts[-1], n * r // world)
        if target >= n:
            cuts.append(n)
            continue
        # forward to the end of the line that contains byte `target - 1` (or stay, if it ends right there)
        if target == 0 or data[target - 1] == 10:
            cuts.append(target)
            continue
        rel = np.flatnonzero(data[target:min(n, target + (1 << 20))] == 10)
        if rel.size == 0:
            rel = np.flatnonzero(data[target:] == 10)
        cuts.append(target + int(rel[0]) + 1 if rel.size else n)
    cuts.append(n)
    return cuts


def bases_from_counts(counts: list[tuple[int, int]], cuts: list[int]) -> list[tuple[int, int, int]]:
    """per rank: (index of its first record in the global list, line-number base, byte-offset base)"""
    out = []
    m = nl = 0
    for r, (matches, newlines) in enumerate(counts):
        out.append((m, nl, cuts[r]))
        m += matches
        nl += newlines
    return out


def all_gather_counts(matches: int, newlines: int, device=None) -> list[tuple[int, int]]:
    """the path's only collective: {matches, newlines} of every rank, in rank order"""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [(int(matches), int(newlines))]
    mine = torch.tensor([int(matches), int(newlines)], dtype=torch.int64, device=device)
    allv = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(allv, mine)
    return [(int(v[0]), int(v[1])) for v in allv]


class CountExchange:
    """all_gather_counts() for a loop: the same collective with its buffers allocated once (one small host-to-device
    copy, one all-gather into a tensor, one read back per call instead of a dozen tiny tensor operations)"""

    def __init__(self, device="cuda"):
        import torch
        import torch.distributed as dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.device = device
        if self.world > 1:
            pinned = str(device).startswith("cuda")
            self.host = torch.zeros(2, dtype=torch.int64, pin_memory=pinned)
            self.mine = torch.zeros(2, dtype=torch.int64, device=device)
            self.all = torch.zeros(2 * self.world, dtype=torch.int64, device=device)

    def __call__(self, matches: int, newlines: int) -> list[tuple[int, int]]:
        if self.world == 1:
            return [(int(matches), int(newlines))]
        import torch.distributed as dist
        self.host[0] = int(matches)
        self.host[1] = int(newlines)
        self.mine.copy_(self.host, non_blocking=True)
        dist.all_gather_into_tensor(self.all, self.mine)
        v = self.all.tolist()
        return [(v[2 * r], v[2 * r + 1]) for r in range(self.world)]


def tiled_cuts(block: np.ndarray, reps: int, world: int) -> list[int]:
    """line_aligned_cuts() for the logical corpus ``block`` repeated ``reps`` times (``block`` ends with a newline),
    computed from the block alone: the nominal cut n*r/world moves forward to the end of the line it falls in."""
    B = int(block.size)
    if B == 0 or block[-1] != 10:
        raise ValueError("the block must end with a newline")
    n = B * reps
    nl = np.flatnonzero(block == 10)
    cuts = [0]
    for r in range(1, world):
        target = max(cuts[-1], n * r // world)
        if target >= n:
            cuts.append(n)
            continue
        b, off = divmod(target, B)
        if off == 0 or block[off - 1] == 10:
            cuts.append(target)
            continue
        j = int(np.searchsorted(nl, off))          # first newline at or after `off`; the block's last byte is one
        cuts.append(b * B + int(nl[j]) + 1)
    cuts.append(n)
    return cuts


def materialize_tiled(dblock, lo: int, hi: int):
    """bytes [lo, hi) of the logical corpus made of repetitions of the device tensor ``dblock``"""
    import torch
    B = int(dblock.numel())
    parts = []
    p = lo
    if p < hi and p % B != 0:
        take = min(B - p % B, hi - p)
        parts.append(dblock[p % B:p % B + take])
        p += take
    whole = (hi - p) // B
    if whole > 0:
        parts.append(dblock.repeat(whole))
        p += whole * B
    if p < hi:
        parts.append(dblock[:hi - p])
    if not parts:
        return dblock[:0].clone()
    return torch.cat(parts) if len(parts) > 1 else parts[0].clone()
