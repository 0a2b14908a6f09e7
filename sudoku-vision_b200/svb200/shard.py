"""Image-sharding of a batch across ranks: frames are independent, so the path has no
data-path collective — rank r owns the contiguous slice [r*n/W, (r+1)*n/W) — and only the final 81-byte
boards are gathered.  Works with any torch.distributed backend (nccl on GPUs, gloo in CPU tests)."""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split: the first n % world ranks get one extra frame."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_boards(local_digits, n_total: int, group=None):
    """All ranks end with the (n_total, 81) uint8 boards in global frame order.  local_digits is this
    rank's (n_local, 81) uint8 tensor; shards may be ragged (padded for the collective)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    mx = max(e - s for s, e in sizes)
    s, e = sizes[rank]
    if local_digits.shape[0] != e - s:
        raise ValueError(f"rank {rank}: expected {e - s} boards, got {local_digits.shape[0]}")
    pad = torch.zeros((mx, 81), dtype=torch.uint8, device=local_digits.device)
    pad[: e - s] = local_digits
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([out[r][: sizes[r][1] - sizes[r][0]] for r in range(world)], 0)
