"""Multi-GPU sharding of independent frames.

A zstd stream may be any concatenation of frames (lib/zstd.h:160-162; precedent contrib/pzstd), so
N ranks compress contiguous, equal-byte partitions of the frame list with no data-path exchange;
the only collective is the final gather of the variable-length compressed buffers:
all_gather(sizes) then gather of payloads padded to the largest size (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def partition_frames(frame_sizes: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) frame ranges per rank, balanced by input bytes (work is ~ bytes)."""
    total = sum(frame_sizes)
    out, f, acc = [], 0, 0
    n = len(frame_sizes)
    for r in range(world_size):
        target = total * (r + 1) / world_size
        b = f
        while f < n and (acc + frame_sizes[f] / 2 <= target or (r == world_size - 1)):
            acc += frame_sizes[f]
            f += 1
        out.append((b, f))
    return out


def split_into_frames(total_size: int, frame_size: int) -> List[Tuple[int, int]]:
    """(offset, size) of the independent frames a buffer is cut into (config 3: 64 MiB frames)."""
    return [(o, min(frame_size, total_size - o)) for o in range(0, max(total_size, 1), frame_size)] if total_size else [(0, 0)]


def gather_compressed(local: torch.Tensor, dst: int = 0, group=None):
    """Gather variable-length uint8 tensors to rank `dst` in rank order.
    Returns (list_of_sizes, concatenated tensor on dst or None elsewhere)."""
    ws = dist.get_world_size(group)
    rank = dist.get_rank(group)
    size = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(size) for _ in range(ws)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes) if sizes else 0
    padded = torch.zeros(mx, dtype=torch.uint8, device=local.device)
    padded[: local.numel()] = local
    bufs = [torch.empty(mx, dtype=torch.uint8, device=local.device) for _ in range(ws)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst, group=group)
    if rank != dst:
        return sizes, None
    return sizes, torch.cat([b[:s] for b, s in zip(bufs, sizes)])


def gather_frame_sizes(c_sizes: Sequence[int], d_sizes: Sequence[int], dst: int = 0, group=None):
    """All ranks' per-frame (compressed, decompressed) sizes in rank order on rank `dst` (None elsewhere): what the seek
    table over the gathered frames needs (zstd_b200.seek_table, contrib/seekable_format)."""
    ws = dist.get_world_size(group)
    rank = dist.get_rank(group)
    objs = [None] * ws if rank == dst else None
    dist.gather_object((list(c_sizes), list(d_sizes)), objs, dst=dst, group=group)
    if rank != dst:
        return None
    cs = [c for o in objs for c in o[0]]
    ds = [d for o in objs for d in o[1]]
    return cs, ds
