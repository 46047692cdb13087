"""Multi-GPU sharding of independent frames.

A zstd stream may be any concatenation of frames (lib/zstd.h:160-162; precedent contrib/pzstd), so
N ranks compress contiguous, equal-byte partitions of the frame list with no data-path exchange;
the only exchange is the final gather of the variable-length compressed buffers: all_gather(sizes), then
point-to-point transfers of exactly each rank's bytes into their final place on the destination rank
(NCCL on GPUs, gloo in CPU tests).
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


def split_one_frame(frame_size: int, world_size: int, alignment: int = 512 << 10) -> List[Tuple[int, int]]:
    """(begin, size) per rank of ONE frame that the ranks compress together (ZSTDB200_compressFramePart; the reference's
    precedent are ZSTDMT's jobs, zstdmt_compress.c:1168-1227): shares start on multiples of `alignment`
    (ZSTDB200_framePartAlignment(): a chunk of the candidate walk), the last one takes the rest.  A rank also needs the
    ZSTDB200_framePartHalo() bytes in front of its share.  A rank with nothing to do gets (-1, 0); an empty frame is
    rank 0's."""
    chunks = (frame_size + alignment - 1) // alignment
    out, done = [], 0
    for r in range(world_size):
        n = (chunks * (r + 1)) // world_size - (chunks * r) // world_size
        b = done * alignment
        e = min(frame_size, (done + n) * alignment)
        out.append((b, e - b) if (n or (r == 0 and frame_size == 0)) else (-1, 0))
        done += n
    return out


def _all_sizes(n: int, device, group=None) -> List[int]:
    """every rank's byte count, in rank order (one tiny all_gather + one host read)"""
    ws = dist.get_world_size(group)
    mine = torch.tensor([n], dtype=torch.int64, device=device)
    allv = torch.empty(ws, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allv, mine, group=group)
    return [int(v) for v in allv.tolist()]


def gather_compressed(local: torch.Tensor, dst: int = 0, group=None, out: torch.Tensor = None, async_op: bool = False):
    """Gather variable-length uint8 tensors to rank `dst` in rank order (what contrib/pzstd's writer thread does with
    the frames of its workers, Pzstd.cpp:335): sizes by one all_gather, then every other rank SENDS exactly its bytes
    and `dst` RECEIVES each payload straight at its final offset of one buffer — no padding, no concatenation copy;
    all transfers are one batched group (NCCL: one grouped launch over NVLink, gloo in the CPU tests).
    out: destination buffer on `dst` (>= sum of sizes); when `local` already is out[:n] the own part is not copied.
    Returns (sizes, gathered tensor or None, works): with async_op=True the transfers may still be in flight, call
    .wait() on every work (zstd_b200.sharding.wait_all) before touching `local` / the gathered bytes."""
    ws = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = _all_sizes(local.numel(), local.device, group)
    ops, gathered = [], None
    if rank == dst:
        total = sum(sizes)
        if out is None:
            out = torch.empty(total, dtype=torch.uint8, device=local.device)
        assert out.numel() >= total
        off = 0
        for r in range(ws):
            if r == rank:
                if sizes[r] and out.data_ptr() + off != local.data_ptr():
                    out[off:off + sizes[r]].copy_(local)
            elif sizes[r]:
                ops.append(dist.P2POp(dist.irecv, out[off:off + sizes[r]], r, group))
            off += sizes[r]
        gathered = out[:total]
    elif local.numel():
        ops.append(dist.P2POp(dist.isend, local, dst, group))
    works = dist.batch_isend_irecv(ops) if ops else []
    if not async_op:
        wait_all(works)
        works = []
    return sizes, gathered, works


def wait_all(works) -> None:
    for w in works:
        w.wait()


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
