"""Multi-GPU sharding of the codec by 8-frame slabs (host-side logic only).

Every slab (cube-depth frames) is an independent key-frame group (reference README.md:10; slab loop
3d-DCT-video-encoding-OpenCL/encoder.c:203-278), so GPU g simply codes the contiguous slab range
[g*n/G, (g+1)*n/G).  The only coupling is the BIT POSITION of the Exp-Golomb stream, which is
continuous across slabs and not byte aligned (ExpGolomb.c:112-122): rank g's bits must start at
B_g = sum of the bit counts of the ranks before it.  That is G scalars: they are exchanged with one
all_gather (gloo or NCCL, or any other channel) and prefix-summed on the host; no data-path
collective exists or is needed.  Each rank's stream (coded from bit 0 of its own buffer) is then
placed at byte B_g/8 shifted right by B_g%8 bits, OR-ing the shared boundary byte.
"""
from __future__ import annotations

import numpy as np


def slab_range(nslabs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slab range [lo, hi) of `rank`."""
    return rank * nslabs // world, (rank + 1) * nslabs // world


def bit_offsets(nbits_per_rank) -> list[int]:
    """Exclusive prefix sum of the per-rank bit counts; the last entry is the total."""
    out = [0]
    for b in nbits_per_rank:
        out.append(out[-1] + int(b))
    return out


def place(dst: np.ndarray, part: np.ndarray, nbits: int, start_bit: int) -> None:
    """OR the first `nbits` bits of `part` (MSB-first bytes, coded from bit 0) into dst at start_bit."""
    if nbits == 0:
        return
    nbytes = (nbits + 7) // 8
    p = np.ascontiguousarray(part[:nbytes], np.uint8)
    byte0, sh = start_bit // 8, start_bit % 8
    if sh == 0:
        dst[byte0:byte0 + nbytes] |= p
        return
    dst[byte0:byte0 + nbytes] |= p >> sh
    spill = (p.astype(np.uint16) << (8 - sh)).astype(np.uint8)
    end = min(byte0 + 1 + nbytes, dst.size)
    dst[byte0 + 1:end] |= spill[: end - byte0 - 1]


def concatenate(parts, nbits_per_rank) -> tuple[np.ndarray, int]:
    """Per-rank streams -> the one stream the reference would have written: floor(bits/8)+1 bytes."""
    offs = bit_offsets(nbits_per_rank)
    total = offs[-1]
    out = np.zeros(total // 8 + 1, np.uint8)
    for part, nb, off in zip(parts, nbits_per_rank, offs):
        place(out, np.asarray(part, np.uint8), int(nb), off)
    return out, total


def gather_bit_counts(nbits: int, group=None) -> list[int]:
    """All ranks learn every rank's bit count (torch.distributed; eight scalars at most)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor([nbits], dtype=torch.int64, device=dev)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [int(t.item()) for t in out]


def decode_range(codec, stream, nslabs: int, rank: int, world: int, cubes_per_slab: int, start_bits=None):
    """Rank `rank` decodes its slab range of ONE concatenated stream (the reference's file layout).

    The stream carries no index, so the range's first bit has to come from somewhere: `start_bits` (the
    exclusive prefix of the encoder ranks' bit counts, `bit_offsets`) when the encoder's side information is at
    hand, otherwise `codec.eg_locate`, which runs index discovery over the stream up to the range's first cube
    (SURVEY.md 8e).  Returns (frames of the range, start bit used)."""
    lo, hi = slab_range(nslabs, rank, world)
    if hi == lo:
        return np.zeros((0, codec.height, codec.width), np.uint8), 0
    start = int(start_bits[rank]) if start_bits is not None else codec.eg_locate(stream, lo * cubes_per_slab)
    s = np.ascontiguousarray(stream, np.uint8)
    byte0 = start // 8                     # the library takes a bit offset below 8 plus whole bytes
    q, _ = codec.eg_decode_i16(s[byte0:], (hi - lo) * cubes_per_slab, start % 8)
    return codec.reconstruct_i16(q, (hi - lo) * codec.cube), start
