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

import os

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


def shift_to_phase(part: np.ndarray, nbits: int, phase: int) -> np.ndarray:
    """Host model of stream_shift_kernel: bit (phase + i) of the result = bit i of `part`, the first `phase` bits
    zero; (phase + nbits) // 8 + 1 bytes."""
    out = np.zeros((phase + nbits) // 8 + 1, np.uint8)
    place(out, np.asarray(part, np.uint8), nbits, phase)
    return out


def place_shifted(dst: np.ndarray, shifted: np.ndarray, start_bit: int, nbits: int, last: bool) -> int:
    """Host model of dct3d_encode_u8_place: whole bytes of a range already at phase start_bit % 8 are COPIED (not
    OR-ed) to byte start_bit // 8 onwards.  The first byte is skipped and returned when the range starts inside a byte
    (the caller ORs it in once the predecessor's bytes have landed); the byte after the last bit is written only when
    the range has bits in it or closes the stream."""
    phase, byte0 = start_bit % 8, start_bit // 8
    if nbits == 0:
        if last and phase == 0:
            dst[byte0] = 0
        return 0
    endp = phase + nbits
    first = 1 if phase else 0
    lastb = endp // 8 if (endp % 8 or last) else endp // 8 - 1
    if lastb >= first:
        dst[byte0 + first:byte0 + lastb + 1] = shifted[first:lastb + 1]
    return int(shifted[0]) if phase else 0


def concatenate_two_phase(parts, nbits_per_rank) -> tuple[np.ndarray, int]:
    """The placement the GPUs perform: shift every part to its phase, copy whole bytes, OR the shared
    boundary bytes afterwards.  The destination starts as garbage to show that every byte is written exactly once."""
    offs = bit_offsets(nbits_per_rank)
    total = offs[-1]
    out = np.full(total // 8 + 1, 0xA5, np.uint8)
    n = len(parts)
    firsts = []
    for g, (part, nb) in enumerate(zip(parts, nbits_per_rank)):
        sh = shift_to_phase(np.asarray(part, np.uint8), int(nb), offs[g] % 8)
        firsts.append(place_shifted(out, sh, offs[g], int(nb), g == n - 1))
    for g in range(1, n):
        if offs[g] % 8:
            out[offs[g] // 8] |= firsts[g]
    return out, total


class SharedStream:
    """The clip's one stream as a shared-memory mapping that every rank (process) places its range into."""

    def __init__(self, name: str, nbytes: int, create: bool):
        import mmap
        self.path = os.path.join("/dev/shm", name)
        flags = os.O_RDWR | (os.O_CREAT if create else 0)
        fd = os.open(self.path, flags, 0o600)
        try:
            if create:
                os.ftruncate(fd, nbytes)
            self.mm = mmap.mmap(fd, nbytes)
        finally:
            os.close(fd)
        self.array = np.frombuffer(self.mm, np.uint8)
        self.nbytes = nbytes

    def unlink(self):
        try:
            os.unlink(self.path)
        except OSError:
            pass


class ShmExchange:
    """The one exchange step of the sharded codec, "N scalars through the host" (SURVEY.md 8e), for ranks that are
    processes of one node: a shared-memory table with one 64-byte line per rank.  all_gather() publishes this rank's
    bit count and returns everybody's; signal() / wait_for() order a rank's boundary byte after its predecessor's
    copy.  Sequence numbers make the table reusable without a reset; values are double-buffered by sequence parity
    because a fast rank may publish exchange k+1 while a slow one still reads exchange k.  (x86 stores are ordered:
    the value is written before the sequence number that announces it.)"""
    SEQ, VAL0, VAL1, FLAG = 0, 1, 2, 3

    def __init__(self, name: str, world: int, rank: int, create: bool):
        self.shm = SharedStream(name, world * 64, create)
        self.tab = self.shm.array.view(np.int64).reshape(world, 8)
        if create:
            self.tab[:] = 0
        self.world, self.rank, self.seq, self.fseq = world, rank, 0, 0

    def all_gather(self, value: int) -> list[int]:
        self.seq += 1
        slot = self.VAL0 + (self.seq & 1)
        self.tab[self.rank, slot] = value
        self.tab[self.rank, self.SEQ] = self.seq
        seqs = self.tab[:, self.SEQ]
        while (seqs < self.seq).any():
            pass
        return [int(v) for v in self.tab[:, slot]]

    def signal(self):
        self.fseq += 1
        self.tab[self.rank, self.FLAG] = self.fseq

    def wait_for(self, peer: int):
        """Blocks until `peer` has signalled as many times as this rank has."""
        while self.tab[peer, self.FLAG] < self.fseq:
            pass

    def unlink(self):
        self.shm.unlink()


def sharded_encode(codec, frames, shared: np.ndarray, rank: int, world: int, group=None):
    """One rank's part of a multi-process encode: phase 1 on the GPU, all_gather of the bit counts, phase 2 straight
    into the shared stream, then the boundary byte.  Returns the exclusive prefix of the bit counts (world + 1)."""
    import torch.distributed as dist
    nbits = codec.encode_u8_range(frames)
    counts = gather_bit_counts(nbits, group) if world > 1 else [nbits]
    offs = bit_offsets(counts)
    fb = codec.encode_u8_place(offs[rank], rank == world - 1, shared, shared.size if isinstance(shared, np.ndarray) else shared.numel())
    if world > 1:
        dist.barrier(group)                 # the predecessor's bytes have landed
        if offs[rank] % 8:
            shared[offs[rank] // 8] |= fb
        dist.barrier(group)
    return offs


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
    hint = int(start_bits[rank + 1]) if start_bits is not None and len(start_bits) > rank + 1 else 0
    frames, _ = codec.decode_u8_range(stream, start, (hi - lo) * codec.cube, hint)
    return frames, start


def bind_to_gpu_numa(index: int):
    """Pins the calling process to the CPUs of GPU `index`'s NUMA node (sysfs local_cpulist of its PCI function), so that
    its host threads and the page-locked buffers it allocates afterwards (first touch) are local to the GPU's PCIe root.
    Returns a short description, or None when the topology cannot be read."""
    import subprocess
    try:
        bdf = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bdf.count(":") == 2 and len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        path = f"/sys/bus/pci/devices/{bdf}/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        node = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()
        return f"gpu {index} ({bdf}): numa node {node}, {len(cpus)} cpus"
    except Exception:   # noqa: BLE001
        return None
