"""CPU tests of the multi-GPU host logic (world_size 2 over gloo): slab-range partition, the
all_gather of bit counts, and the bit-shifted concatenation, checked against the oracle's
whole-clip stream.  The per-rank streams come from the oracle here (no GPU in this tier); on the
GPU box the same functions are fed by libdct3d (tests/test_gpu_parity.py::test_sharded_*)."""
import os
import sys

import numpy as np
import pytest

from conftest import pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_ranges_cover_everything():
    sh = pkg("sharding")
    for n in (1, 7, 32, 128):
        for world in (1, 2, 4, 8):
            r = [sh.slab_range(n, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))


def test_concatenate_matches_one_shot(oracle, synth):
    sh = pkg("sharding")
    clip = synth.natural(64, 48, 40, 1)
    whole, bits = oracle.encode_u8(clip, 8, 0)
    for world in (2, 3, 5):
        parts, nb = [], []
        for g in range(world):
            lo, hi = sh.slab_range(5, g, world)
            s, b = oracle.encode_u8(clip[lo * 8:hi * 8], 8, 0) if hi > lo else (np.zeros(1, np.uint8), 0)
            parts.append(s)
            nb.append(b)
        cat, total = sh.concatenate(parts, nb)
        assert total == bits and cat.tobytes() == whole.tobytes()


def test_two_phase_placement_matches_one_shot(oracle, synth):
    """The rule dct3d_encode_u8_place implements (shift to the phase, copy whole bytes, OR the boundary bytes) on the
    host model: every byte of the stream is written exactly once and the result is the one-shot stream, also when a
    range ends on a byte boundary or is empty."""
    sh = pkg("sharding")
    clip = synth.natural(64, 48, 40, 1)
    whole, bits = oracle.encode_u8(clip, 8, 0)
    for world in (1, 2, 3, 5, 8):
        parts, nb = [], []
        for g in range(world):
            lo, hi = sh.slab_range(5, g, world)
            s, b = oracle.encode_u8(clip[lo * 8:hi * 8], 8, 0) if hi > lo else (np.zeros(1, np.uint8), 0)
            parts.append(s)
            nb.append(b)
        cat, total = sh.concatenate_two_phase(parts, nb)
        assert total == bits and cat.tobytes() == whole.tobytes()
    # synthetic parts with every combination of phases, including ranges that end exactly on a byte boundary
    rng = np.random.default_rng(5)
    for trial in range(200):
        n = int(rng.integers(1, 6))
        nb = [int(rng.choice([0, 8, 16, 24, int(rng.integers(1, 200))])) for _ in range(n)]
        parts = []
        for b in nb:
            p = rng.integers(0, 256, b // 8 + 1).astype(np.uint8)
            if b % 8:
                p[-1] &= (0xFF << (8 - b % 8)) & 0xFF
            else:
                p[-1] = 0
            parts.append(p)
        a, ta = sh.concatenate(parts, nb)
        b2, tb = sh.concatenate_two_phase(parts, nb)
        assert ta == tb and a.tobytes() == b2.tobytes(), (trial, nb)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    import torch.distributed as dist
    from oracle import oracle as O
    sh = importlib.import_module("3ddctvideoencoding_b200.sharding")
    synth = importlib.import_module("3ddctvideoencoding_b200.synth")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    clip = synth.natural(64, 48, 32, 1)                 # every rank can regenerate the clip from the seed
    lo, hi = sh.slab_range(4, rank, world)
    stream, nbits = O.encode_u8(clip[lo * 8:hi * 8], 8, 0)
    counts = sh.gather_bit_counts(nbits)
    offs = sh.bit_offsets(counts)
    # rank 0 collects the parts (tiny here) and concatenates
    parts = [None] * world
    dist.all_gather_object(parts, stream.tobytes())
    if rank == 0:
        cat, total = sh.concatenate([np.frombuffer(p, np.uint8) for p in parts], counts)
        whole, bits = O.encode_u8(clip, 8, 0)
        q.put((total == bits and cat.tobytes() == whole.tobytes(), offs))
    # decode side: each rank decodes its own slab range from its start bit (side information = offs)
    whole, bits = O.encode_u8(clip, 8, 0)
    qd, end = O.eg_decode_cubes(whole, (hi - lo) * 48, 8, offs[rank])
    ok = end == offs[rank + 1] and (qd == O.quantized_cubes(clip[lo * 8:hi * 8], 8, 0)).all()
    oks = [None] * world
    dist.all_gather_object(oks, bool(ok))
    if rank == 0:
        q.put(all(oks))
    dist.destroy_process_group()


def test_two_ranks_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    same, offs = q.get(timeout=10)
    assert same and offs[0] == 0 and offs[1] > 0
    assert q.get(timeout=10)


def _worker_shm(rank, world, tag, q):
    """The multi-process flow of bench.py's e2e path with the GPU replaced by the oracle and the host model of the
    placement kernel: code the rank's slab range from bit 0, exchange the bit counts through the shared-memory table,
    place whole bytes into the shared stream, OR the boundary byte once the predecessor has landed."""
    sys.path.insert(0, ROOT)
    import importlib
    from oracle import oracle as O
    sh = importlib.import_module("3ddctvideoencoding_b200.sharding")
    synth = importlib.import_module("3ddctvideoencoding_b200.synth")
    W, H, nslabs = 64, 48, 7
    clip = synth.natural_slabs(W, H, 8, 0, nslabs, 3)
    whole, bits = O.encode_u8(clip, 8, 0)
    xch = sh.ShmExchange("dct3d_test_x_%s" % tag, world, rank, create=False)
    shared = sh.SharedStream("dct3d_test_s_%s" % tag, bits // 8 + 1, create=False)
    ok = True
    for rep in range(20):                                # the table is reused without a reset
        lo, hi = sh.slab_range(nslabs, rank, world)
        part, nb = O.encode_u8(clip[lo * 8:hi * 8], 8, 0) if hi > lo else (np.zeros(1, np.uint8), 0)
        offs = sh.bit_offsets(xch.all_gather(nb))
        fb = sh.place_shifted(shared.array, sh.shift_to_phase(part, nb, offs[rank] % 8), offs[rank], nb, rank == world - 1)
        xch.signal()
        if offs[rank] % 8:
            xch.wait_for(rank - 1)
            shared.array[offs[rank] // 8] |= fb
        # wait for everybody before checking (and before the next repetition overwrites the stream)
        done = xch.all_gather(1)
        ok = ok and offs[-1] == bits and shared.array.tobytes() == whole.tobytes() and sum(done) == world
        xch.all_gather(2)
    q.put(bool(ok))


@pytest.mark.parametrize("world", [2, 3])
def test_shared_memory_exchange_and_placement(world):
    import torch.multiprocessing as mp
    sh = pkg("sharding")
    tag = "%d_%d" % (os.getpid(), world)
    # rank 0's role of creating the mappings is played by the test itself
    xch = sh.ShmExchange("dct3d_test_x_%s" % tag, world, 0, create=True)
    shared = sh.SharedStream("dct3d_test_s_%s" % tag, 1 << 20, create=True)
    try:
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker_shm, args=(r, world, tag, q)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=300)
            assert p.exitcode == 0
        assert all(q.get(timeout=10) for _ in range(world))
    finally:
        xch.unlink()
        shared.unlink()
