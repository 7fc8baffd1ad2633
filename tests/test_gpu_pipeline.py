"""GPU tests of the pipelined host-buffer paths, the sharded (range + place) encode, the multi-GPU object and the
error edges of the stream reader.  Everything goes through the C ABI (ctypes); the checker is the CPU oracle or the
library's own one-shot path where the oracle has already pinned it (tests/test_gpu_parity.py)."""
import ctypes as C

import numpy as np
import pytest

from conftest import pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codec_mod():
    return pkg("codec")


def one_shot(codec_mod, clip, W, H, cube):
    """The stream of the whole clip coded as ONE chunk (the path tests/test_gpu_parity.py pins to the oracle)."""
    with codec_mod.Codec(W, H, cube) as c:
        c.set_option("chunk_frames", 1 << 20 if cube == 8 else (1 << 20))
        stream, nbits = c.encode_u8(clip)
        assert c.stat("chunks") <= 1
        return stream, nbits, c.decode_u8(stream, clip.shape[0])


@pytest.mark.parametrize("W,H,F,cube,chunk", [(128, 64, 40, 8, 8), (200, 24, 64, 8, 24), (64, 32, 20, 4, 4), (256, 64, 16, 8, 16)])
def test_pipelined_calls_equal_one_shot(codec_mod, oracle, synth, W, H, F, cube, chunk):
    """dct3d_encode_u8 / dct3d_decode_u8 as a pipeline of chunks (device-chained bit position) give the bytes and the
    frames of the one-shot calls, and the stream is the oracle's."""
    clip = synth.natural(W, H, F, 21)
    want, wbits, wdec = one_shot(codec_mod, clip, W, H, cube)
    with codec_mod.Codec(W, H, cube) as c:
        c.set_option("chunk_frames", chunk)
        stream, nbits = c.encode_u8(clip)
        assert c.stat("chunks") == -(-(F // cube * cube) // chunk)
        assert nbits == wbits and stream.tobytes() == want.tobytes()
        dec = c.decode_u8(stream, F)
        assert c.stat("chunks") == -(-(F // cube * cube) // chunk)
        assert (dec == wdec).all()
        q = c.quantize_u8(clip).astype(np.int32)
        ref, ref_bits = oracle.eg_encode_cubes(q, cube, cap=5 * q.size + 64)
        assert ref_bits == nbits and ref[: nbits // 8 + 1].tobytes() == stream.tobytes()
        # too small a buffer is reported, not overrun
        small = np.zeros(stream.size - 3, np.uint8)
        nb, ny = C.c_uint64(), C.c_size_t()
        rc = c.L.dct3d_encode_u8(c.h, clip.ctypes.data, F, small.ctypes.data, small.size, C.byref(nb), C.byref(ny))
        assert rc == pkg("_lib").E_OVERFLOW
        # and the context is still usable
        again, nbits2 = c.encode_u8(clip)
        assert nbits2 == nbits and again.tobytes() == stream.tobytes()


@pytest.mark.parametrize("kind,piece", [("natural", 4096), ("natural", 20480), ("noise", 8192), ("constant", 4096)])
def test_decode_parses_the_stream_piece_by_piece(codec_mod, synth, kind, piece):
    """The pipelined decoder uploads the stream in pieces and parses each piece as it arrives (counts, overhang and list
    ranks carried from piece to piece): same frames and end bit as the one-shot decode, for piece sizes that cut codes,
    cubes and slabs anywhere; a stream that ends early is reported."""
    W, H, F = 128, 64, 64
    clip = synth.natural(W, H, F, 23) if kind == "natural" else synth.noise(W, H, F, 24) if kind == "noise" else synth.constant(W, H, F, 77)
    with codec_mod.Codec(W, H, 8) as c:
        c.set_option("chunk_frames", 1 << 20)
        stream, nbits = c.encode_u8(clip, cap=4 * clip.size + 4096)
        want = c.decode_u8(stream, F)
        c.set_option("chunk_frames", 16)
        c.set_option("piece_bytes", piece)
        assert stream.size > 2 * piece
        got, end = c.decode_u8_range(stream, 0, F)
        assert end == nbits and (got == want).all()
        # from a bit offset inside a longer buffer, with trailing bytes after the clip
        pad = np.concatenate([np.full(5, 0xFF, np.uint8), np.zeros(0, np.uint8)])
        shifted = pkg("sharding").shift_to_phase(stream, nbits, 3)
        buf = np.concatenate([pad, shifted, np.full(9000, 0xFF, np.uint8)])
        got2, end2 = c.decode_u8_range(buf, 5 * 8 + 3, F)
        assert end2 == 5 * 8 + 3 + nbits and (got2 == want).all()
        # truncated streams: an error at every cut position class
        for cut in (stream.size // 3, stream.size - piece - 7, stream.size - 2):
            with pytest.raises(codec_mod.Dct3dError):
                c.decode_u8(stream[:cut], F)
        # streaming decode asks for more instead
        assert c.stream_decode(stream[: stream.size // 2], 0, F) is None
        res = c.stream_decode(stream, 0, F)
        assert res is not None and res[1] == nbits and (res[0] == want).all()


def test_pipelined_noise_content(codec_mod, synth):
    """Dense content (3.3 bit/sample): long lists, large stream; chunks of one slab."""
    W, H, F = 128, 64, 32
    clip = synth.noise(W, H, F, 2)
    want, wbits, wdec = one_shot(codec_mod, clip, W, H, 8)
    with codec_mod.Codec(W, H, 8) as c:
        c.set_option("chunk_frames", 8)
        stream, nbits = c.encode_u8(clip, cap=4 * clip.size + 4096)
        assert nbits == wbits and stream.tobytes() == want.tobytes()
        assert (c.decode_u8(stream, F) == wdec).all()


def test_stream_shift_dev_all_phases(codec_mod):
    import torch
    sh = pkg("sharding")
    rng = np.random.default_rng(3)
    dev = torch.device("cuda", 0)
    with codec_mod.Codec(64, 64, 8) as c:
        for nbits in (1, 7, 8, 31, 32, 33, 1000, 65536 + 5):
            src = rng.integers(0, 256, nbits // 8 + 1).astype(np.uint8)
            src[-1] &= (0xFF << (8 - nbits % 8)) & 0xFF if nbits % 8 else 0
            d_src = torch.zeros(src.size + 16, dtype=torch.uint8, device=dev)
            d_src[: src.size] = torch.from_numpy(src).to(dev)
            for phase in range(8):
                cap = ((nbits + phase + 31) // 32 + 1) * 4
                d_dst = torch.full((cap + 8,), 0xEE, dtype=torch.uint8, device=dev)
                torch.cuda.synchronize()
                c.stream_shift_dev(d_src, nbits, phase, d_dst, cap)
                torch.cuda.synchronize()
                got = d_dst.cpu().numpy()
                want = sh.shift_to_phase(src, nbits, phase)
                assert (got[: want.size] == want).all(), (nbits, phase)
                assert (got[want.size:cap] == 0).all() and (got[cap:] == 0xEE).all()
        with pytest.raises(codec_mod.Dct3dError):
            c.stream_shift_dev(d_src, 100, 3, d_dst, 8)


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_range_and_place_build_the_one_stream(codec_mod, synth, world):
    """Phase 1 + phase 2 per rank (dct3d_encode_u8_range / _place) into a garbage-filled buffer, boundary bytes OR-ed
    afterwards: the result is the one-shot stream; every range then decodes from its global start bit."""
    sh = pkg("sharding")
    W, H, F = 128, 64, 40
    clip = synth.natural(W, H, F, 9)
    want, wbits, wdec = one_shot(codec_mod, clip, W, H, 8)
    nslabs = F // 8
    ctxs = [codec_mod.Codec(W, H, 8) for _ in range(world)]
    try:
        for c in ctxs:
            c.set_option("chunk_frames", 8)
        counts = []
        for g, c in enumerate(ctxs):
            lo, hi = sh.slab_range(nslabs, g, world)
            counts.append(c.encode_u8_range(clip[lo * 8:hi * 8]))
        offs = sh.bit_offsets(counts)
        assert offs[-1] == wbits
        out = np.full(wbits // 8 + 1, 0x5A, np.uint8)
        firsts = [c.encode_u8_place(offs[g], g == world - 1, out) for g, c in enumerate(ctxs)]
        for g in range(1, world):
            if offs[g] % 8:
                out[offs[g] // 8] |= firsts[g]
        assert out.tobytes() == want.tobytes()
        # placing into too small a buffer is refused
        with pytest.raises(codec_mod.Dct3dError):
            ctxs[-1].encode_u8_place(offs[world - 1], True, out[: out.size - 1])
        for g, c in enumerate(ctxs):
            lo, hi = sh.slab_range(nslabs, g, world)
            if hi == lo:
                continue
            for hint in (0, offs[g + 1]):
                fr, end = c.decode_u8_range(out, offs[g], (hi - lo) * 8, hint)
                assert end == offs[g + 1] and (fr == wdec[lo * 8:hi * 8]).all()
    finally:
        for c in ctxs:
            c.close()


def test_place_without_range_is_an_error(codec_mod):
    with codec_mod.Codec(64, 64, 8) as c:
        with pytest.raises(codec_mod.Dct3dError):
            c.encode_u8_place(0, True, np.zeros(64, np.uint8))


@pytest.mark.parametrize("ndev,cube", [(1, 8), (2, 8), (3, 8), (4, 4), (8, 8)])
def test_multi_object_equals_single_gpu(codec_mod, synth, ndev, cube):
    """dct3d_multi_* with `ndev` contexts (all on GPU 0 here; the 8-GPU box runs the same code with 8 ordinals): one
    stream, bit-identical to the single-GPU call; decode with the encoder's side information and without any."""
    W, H, F = 128, 64, 48 if cube == 8 else 24
    clip = synth.natural(W, H, F, 31)
    want, wbits, wdec = one_shot(codec_mod, clip, W, H, cube)
    with codec_mod.MultiCodec(W, H, cube, devices=[0] * ndev) as m:
        m.set_option("chunk_frames", cube)
        stream, nbits, starts = m.encode_u8(clip)
        assert nbits == wbits and stream.tobytes() == want.tobytes()
        assert starts[0] == 0 and starts[-1] == nbits and all(a <= b for a, b in zip(starts, starts[1:]))
        assert (m.decode_u8(stream, F, starts) == wdec).all()
        assert (m.decode_u8(stream, F) == wdec).all()                       # index discovery
        if ndev > 1:
            found = m.locate(stream, F)
            assert found[:ndev] == starts[:ndev]
        with pytest.raises(codec_mod.Dct3dError):
            m.decode_u8(stream[: stream.size // 2], F)


def test_multi_more_gpus_than_slabs(codec_mod, synth):
    W, H, F = 64, 32, 16
    clip = synth.natural(W, H, F, 5)
    want, wbits, wdec = one_shot(codec_mod, clip, W, H, 8)
    with codec_mod.MultiCodec(W, H, 8, devices=[0] * 5) as m:
        stream, nbits, starts = m.encode_u8(clip)
        assert nbits == wbits and stream.tobytes() == want.tobytes()
        assert (m.decode_u8(stream, F, starts) == wdec).all()
        assert (m.decode_u8(stream, F) == wdec).all()


def test_truncated_stream_is_reported_at_every_cut(codec_mod, synth):
    """a stream cut inside the last code must not decode.  Noise content makes the final (7,7,7) coefficients
    non-zero, so the last code is long; every cut of the last bytes is tried, for the one-shot decoder, the stage decoder
    and the streaming decoder (which must ask for more instead)."""
    lib = pkg("_lib")
    W, H, F = 32, 16, 8
    for seed in range(2, 8):
        clip = synth.noise(W, H, F, seed)
        with codec_mod.Codec(W, H, 8) as c:
            stream, nbits = c.encode_u8(clip, cap=4 * clip.size + 4096)
            q = c.quantize_u8(clip)
            full = c.decode_u8(stream, F)
            last_code_byte = (nbits - 1) // 8                 # byte that holds the last bit of the last code
            for cut in range(max(1, last_code_byte - 2), stream.size + 1):
                part = np.ascontiguousarray(stream[:cut])
                ok = cut > last_code_byte
                if ok:
                    assert (c.decode_u8(part, F) == full).all()
                    qd, end = c.eg_decode_i16(part, q.shape[0])
                    assert end == nbits and (qd == q).all()
                    res = c.stream_decode(part, 0, F)
                    assert res is not None and res[1] == nbits
                else:
                    with pytest.raises(codec_mod.Dct3dError) as e:
                        c.decode_u8(part, F)
                    assert e.value.code == lib.E_STREAM
                    with pytest.raises(codec_mod.Dct3dError) as e:
                        c.eg_decode_i16(part, q.shape[0])
                    assert e.value.code == lib.E_STREAM
                    assert c.stream_decode(part, 0, F) is None


def test_out_of_range_code_numbers_are_malformed(codec_mod, oracle):
    """17-bit code numbers other than 65537 (v = -32768) do not fit the codec's int16 cubes: E_STREAM, not a wrapped
    value.  m = 65537 itself decodes."""
    lib = pkg("_lib")

    def stream_with(m):
        bits = "1" * 100 + "0" * 16 + format(m, "017b") + "1" * (512 - 101)
        bits += "0" * (-len(bits) % 8)
        return np.frombuffer(int(bits, 2).to_bytes(len(bits) // 8, "big") + b"\0", np.uint8).copy()

    with codec_mod.Codec(8, 8, 8) as c:
        q, end = c.eg_decode_i16(stream_with(65537), 1)
        assert end == 100 + 33 + 411 and int(q.reshape(-1)[np.flatnonzero(q.reshape(-1))[0]]) == -32768
        for m in (65536, 65538, 0x1FFFF):
            with pytest.raises(codec_mod.Dct3dError) as e:
                c.eg_decode_i16(stream_with(m), 1)
            assert e.value.code == lib.E_STREAM


def test_misaligned_device_pointers_are_rejected(codec_mod):
    import torch
    lib = pkg("_lib")
    dev = torch.device("cuda", 0)
    W, H, F = 64, 32, 8
    with codec_mod.Codec(W, H, 8) as c:
        frames = torch.zeros(W * H * F + 64, dtype=torch.uint8, device=dev)
        q = torch.zeros(W * H * F + 64, dtype=torch.int16, device=dev)
        stream = torch.zeros(W * H * F, dtype=torch.uint8, device=dev)
        for call in (lambda: c.encode_u8_dev(frames[4:], F, stream, stream.numel()),
                     lambda: c.quantize_u8_dev(frames, F, q[1:]),
                     lambda: c.reconstruct_i16_dev(q[2:], F, frames),
                     lambda: c.reconstruct_i16_dev(q, F, frames[4:]),
                     lambda: c.decode_u8_dev(stream, 64, F, frames[2:])):
            with pytest.raises(codec_mod.Dct3dError) as e:
                call()
            assert e.value.code == lib.E_INVALID
        torch.cuda.synchronize()
        # an 8-byte aligned (not 16) frame buffer is fine: plain loads instead of TMA
        c.encode_u8_dev(frames[8:], F, stream, stream.numel())
        torch.cuda.synchronize()


def test_contexts_created_concurrently(codec_mod, synth):
    """One context per host thread, created and first used at the same moment (the lazy globals)."""
    import threading
    W, H, F = 64, 32, 16
    clip = synth.natural(W, H, F, 3)
    out, errs = [None] * 8, []

    def work(i):
        try:
            with codec_mod.Codec(W, H, 8) as c:
                out[i] = c.encode_u8(clip)[0].tobytes()
        except Exception as ex:   # noqa: BLE001
            errs.append(ex)

    th = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs and all(o == out[0] for o in out)


def test_multi_object_on_every_visible_gpu(codec_mod, synth):
    """The same on real distinct GPUs when the box has more than one (skipped on a single-GPU box)."""
    n = pkg("_lib").load().dct3d_device_count()
    if n < 2:
        pytest.skip("one GPU visible")
    W, H, F = 256, 128, 8 * 3 * n + 8
    clip = synth.natural(W, H, F, 41)
    want, wbits, wdec = one_shot(codec_mod, clip, W, H, 8)
    with codec_mod.MultiCodec(W, H, 8, devices=list(range(n))) as m:
        m.set_option("chunk_frames", 8)
        stream, nbits, starts = m.encode_u8(clip)
        assert nbits == wbits and stream.tobytes() == want.tobytes()
        assert (m.decode_u8(stream, F, starts) == wdec).all()
        assert (m.decode_u8(stream, F) == wdec).all()


@pytest.mark.parametrize("batch,chunk", [(1, 8), (3, 8), (2, 16), (5, 0)])
def test_streaming_in_batches_of_slabs(codec_mod, synth, batch, chunk):
    """dct3d_stream_encode / _decode with several slabs per call (the C codec's batch loop, host/encoder.c): the carried
    partial byte continues through the chunk pipeline (device-chained start bit), the concatenated output is the
    one-shot stream, and the streaming decoder returns NEED_MORE until a batch is complete."""
    W, H, F = 128, 64, 88                                   # 11 slabs: the last batch is short
    clip = synth.natural(W, H, F, 17)
    want, wbits, wdec = one_shot(codec_mod, clip, W, H, 8)
    with codec_mod.Codec(W, H, 8) as c:
        c.set_option("chunk_frames", chunk)
        c.stream_begin()
        parts = []
        for s in range(0, 11, batch):
            n = min(batch, 11 - s)
            parts.append(c.stream_encode(clip[s * 8:(s + n) * 8], last=(s + n == 11)))
        got = np.concatenate(parts)
        assert got.tobytes() == want.tobytes()
        # decode side: feed the stream in pieces of 1000 bytes
        buf, pos, done, frames = np.zeros(0, np.uint8), 0, 0, []
        fed = 0
        while done < 11:
            n = min(batch, 11 - done)
            res = c.stream_decode(buf, pos, n * 8) if buf.size else None
            if res is None:
                assert fed < got.size, "decoder asks for more than the stream holds"
                buf = np.concatenate([buf, got[fed:fed + 1000]])
                fed += 1000
                continue
            out, pos = res
            frames.append(out)
            buf = buf[pos // 8:]
            pos %= 8
            done += n
        assert (np.concatenate(frames) == wdec).all()


@pytest.mark.parametrize("ndev", [2, 3, 8])
@pytest.mark.parametrize("kind", ["natural", "noise"])
def test_distributed_index_discovery(codec_mod, synth, ndev, kind):
    """dct3d_multi_locate on a stream large enough for the distributed path (every GPU counts the codes of 1/G of the
    bytes, the host checks the entry guesses and locates the ranges): the start bits are the encoder's, and a decode
    without side information gives the one-shot frames."""
    W, H, F = 640, 480, 64 if kind == "natural" else 32
    clip = synth.natural(W, H, F, 13) if kind == "natural" else synth.noise(W, H, F, 14)
    with codec_mod.Codec(W, H, 8) as c:
        stream, nbits = c.encode_u8(clip, cap=4 * clip.size + 4096)
        wdec = c.decode_u8(stream, F)
        cps = (W // 8) * (H // 8)
    assert stream.size * 8 >= ndev * (1 << 19)                   # above the small-stream fallback
    with codec_mod.MultiCodec(W, H, 8, devices=[0] * ndev) as m:
        _, nb2, starts = m.encode_u8(clip, cap=4 * clip.size + 4096)
        assert nb2 == nbits
        found = m.locate(stream, F)
        assert found[:ndev] == starts[:ndev]
        assert (m.decode_u8(stream, F) == wdec).all()
        # truncated in the middle of the last range: an error, not garbage
        with pytest.raises(codec_mod.Dct3dError):
            m.decode_u8(stream[: stream.size - 1000], F)
        # truncated before the last range starts
        with pytest.raises(codec_mod.Dct3dError):
            m.decode_u8(stream[: starts[ndev - 1] // 8 - 10], F)


def test_distributed_discovery_with_wrong_entry_guesses(codec_mod):
    """A stream of equal 33-bit codes never resynchronises: the entry guess of every part is wrong, so every part is
    recounted from its predecessor's verified overhang (and every segment inside it needs its own fix-up round)."""
    W = H = 64
    nslabs, cps = 8, 64
    q = np.full((nslabs * cps, 8, 8, 8), -32768, np.int16)
    with codec_mod.Codec(W, H, 8) as c:
        stream, end = c.eg_encode_i16(q)
    assert end == q.size * 33
    sh = pkg("sharding")
    with codec_mod.MultiCodec(W, H, 8, devices=[0] * 4) as m:
        found = m.locate(stream, nslabs * 8)
        for g in range(4):
            lo, _ = sh.slab_range(nslabs, g, 4)
            assert found[g] == lo * cps * 512 * 33


@pytest.mark.parametrize("ndev,batch", [(2, 4), (3, 5)])
def test_multi_streaming_in_batches(codec_mod, synth, ndev, batch):
    """dct3d_multi_stream_encode / _decode: batches of slabs shared out over `ndev` contexts with the partial byte carried
    between calls give the one-shot stream; the decoder discovers the range boundaries of every batch and asks for more
    input until a batch is complete."""
    W, H, nsl = 128, 64, 11
    clip = synth.natural(W, H, nsl * 8, 19)
    want, wbits, wdec = one_shot(codec_mod, clip, W, H, 8)
    with codec_mod.MultiCodec(W, H, 8, devices=[0] * ndev) as m:
        m.set_option("chunk_frames", 8)
        m.stream_begin()
        parts = []
        for s in range(0, nsl, batch):
            n = min(batch, nsl - s)
            parts.append(m.stream_encode(clip[s * 8:(s + n) * 8], last=(s + n == nsl)))
        got = np.concatenate(parts)
        assert got.tobytes() == want.tobytes()
        buf, pos, done, frames, fed = np.zeros(0, np.uint8), 0, 0, [], 0
        while done < nsl:
            n = min(batch, nsl - done)
            res = m.stream_decode(buf, pos, n * 8) if buf.size else None
            if res is None:
                assert fed < got.size, "decoder asks for more than the stream holds"
                buf = np.concatenate([buf, got[fed:fed + 3000]])
                fed += 3000
                continue
            out, pos = res
            frames.append(out)
            buf = buf[pos // 8:]
            pos %= 8
            done += n
        assert (np.concatenate(frames) == wdec).all()


def test_multi_weighted_shares_give_the_same_stream(codec_mod, synth):
    """Slab ranges that follow per-GPU weights (dct3d_multi_set_weights; the shares a link probe proposes on hosts whose GPUs
    do not get equal host bandwidth): any shares, including zero for a GPU, give the one-shot stream and frames."""
    W, H, F = 128, 64, 88
    clip = synth.natural(W, H, F, 29)
    want, wbits, wdec = one_shot(codec_mod, clip, W, H, 8)
    with codec_mod.MultiCodec(W, H, 8, devices=[0] * 4) as m:
        for weights in ([1, 2, 3, 4], [5, 1, 1, 0.2], [0, 1, 0, 1], [1, 0, 0, 0], None):
            m.set_weights(weights)
            stream, nbits, starts = m.encode_u8(clip)
            assert nbits == wbits and stream.tobytes() == want.tobytes(), weights
            assert (m.decode_u8(stream, F, starts) == wdec).all(), weights
            assert (m.decode_u8(stream, F) == wdec).all(), weights
        up, down, w = m.probe_links()
        assert len(w) == 4 and all(x > 0 for x in up + down + w)
        m.set_weights(w)
        stream, nbits, _ = m.encode_u8(clip)
        assert stream.tobytes() == want.tobytes()
        with pytest.raises(codec_mod.Dct3dError):
            m.set_weights([0, 0, 0, 0])


@pytest.mark.parametrize("W,H,F,cube,kind", [(256, 64, 24, 8, "natural"), (1920, 136, 16, 8, "natural"), (128, 64, 16, 8, "noise"),
                                             (96, 32, 16, 8, "constant"), (200, 24, 16, 8, "natural"), (128, 32, 8, 4, "natural"),
                                             (64, 64, 16, 8, "sparse")])
def test_kernel_variants_are_bit_identical(codec_mod, oracle, synth, W, H, F, cube, kind):
    """Round-2 kernel variants -- the encoder's zero-run skip (option zero_skip), the inverse kernel's TMA tile store
    (tma_store) and the packer's sorted deal (pack_sort) -- against the plain paths: same stream, same frames, and the
    frames within +-1 of the oracle.  Widths that are not a multiple of 32 must fall back to row stores by themselves."""
    if kind == "sparse":                       # mostly flat cubes with a few textured ones: all three column classes occur
        rng = np.random.default_rng(3)
        clip = np.full((F, H, W), 120, np.uint8)
        clip[:, :16, :24] = rng.integers(0, 256, size=(F, 16, 24), dtype=np.uint8)
        clip[:, 32:40, :] = (np.arange(W)[None, None, :] * 3 % 256).astype(np.uint8)
        clip[:, 48:, 32:] = synth.natural(W - 32, H - 48, F, 5)
    else:
        clip = getattr(synth, kind)(W, H, F, 9) if kind != "constant" else synth.constant(W, H, F, 77)
    outs = {}
    for skip, tma, psort in [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 1)]:
        with codec_mod.Codec(W, H, cube) as c:
            c.set_option("zero_skip", skip)
            c.set_option("tma_store", tma)
            c.set_option("pack_sort", psort)
            stream, nbits = c.encode_u8(clip)
            dec = c.decode_u8(stream, F)
            assert c.stat("tma_store_used") == (1 if tma and W % 32 == 0 else 0)
            outs[(skip, tma, psort)] = (stream.tobytes(), nbits, dec)
    base = outs[(0, 0, 0)]
    for key, (s, n, d) in outs.items():
        assert n == base[1] and s == base[0], f"stream differs for variant {key}"
        assert (d == base[2]).all(), f"frames differ for variant {key}: {int((d != base[2]).sum())} pixels"
    odec = oracle.decode_u8(np.frombuffer(base[0], np.uint8), W, H, F // cube * cube, cube)
    assert np.abs(base[2][: F // cube * cube].astype(int) - odec.astype(int)).max() <= 1
