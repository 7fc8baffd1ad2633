"""GPU parity tests (run with -m gpu on the B200 box): every call goes through the C ABI of
libdct3d.so (ctypes), and is compared with the CPU oracle on the same seeded inputs.

Parity rules (BASELINE.json north_star):
  (1) float DCT coefficients within 1e-4 relative (to the block's magnitude) of the fp64 oracle;
  (2) quantised cubes identical except +-1 flips at rounding near-ties, counted and bounded;
  (3) Exp-Golomb stream bit-exact given identical quantised cubes;
  (4) decoded pixels within +-1.
"""
import numpy as np
import pytest

from conftest import pkg

pytestmark = pytest.mark.gpu

FLIP_RATE_MAX = 2e-5   # fp32 near-tie flips; SURVEY.md App. D measured 1.4e-6


@pytest.fixture(scope="module")
def codec_mod():
    return pkg("codec")


def make(codec_mod, W, H, cube=8, tma=None):
    c = codec_mod.Codec(W, H, cube)
    if tma is not None:
        c.set_option("tma", tma)
    return c


CLIPS = [
    # name, W, H, F, cube, generator, seed
    ("golden", 64, 48, 24, 8, "natural", 1),
    ("wide", 384, 16, 8, 8, "natural", 3),
    ("ragged", 200, 24, 16, 8, "natural", 4),      # 25 cubes per row: partial boxes, W % 16 != 0 -> plain loads
    ("noise", 128, 32, 8, 8, "noise", 2),
    ("const", 32, 16, 8, 8, "constant", 0),
    ("tail", 32, 32, 19, 8, "natural", 5),         # 3 trailing frames dropped
    ("c4", 128, 16, 8, 4, "natural", 6),
    ("c4ragged", 40, 12, 12, 4, "noise", 7),
]


def gen(synth, kind, W, H, F, seed):
    if kind == "natural":
        return synth.natural(W, H, F, seed)
    if kind == "noise":
        return synth.noise(W, H, F, seed)
    return synth.constant(W, H, F, 128)


def test_baseline_config1_in_full_against_oracle(codec_mod, oracle, synth):
    """BASELINE configs[0] (640x480, 64 frames, 8^3) in full: all four parity rules on every slab, and the whole clip's
    stream is the concatenation of the oracle's coding of the GPU's cubes (J/Encoder.java:14-129)."""
    from oracle import parity as PAR
    W, H, F = 640, 480, 64
    clip = synth.natural_slabs(W, H, 8, 0, F // 8, 1)
    with make(codec_mod, W, H, 8) as c:
        parts = [PAR.piece_parity(c, clip[s * 8:(s + 1) * 8], 8) for s in range(F // 8)]
        m = PAR.merge(parts)
        print("config 1 parity:", m)
        assert m["coef_max_rel"] <= 1e-4 and m["stream_bit_exact"] and m["pixel_max_abs"] <= 1
        assert m["flips_exact"] == 0 and m["flips_near"] <= max(2, FLIP_RATE_MAX * m["coefficients"])
        stream, nbits = c.encode_u8(clip)
        q = c.quantize_u8(clip).astype(np.int32)
        want, want_bits = oracle.eg_encode_cubes(q, 8, cap=5 * q.size + 64)
        assert nbits == want_bits == sum(p["nbits"] for p in parts) and stream.tobytes() == want[: nbits // 8 + 1].tobytes()
        dec = c.decode_u8(stream, F)
        odec = oracle.decode_u8(stream, W, H, F, 8)
        assert np.abs(dec.astype(int) - odec.astype(int)).max() <= 1


@pytest.mark.parametrize("cube,F", [(8, 256), (4, 256)])
def test_baseline_1080p_sampled_slabs_against_oracle(codec_mod, oracle, synth, cube, F):
    """BASELINE configs[1] (1920x1080x256, 8^3) and configs[3] (4^3): the whole clip goes through the fused calls; four
    sampled slabs (first, two in the middle, last) are checked against the oracle under all four rules, and the bits
    and pixels of those slabs inside the whole clip's stream are the ones the slab gives on its own (slabs are
    independent key-frame groups, only the bit position is shared: C/encoder.c:203-278)."""
    from oracle import parity as PAR
    sh = pkg("sharding")
    W, H = 1920, 1080
    if cube == 4:
        H = 1080 - 1080 % 4
    nslabs = F // cube
    picks = sorted({0, nslabs // 3, (2 * nslabs) // 3, nslabs - 1})
    cps = (W // cube) * (H // cube)
    with make(codec_mod, W, H, cube) as c:
        clip = synth.natural_slabs(W, H, cube, 0, nslabs, 1)
        stream, nbits = c.encode_u8(clip)
        dec = c.decode_u8(stream, F)
        parts = []
        for s in picks:
            piece = clip[s * cube:(s + 1) * cube]
            p = PAR.piece_parity(c, piece, cube)
            parts.append(p)
            # where the slab sits in the whole clip's stream
            b0 = c.eg_locate(stream, s * cps) if s else 0
            b1 = c.eg_locate(stream, (s + 1) * cps)
            assert b1 - b0 == p["nbits"]
            own, own_bits = c.encode_u8(piece)
            moved = sh.shift_to_phase(own, own_bits, b0 % 8)
            got = stream[b0 // 8: b0 // 8 + moved.size].copy()
            got[0] &= 0xFF >> (b0 % 8)
            if (b0 % 8 + own_bits) % 8:
                got[-1] &= (0xFF << (8 - (b0 % 8 + own_bits) % 8)) & 0xFF
            else:
                got = got[:-1]
                moved = moved[:-1]
            assert got.tobytes() == moved.tobytes()
            assert (dec[s * cube:(s + 1) * cube] == c.decode_u8(own, cube)).all()
        m = PAR.merge(parts)
        print(f"1080p x{F} cube {cube} parity on slabs {picks}:", m)
        assert m["coef_max_rel"] <= 1e-4 and m["stream_bit_exact"] and m["pixel_max_abs"] <= 1
        assert m["flips_near"] <= max(2, FLIP_RATE_MAX * m["coefficients"])
        # "exact" = |frac - 0.5| < 1e-9, where the fp64 oracle itself cannot decide; at 66 M coefficients one or two such
        # values occur by chance even for the irrational 8^3 basis (the small clips above assert none)
        if cube == 8:
            assert m["flips_exact"] <= 3


def classify_flips(q, ref, coef_planar, oracle, cube):
    """Every mismatch must be a +-1 flip at a rounding tie of the fp64 value: an EXACT tie (the 4^3
    transform has rational coefficients, e.g. DC = sum/8; the reference's own Java and C results differ
    there) or a NEAR tie within fp32 error.  Returns (exact, near)."""
    bad = np.argwhere(q != ref)
    if bad.size == 0:
        return 0, 0
    assert np.abs(q - ref).max() <= 1
    cc = oracle.frames_to_cubes(coef_planar, cube)
    k = np.indices((cube, cube, cube)).sum(axis=0)
    div = np.maximum(1, 5 * k)[None]
    v = (cc / div)[tuple(bad.T)]
    dist = np.abs(np.abs(v - np.floor(v)) - 0.5)
    assert dist.max() < 2e-3, "a quantised value differs from the oracle away from any rounding tie"
    exact = int((dist < 1e-9).sum())
    return exact, int(bad.shape[0] - exact)


@pytest.mark.parametrize("name,W,H,F,cube,kind,seed", CLIPS)
@pytest.mark.parametrize("tma", [1, 0])
def test_quantised_cubes_match_oracle(codec_mod, oracle, synth, name, W, H, F, cube, kind, seed, tma):
    """Rule (2): identical quantised cubes except counted +-1 flips at rounding ties."""
    if tma and W % 16:
        pytest.skip("TMA needs 16-byte row pitch")
    clip = gen(synth, kind, W, H, F, seed)
    with make(codec_mod, W, H, cube, tma) as c:
        q = c.quantize_u8(clip).astype(np.int32)
    ref, coef = oracle.quantized_cubes(clip, cube, mode=0, want_coef=True)
    assert q.shape == ref.shape
    exact, near = classify_flips(q, ref, coef, oracle, cube)
    print(f"[{name} tma={tma}] flips vs fp64 oracle: {exact} at exact ties, {near} at near ties, of {q.size}")
    assert near <= max(2, FLIP_RATE_MAX * q.size)
    if cube == 8:
        assert exact == 0      # the 8^3 basis is irrational: no exact ties
        assert (ref == oracle.quantized_cubes(clip, cube, mode=1)).all()   # Java and C rounding agree


@pytest.mark.parametrize("cube", [8, 4])
def test_forward_inverse_f32_seam(codec_mod, oracle, synth, cube):
    W, H, F = 96, 32, 2 * cube
    clip = synth.natural(W, H, F, 11)
    cubes = oracle.frames_to_cubes(clip, cube).astype(np.float32)
    with make(codec_mod, W, H, cube) as c:
        coef = c.forward_f32(cubes).reshape(cubes.shape)
        ref = oracle.frames_to_cubes(oracle.dct_sep(clip.astype(np.float64), cube), cube)
        blockmax = np.abs(ref).reshape(ref.shape[0], -1).max(axis=1).reshape(-1, 1, 1, 1)
        rel = np.abs(coef - ref) / np.maximum(blockmax, 1.0)
        assert rel.max() <= 1e-4            # rule (1)
        assert np.abs(coef - ref).max() < 2e-3
        back = c.inverse_f32(coef).reshape(cubes.shape)
        assert np.abs(back - cubes).max() < 2e-3
        # clamp: coefficients of an out-of-range signal
        big = c.inverse_f32(c.forward_f32(cubes * 4.0 - 300.0))
        assert big.min() >= 0.0 and big.max() <= 255.0


@pytest.mark.parametrize("cube", [8, 4])
def test_forward_inverse_f64_seam(codec_mod, oracle, synth, cube):
    W, H, F = 64, 24, 2 * cube
    px = synth.noise(W, H, F, 12).astype(np.float64)
    with make(codec_mod, W, H, cube) as c:
        coef = c.forward_f64(px)
        ref = oracle.dct_sep(px, cube)
        assert np.abs(coef - ref).max() <= 1e-10 * np.abs(ref).max()
        back = c.inverse_f64(coef)
        assert np.abs(back - px).max() < 1e-9
    # Java-shaped object API
    out = np.zeros_like(px)
    codec_mod.DCT(px, out, W, H, cube, cube, cube).run()
    assert np.abs(out - ref).max() <= 1e-10 * np.abs(ref).max()
    pix = np.zeros_like(px)
    codec_mod.InverseDCT(out, pix, W, H, cube, cube, cube).run()
    assert np.abs(pix - px).max() < 1e-9


def _rand_cubes(rng, ncubes, cs, density, hi):
    q = np.zeros((ncubes, cs), np.int16)
    m = rng.random((ncubes, cs)) < density
    q[m] = rng.integers(-hi - 1, hi + 1, size=int(m.sum())).astype(np.int16)
    return q


@pytest.mark.parametrize("cube", [8, 4])
@pytest.mark.parametrize("ncubes,density,hi,start", [(1, 0.03, 300, 0), (33, 0.0, 1, 3), (70, 0.05, 5770, 13),
                                                     (97, 1.0, 32767, 0), (1000, 0.03, 200, 7)])
def test_expgolomb_bit_exact(codec_mod, oracle, cube, ncubes, density, hi, start):
    """Rule (3): identical cubes -> identical bits, for any start phase, plus decode round trip."""
    cs = cube ** 3
    rng = np.random.default_rng(ncubes + cube)
    q = _rand_cubes(rng, ncubes, cs, density, hi).reshape(-1, cube, cube, cube)
    ref, ref_end = oracle.eg_encode_cubes(q.astype(np.int32), cube, start_bit=start)
    prefix = np.array([0xA0], np.uint8) if start else None     # pre-existing bits of the partial byte survive
    if start:
        ref = ref.copy()
        ref[0] |= 0xA0      # 101 then zeros: only bits before start_bit (>= 3) are set
    with make(codec_mod, 8 * cube, 8 * cube, cube) as c:
        out, end = c.eg_encode_i16(q, start, prefix)
        assert end == ref_end
        assert out.tobytes() == ref.tobytes()
        back, dend = c.eg_decode_i16(out, ncubes, start)
        assert dend == ref_end
        assert (back == q).all()


def test_reference_golden_stream(codec_mod, oracle, golden):
    """The reference CLI's own stream and pixels (tests/golden/make_golden.py)."""
    stream, refq = golden["flow_stream"], golden["flow_qcubes"]
    with make(codec_mod, 64, 48, 8) as c:
        q, end = c.eg_decode_i16(stream, refq.shape[0])
        assert end == int(golden["flow_bits"]) == 84744
        assert (q == refq).all()
        out, oend = c.eg_encode_i16(refq)
        assert oend == 84744 and out.tobytes() == stream.tobytes()
        dec = c.decode_u8(stream, 24)
        assert np.abs(dec.astype(int) - golden["flow_decoded"].astype(int)).max() <= 1   # rule (4)
        assert (dec != golden["flow_decoded"]).mean() < 0.01


@pytest.mark.parametrize("name,W,H,F,cube,kind,seed", CLIPS)
def test_fused_encode_decode(codec_mod, oracle, synth, name, W, H, F, cube, kind, seed):
    clip = gen(synth, kind, W, H, F, seed)
    Fe = F - F % cube
    with make(codec_mod, W, H, cube) as c:
        q = c.quantize_u8(clip)
        stream, nbits = c.encode_u8(clip)
        # the fused kernel's stream is exactly the Exp-Golomb coding of the cubes it quantised
        ref, ref_bits = oracle.eg_encode_cubes(q.astype(np.int32), cube)
        assert nbits == ref_bits and stream.size == nbits // 8 + 1
        assert stream.tobytes() == ref.tobytes()
        oq = oracle.quantized_cubes(clip, cube, 0)
        if (oq == q).all():
            full, fbits = oracle.encode_u8(clip, cube, 0)
            assert fbits == nbits and full.tobytes() == stream.tobytes()
        dec = c.decode_u8(stream, F)
        assert dec.shape == (Fe, H, W)
        odec = oracle.decode_u8(stream, W, H, F, cube)
        d = np.abs(dec.astype(int) - odec.astype(int))
        print(f"[{name}] decoded pixels differing from oracle: {(d != 0).mean():.2e}")
        assert d.max(initial=0) <= 1                                        # rule (4)
        rec = c.reconstruct_i16(q, F)
        assert (rec == dec).all()


def test_streaming_equals_one_shot(codec_mod, oracle, synth):
    W, H, F = 64, 48, 24
    clip = synth.natural(W, H, F, 1)
    with make(codec_mod, W, H, 8) as c:
        one, nbits = c.encode_u8(clip)
        c.stream_begin()
        parts = [c.stream_encode(clip[i:i + 8], last=(i + 8 >= F)) for i in range(0, F, 8)]
        cat = np.concatenate(parts)
        assert cat.tobytes() == one.tobytes()     # C/encoder.c:263-271 slab loop == Java one-shot
        # slab-by-slab decode with a carried bit position
        pos, frames, buf = 0, [], one
        for i in range(0, F, 8):
            res = c.stream_decode(buf, pos, 8)
            assert res is not None
            fr, pos = res
            frames.append(fr)
            buf, pos = buf[pos // 8:], pos % 8
        assert (np.concatenate(frames) == c.decode_u8(one, F)).all()
        assert c.stream_decode(one[: one.size // 2], 0, F) is None   # not enough input yet


def test_error_behaviour(codec_mod, synth):
    with pytest.raises(codec_mod.Dct3dError):
        codec_mod.Codec(100, 48, 8)            # width not a multiple of the cube edge
    with pytest.raises(codec_mod.Dct3dError):
        codec_mod.Codec(64, 48, 5)
    with pytest.raises(codec_mod.Dct3dError):
        codec_mod.Codec(64, 48, 8, device=99)
    clip = synth.noise(64, 48, 8, 1)
    with make(codec_mod, 64, 48, 8) as c:
        L = c.L
        out = np.zeros(64, np.uint8)
        import ctypes as C
        nb, ny = C.c_uint64(), C.c_size_t()
        rc = L.dct3d_encode_u8(c.h, clip.ctypes.data, 8, out.ctypes.data, out.size, C.byref(nb), C.byref(ny))
        assert rc == -3 and b"small" in L.dct3d_last_error(c.h)            # DCT3D_E_OVERFLOW
        stream, _ = c.encode_u8(clip)
        with pytest.raises(codec_mod.Dct3dError) as e:
            c.decode_u8(stream[: stream.size // 3], 8)
        assert e.value.code == -4                                           # DCT3D_E_STREAM
        with pytest.raises(codec_mod.Dct3dError):
            c.decode_u8(np.zeros(4096, np.uint8), 8)                        # endless zero prefix: malformed
        empty, bits = c.encode_u8(np.zeros((3, 48, 64), np.uint8))          # fewer frames than a cube
        assert bits == 0 and empty.size == 1


def test_full_hd_slabs_properties(codec_mod, oracle, synth):
    """1080p, 16 frames: size-independent properties + oracle spot check on the first cube row."""
    W, H, F = 1920, 1080, 16
    rng = np.random.default_rng(21)
    base = synth.natural(W, 136, F, 8)
    clip = np.tile(base, (1, 8, 1))[:, :H, :].copy()
    clip[:, 500:508, :] = rng.integers(0, 256, size=(F, 8, W), dtype=np.uint8)
    with make(codec_mod, W, H, 8) as c:
        stream, nbits = c.encode_u8(clip)
        q = c.quantize_u8(clip)
        ref, ref_bits = oracle.eg_encode_cubes(q.astype(np.int32), 8)
        assert ref_bits == nbits and ref.tobytes() == stream.tobytes()
        qd, end = c.eg_decode_i16(stream, q.shape[0])
        assert end == nbits and (qd == q).all()
        dec = c.decode_u8(stream, F)
        assert (dec == c.reconstruct_i16(q, F)).all()
        # oracle on the first row of cubes of both slabs
        oq = oracle.quantized_cubes(clip[:, :8, :], 8, 0)
        mine = q.reshape(2, 135, 240, 512)[:, 0].reshape(-1, 8, 8, 8)
        assert np.abs(mine.astype(int) - oq).max() <= 1 and (mine != oq).sum() <= 3
        # TMA and plain loads agree bit for bit
        c.set_option("tma", 0)
        s2, b2 = c.encode_u8(clip)
        assert b2 == nbits and s2.tobytes() == stream.tobytes()


def test_c_cli_interoperates_with_reference_cli(codec_mod, oracle, synth, tmp_path):
    """Files written by our C codec (host/codec, the reference's encoder.c/decoder.c flow over libdct3d)
    decode with the reference's own CLI (oracle/_ref/codec_ref: unmodified reference C + CPU OpenCL shim)
    and vice versa; the Python mirrors of the Java/C command lines read and write the same files."""
    import os, subprocess, zlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ours = os.path.join(root, "host", "codec")
    ref = os.path.join(root, "oracle", "_ref", "codec_ref")
    if not os.path.exists(ours):
        subprocess.check_call(["make", "-C", os.path.join(root, "host"), "CC=gcc"])
    W, H, F = 64, 48, 24
    clip = synth.natural(W, H, F, 1)
    raw = tmp_path / "in.raw"
    clip.tofile(raw)
    args = [str(W), str(H), str(F)]
    run = lambda exe, *a: subprocess.run([exe, *a], check=True, stdout=subprocess.DEVNULL)
    run(ours, "encode", str(raw), str(tmp_path / "ours.enc"), *args)
    stream = np.frombuffer(zlib.decompress(open(tmp_path / "ours.enc", "rb").read()), np.uint8)
    with make(codec_mod, W, H, 8) as c:
        one, nbits = c.encode_u8(clip)
    assert stream.tobytes() == one.tobytes()                       # slab loop + carried partial byte == one shot
    run(ours, "decode", str(tmp_path / "ours.enc"), str(tmp_path / "ours.dec"), *args)
    mine = np.fromfile(tmp_path / "ours.dec", np.uint8).reshape(F, H, W)
    assert np.abs(mine.astype(int) - oracle.decode_u8(stream, W, H, F).astype(int)).max() <= 1
    # Python mirrors of the Java command lines and of the C command line
    assert codec_mod.Encoder.main([str(raw), str(tmp_path / "j.enc"), *args]) == 0
    assert zlib.decompress(open(tmp_path / "j.enc", "rb").read()) == one.tobytes()
    assert codec_mod.Decoder.main([str(tmp_path / "ours.enc"), str(tmp_path / "j.dec"), *args]) == 0
    assert (np.fromfile(tmp_path / "j.dec", np.uint8).reshape(F, H, W) == mine).all()
    assert codec_mod.codec_main(["codec", "decode", str(tmp_path / "j.enc"), str(tmp_path / "p.dec"), *args]) == 0
    assert (np.fromfile(tmp_path / "p.dec", np.uint8).reshape(F, H, W) == mine).all()
    if os.path.exists(ref):
        run(ref, "decode", str(tmp_path / "ours.enc"), str(tmp_path / "ref.dec"), *args)      # reference reads our file
        theirs = np.fromfile(tmp_path / "ref.dec", np.uint8).reshape(F, H, W)
        assert np.abs(theirs.astype(int) - mine.astype(int)).max() <= 1                          # rule (4)
        run(ref, "encode", str(raw), str(tmp_path / "ref.enc"), *args)                          # we read the reference's file
        run(ours, "decode", str(tmp_path / "ref.enc"), str(tmp_path / "x.dec"), *args)
        run(ref, "decode", str(tmp_path / "ref.enc"), str(tmp_path / "y.dec"), *args)
        a = np.fromfile(tmp_path / "x.dec", np.uint8).astype(int)
        b = np.fromfile(tmp_path / "y.dec", np.uint8).astype(int)
        assert np.abs(a - b).max() <= 1


def test_zero_source_change_mode_runs_the_reference_on_the_gpu(codec_mod, oracle, synth, tmp_path):
    """the reference's own main.c / encoder.c / decoder.c / ExpGolomb.c / CubeUtils.c, compiled unmodified
    against host/clshim (host/_refgpu/codec_ref_gpu: its cl* calls are served by dct3d_forward_f32 / dct3d_inverse_f32),
    produce files that our C codec and the CPU-shimmed reference decode to the same frames (+-1), and decode our files."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gpu_ref = os.path.join(root, "host", "_refgpu", "codec_ref_gpu")
    ours = os.path.join(root, "host", "codec")
    cpu_ref = os.path.join(root, "oracle", "_ref", "codec_ref")
    if not os.path.exists(gpu_ref):
        pytest.skip("host/_refgpu/codec_ref_gpu is built only where the reference checkout exists")
    W, H, F = 64, 48, 24
    clip = synth.natural(W, H, F, 1)
    raw = tmp_path / "in.raw"
    clip.tofile(raw)
    args = [str(W), str(H), str(F)]
    run = lambda exe, *a: subprocess.run([exe, *a], check=True, stdout=subprocess.DEVNULL, cwd=str(tmp_path))
    read = lambda name: np.fromfile(tmp_path / name, np.uint8).reshape(F, H, W).astype(int)
    assert b"B200" in subprocess.run([gpu_ref, "list_platforms"], capture_output=True, check=True).stdout
    run(gpu_ref, "encode", str(raw), "g.enc", *args)
    run(gpu_ref, "decode", "g.enc", "gg.dec", *args)
    run(ours, "decode", str(tmp_path / "g.enc"), str(tmp_path / "go.dec"), *args)
    assert np.abs(read("gg.dec") - read("go.dec")).max() <= 1                  # rule (4), same file, two decoders
    assert np.abs(read("gg.dec") - clip.astype(int)).mean() < 6.0              # and it is the clip
    run(ours, "encode", str(raw), str(tmp_path / "o.enc"), *args)
    run(gpu_ref, "decode", "o.enc", "og.dec", *args)
    run(ours, "decode", str(tmp_path / "o.enc"), str(tmp_path / "oo.dec"), *args)
    assert np.abs(read("og.dec") - read("oo.dec")).max() <= 1
    if os.path.exists(cpu_ref):
        run(cpu_ref, "decode", "g.enc", "gc.dec", *args)
        assert np.abs(read("gg.dec") - read("gc.dec")).max() <= 1
        # the two encoders (GPU butterflies vs the CPU restatement of 3dDCT.cl) differ at most by rounding-tie flips
        run(cpu_ref, "encode", str(raw), "c.enc", *args)
        run(cpu_ref, "decode", "c.enc", "cc.dec", *args)
        assert (read("gg.dec") != read("cc.dec")).mean() < 0.02


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_ranges_equal_one_shot(codec_mod, synth, world):
    """The multi-GPU path, emulated on one GPU: every 'rank' codes its slab range from bit 0 of its own
    buffer; prefix-summed bit counts + shifted concatenation give the one-shot stream bit for bit, and
    every range decodes from its start bit."""
    sh = pkg("sharding")
    W, H, F = 128, 64, 64
    clip = synth.natural(W, H, F, 3)
    with make(codec_mod, W, H, 8) as c:
        one, bits = c.encode_u8(clip)
        parts, nb = [], []
        for g in range(world):
            lo, hi = sh.slab_range(F // 8, g, world)
            s, b = c.encode_u8(clip[lo * 8:hi * 8])
            parts.append(s)
            nb.append(b)
        cat, total = sh.concatenate(parts, nb)
        assert total == bits and cat.tobytes() == one.tobytes()
        offs = sh.bit_offsets(nb)
        assert any(o % 8 for o in offs[1:-1])          # ranges do not start on byte boundaries
        full = c.decode_u8(one, F)
        for g in range(world):
            lo, hi = sh.slab_range(F // 8, g, world)
            fr, end = c.stream_decode(one, offs[g], (hi - lo) * 8)
            assert end == offs[g + 1] and (fr == full[lo * 8:hi * 8]).all()


def _gpu_clip(W, H, F, seed):
    import torch
    bench = __import__("bench")
    return bench.synth_clip_torch(W, H, F, seed, torch.device("cuda", 0))


@pytest.mark.parametrize("W,H,F,cube", [(1920, 1080, 256, 8), (3840, 2160, 32, 8), (1920, 1080, 64, 4)])
def test_full_size_configs_properties(codec_mod, W, H, F, cube):
    """BASELINE configs at full frame size (config 2 in full; config 3's 4K frames and config 4's 4^3 cubes on
    a shorter clip): size-independent properties, all device-resident.
      * the fused stream equals the Exp-Golomb coding of the cubes the quantise entry point returns;
      * decode(encode(x)) equals reconstruct(quantise(x));
      * index discovery finds exactly the cubes that were coded (eg_decode == quantise);
      * slab-range sharding (4 ranges) concatenates to the one-shot stream;
      * the lossy round trip stays close to the source."""
    import torch
    sh = pkg("sharding")
    frames = _gpu_clip(W, H, F, 5)
    N = W * H * F
    cap = N // 2 + 4096
    dev = frames.device
    with make(codec_mod, W, H, cube) as c:
        d_stream = torch.zeros(cap, dtype=torch.uint8, device=dev)
        # the context runs on its own non-blocking stream (stream argument 0): torch's pending work on
        # these tensors (clip generation, zero fills) must be complete before each call
        torch.cuda.synchronize()
        end = c.encode_u8_dev(frames, F, d_stream, cap)
        nbytes = end // 8 + 1
        q = torch.empty(N, dtype=torch.int16, device=dev)
        c.quantize_u8_dev(frames, F, q)
        torch.cuda.synchronize()
        # stage path: cubes -> stream must be the same bits
        d2 = torch.zeros(cap, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        e2 = C_u64()
        c._check(c.L.dct3d_eg_encode_i16_dev(c.h, q.data_ptr(), N // cube ** 3, 0, d2.data_ptr(), cap, e2, 0))
        torch.cuda.synchronize()
        assert e2.value == end and torch.equal(d2[:nbytes], d_stream[:nbytes])
        # decode side
        out = torch.empty_like(frames)
        dend = c.decode_u8_dev(d_stream, nbytes, F, out)
        rec = torch.empty_like(frames)
        c.reconstruct_i16_dev(q, F, rec)
        q2 = torch.empty_like(q)
        e3 = C_u64()
        c._check(c.L.dct3d_eg_decode_i16_dev(c.h, d_stream.data_ptr(), nbytes, 0, N // cube ** 3, q2.data_ptr(), e3, 0))
        torch.cuda.synchronize()
        assert dend == end == e3.value
        assert torch.equal(out, rec) and torch.equal(q, q2)
        err = (out.to(torch.int16) - frames.to(torch.int16)).abs()
        assert float(err.float().mean()) < 6.0 and int(err.max()) < 80
        # sharding: 4 slab ranges from bit 0 each, concatenated on the host
        host_one = d_stream[:nbytes].cpu().numpy()
        parts, nb = [], []
        nsl = F // cube
        for g in range(4):
            lo, hi = sh.slab_range(nsl, g, 4)
            part = torch.zeros(cap, dtype=torch.uint8, device=dev)
            torch.cuda.synchronize()
            e = c.encode_u8_dev(frames[lo * cube:hi * cube], (hi - lo) * cube, part, cap)
            parts.append(part[: e // 8 + 1].cpu().numpy())
            nb.append(e)
        cat, total = sh.concatenate(parts, nb)
        assert total == end and cat.tobytes() == host_one.tobytes()


def C_u64():
    import ctypes
    return ctypes.c_uint64()


def test_decode_worst_case_resynchronisation(codec_mod, oracle):
    """A stream of equal 33-bit codes never resynchronises: every 1024-bit segment's entry point depends on
    its predecessor's, so index discovery needs one fix-up round per segment (the host fallback loop)."""
    q = np.full((4, 8, 8, 8), -32768, np.int16)
    q[1, 0, 0, 0] = 5          # a little structure so that the cubes differ
    q[3, 7, 7, 7] = 0
    ref, ref_end = oracle.eg_encode_cubes(q.astype(np.int32), 8, cap=5 * q.size + 16)
    with make(codec_mod, 64, 64, 8) as c:
        out, end = c.eg_encode_i16(q)
        assert end == ref_end and out.tobytes() == ref.tobytes()
        back, dend = c.eg_decode_i16(out, 4)
        assert dend == ref_end and (back == q).all()


# ---- colour planes ---------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 2, 3, 47, 48, 49, 50, 3 * 4097, 1_000_001])
def test_rgb_split_mix_host(codec_mod, n):
    """RGBUtils.split sends byte i to plane i % 3 (J/RGBUtils.java:67-80); mix interleaves back (:115-119)."""
    rng = np.random.default_rng(n)
    rgb = rng.integers(0, 256, n, dtype=np.uint8)
    with make(codec_mod, 16, 16, 8) as c:
        r, g, b = c.rgb_split(rgb)
        assert (r == rgb[0::3]).all() and (g == rgb[1::3]).all() and (b == rgb[2::3]).all()
        m = n // 3
        back = c.rgb_mix(r[:m], g[:m], b[:m])
        assert (back == rgb[: 3 * m]).all()


@pytest.mark.gpu
def test_rgb_dev_aligned_and_unaligned(codec_mod):
    import torch
    dev = torch.device("cuda", 0)
    n = 1920 * 1080 * 3 * 4
    rgb = torch.randint(0, 256, (n + 16,), dtype=torch.uint8, device=dev)
    with make(codec_mod, 1920, 1080, 8) as c:
        for off in (0, 1):                          # 16-byte aligned -> vector path, otherwise the byte path
            src = rgb[off:off + n]
            planes = [torch.empty(n // 3 + 16, dtype=torch.uint8, device=dev)[off:off + n // 3] for _ in range(3)]
            torch.cuda.synchronize()
            c.rgb_split_dev(src, n, *planes)
            torch.cuda.synchronize()
            for p in range(3):
                assert torch.equal(planes[p], src[p::3])
            out = torch.zeros(n + 16, dtype=torch.uint8, device=dev)[off:off + n]
            torch.cuda.synchronize()
            c.rgb_mix_dev(*planes, n // 3, out)
            torch.cuda.synchronize()
            assert torch.equal(out, src)


@pytest.mark.gpu
def test_rgbutils_cli_and_colour_clip(codec_mod, synth, tmp_path):
    """A colour clip is three gray clips: split, code each plane, decode, mix (README of the reference)."""
    W, H, F = 64, 48, 8
    planes = [synth.natural(W, H, F, s) for s in (1, 2, 3)]
    rgb = np.stack(planes, axis=-1).reshape(-1)
    (tmp_path / "in.rgb").write_bytes(rgb.tobytes())
    assert codec_mod.RGBUtils.main(["split", str(tmp_path / "in.rgb"), str(tmp_path / "p")]) == 0
    for ext, p in zip((".red", ".green", ".blue"), planes):
        assert (np.fromfile(str(tmp_path / "p") + ext, np.uint8) == p.reshape(-1)).all()
    with make(codec_mod, W, H, 8) as c:
        for ext, p in zip((".red", ".green", ".blue"), planes):
            stream, nbits = c.encode_u8(p)
            c.decode_u8(stream, F).tofile(str(tmp_path / "q") + ext)
    assert codec_mod.RGBUtils.main(["mix", str(tmp_path / "q"), str(tmp_path / "out.rgb")]) == 0
    out = np.fromfile(str(tmp_path / "out.rgb"), np.uint8)
    assert out.size == rgb.size and np.abs(out.astype(int) - rgb.astype(int)).mean() < 6.0


# ---- fp64 mode (option "precision" = 64): the Java flavour's arithmetic, no tie flips -----------------
@pytest.mark.gpu
@pytest.mark.parametrize("kind,seed", [("natural", 1), ("noise", 2)])
@pytest.mark.parametrize("mode", [0, 1])
def test_fp64_mode_matches_oracle_without_flips(codec_mod, oracle, synth, kind, seed, mode):
    """In fp64 the quantiser sees the same values as the fp64 oracle to ~1e-12, so for the 8^3 transform (no exact
    ties) the quantised cubes and therefore the whole Exp-Golomb stream are identical, in both rounding flavours
    (J/Encoder.java:82 Math.round, C/encoder.c:53 round)."""
    W, H, F = 128, 64, 16
    clip = gen(synth, kind, W, H, F, seed)
    ref = oracle.quantized_cubes(clip, 8, mode=mode)
    ref_stream, ref_bits = oracle.encode_u8(clip, 8, mode)
    with make(codec_mod, W, H, 8) as c:
        c.set_option("precision", 64)
        c.set_option("rounding", mode)
        q = c.quantize_u8(clip).astype(np.int32)
        assert int((q != ref).sum()) == 0
        stream, nbits = c.encode_u8(clip)
        assert nbits == ref_bits and stream.tobytes() == ref_stream[: nbits // 8 + 1].tobytes()
        dec = c.decode_u8(stream, F)
        want = oracle.decode_u8(ref_stream, W, H, F, 8)
        # truncation after fp64 arithmetic in a different summation order: a pixel may differ by one only where
        # the exact value is an integer to within 1e-9
        assert np.abs(dec.astype(int) - want.astype(int)).max() <= 1
        assert (dec != want).mean() < 1e-3
        assert (c.reconstruct_i16(q.astype(np.int16), F) == dec).all()
        c.set_option("precision", 32)
        q32 = c.quantize_u8(clip).astype(np.int32)
        assert np.abs(q32 - q).max() <= 1           # and the fp32 path differs from it only by tie flips


@pytest.mark.gpu
def test_fp64_mode_cube4_flips_only_at_exact_ties(codec_mod, oracle, synth):
    W, H, F = 64, 32, 8
    clip = gen(synth, "natural", W, H, F, 4)
    ref, coef = oracle.quantized_cubes(clip, 4, mode=0, want_coef=True)
    with make(codec_mod, W, H, 4) as c:
        c.set_option("precision", 64)
        q = c.quantize_u8(clip).astype(np.int32)
    exact, near = classify_flips(q, ref, coef, oracle, 4)
    assert near == 0                                # the 4^3 basis is rational: only exact ties can differ


@pytest.mark.gpu
def test_randomised_shapes_and_contents(codec_mod, oracle):
    """40 seeded random clips: odd cube counts per row (partial warp units, TMA on and off), every content mix.
    For each: the fused stream equals the oracle's Exp-Golomb coding of the cubes the quantise entry point
    returns, index discovery recovers exactly those cubes, and the fused decode equals the staged one."""
    rng = np.random.default_rng(2026)
    for it in range(40):
        cube = 8 if it % 5 else 4
        W = cube * int(rng.integers(1, 41))
        H = cube * int(rng.integers(1, 13))
        F = cube * int(rng.integers(1, 4)) + int(rng.integers(0, cube))      # trailing frames are ignored
        kind = it % 4
        if kind == 0:
            clip = rng.integers(0, 256, (F, H, W), dtype=np.uint8)
        elif kind == 1:
            clip = np.full((F, H, W), int(rng.integers(0, 256)), np.uint8)
        elif kind == 2:
            clip = (128 + 100 * np.sin(np.arange(W) / 7.0)[None, None, :] * np.cos(np.arange(F) / 3.0)[:, None, None]
                    + rng.normal(0, 3, (F, H, W))).clip(0, 255).astype(np.uint8)
        else:
            clip = np.zeros((F, H, W), np.uint8)
            clip[rng.random((F, H, W)) < 0.01] = 255
        Fe = F - F % cube
        with make(codec_mod, W, H, cube) as c:
            stream, nbits = c.encode_u8(clip)
            q = c.quantize_u8(clip)
            ref, ref_bits = oracle.eg_encode_cubes(q.astype(np.int32), cube, cap=5 * q.size + 64)
            assert nbits == ref_bits and stream.tobytes() == ref[: nbits // 8 + 1].tobytes(), (it, W, H, F, cube)
            qd, end = c.eg_decode_i16(stream, q.shape[0])
            assert end == nbits and (qd == q).all(), (it, W, H, F, cube)
            dec = c.decode_u8(stream, F)
            assert dec.shape == (Fe, H, W) and (dec == c.reconstruct_i16(q, Fe)).all(), (it, W, H, F, cube)


@pytest.mark.gpu
def test_locate_and_sharded_decode_without_side_information(codec_mod, oracle, synth):
    """SURVEY.md 8e, decode side: a rank that decodes a later slab range of the one concatenated stream finds its
    first bit by index discovery alone (dct3d_eg_locate), and its frames equal the one-shot decode's."""
    sh = pkg("sharding")
    W, H, F = 128, 64, 48
    clip = gen(synth, "natural", W, H, F, 11)
    with make(codec_mod, W, H, 8) as c:
        stream, nbits = c.encode_u8(clip)
        whole = c.decode_u8(stream, F)
        nslabs, cps = F // 8, (W // 8) * (H // 8)
        # every slab boundary, against the oracle's sequential reader
        q = c.quantize_u8(clip).astype(np.int32)
        for s in range(nslabs + 1):
            _, want = oracle.eg_encode_cubes(q[: s * cps], 8, cap=5 * q.size + 64) if s else (None, 0)
            assert c.eg_locate(stream, s * cps) == want
        assert c.eg_locate(stream, nslabs * cps) == nbits
        for world in (2, 3, 4):
            parts = [sh.decode_range(c, stream, nslabs, r, world, cps)[0] for r in range(world)]
            assert (np.concatenate(parts) == whole).all()
        with pytest.raises(codec_mod.Dct3dError):
            c.eg_locate(stream, nslabs * cps + 1)           # more cubes than the stream holds
