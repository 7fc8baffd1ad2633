"""CPU tests: pin the oracle (oracle/dct3d_oracle.c) against the golden vectors
generated from the reference's own C code (tests/golden/make_golden.py) and,
when oracle/_ref is present, against that code directly."""
import hashlib

import numpy as np
import pytest


def fnv1a32(a):
    h = 0x811C9DC5
    for b in np.asarray(a, "<i4").tobytes():
        h = ((h ^ b) * 0x01000193) & 0xFFFFFFFF
    return "%08x" % h


# ---- zig-zag (C/CubeUtils.c:5-46, J/CubeUtils.java:7-41) ---------------------
@pytest.mark.parametrize("cube,fnv", [(8, "2f207a41"), (4, "be688055")])
def test_zigzag_matches_reference_golden(oracle, kat, cube, fnv):
    zz = oracle.zigzag(cube)
    assert zz.tolist() == kat["zigzag%d" % cube]["linear"]
    assert fnv1a32(zz) == fnv == kat["zigzag%d" % cube]["fnv1a32"]
    assert sorted(zz.tolist()) == list(range(cube ** 3))
    c = cube
    sums = [(i % c) + (i // c) % c + i // (c * c) for i in zz.tolist()]
    assert sums == sorted(sums)
    assert np.bincount(sums).tolist() == kat["zigzag%d" % cube]["slice_sizes"]


def test_zigzag_survey_prefix(oracle):
    assert oracle.zigzag(8)[:32].tolist() == [0, 1, 64, 8, 2, 65, 128, 9, 72, 16, 3, 66, 129, 192, 10, 73, 136, 17, 80,
                                              24, 4, 67, 130, 193, 256, 11, 74, 137, 200, 18, 81, 144]
    assert oracle.zigzag(8)[-5:].tolist() == [5 + 7 * 8 + 7 * 64, 7 + 6 * 8 + 7 * 64, 7 + 7 * 8 + 6 * 64,
                                              6 + 7 * 8 + 7 * 64, 511]


# ---- Exp-Golomb (C/ExpGolomb.c:32-110, J/ExpGolombWriter/Reader) -------------
def test_eg_single_value_kats(oracle, kat):
    for e in kat["eg_single"]:
        b, bits = oracle.eg_encode([e["v"]])
        assert bits == e["bits"] == oracle.lib().orc_eg_codelen(e["v"])
        assert b.tobytes().hex() == e["hex"], e
        assert oracle.eg_decode(b, 1)[0][0] == e["v"]


def test_eg_sequence_kat(oracle, kat):
    e = kat["eg_sequence"]
    b, bits = oracle.eg_encode(e["values"])
    assert bits == 56 and b.tobytes().hex() == "005ac1915a7c9f00" == e["hex"]
    assert oracle.eg_decode(b, 16)[0].tolist() == e["values"]


def test_eg_random_stream_golden(oracle, golden):
    vals = golden["eg_values"]
    b, bits = oracle.eg_encode(vals)
    assert bits == int(golden["eg_bits"])
    assert b.tobytes() == golden["eg_bytes"].tobytes()
    back, end = oracle.eg_decode(golden["eg_bytes"], vals.size)
    assert end == bits and (back == vals).all()


def test_eg_start_bit_carry(oracle, golden):
    """Slab-by-slab writing with a carried bit position equals one-shot writing (C/ExpGolomb.c:112-130)."""
    vals = golden["eg_values"]
    one, bits = oracle.eg_encode(vals)
    buf = np.zeros(one.size + 8, np.uint8)
    pos = 0
    for chunk in np.array_split(vals, 7):
        pos = oracle.lib().orc_eg_encode(np.ascontiguousarray(chunk), chunk.size, buf, buf.size, pos)
    assert pos == bits and buf[: bits // 8 + 1].tobytes() == one.tobytes()


def test_eg_overflow_and_truncation(oracle):
    with pytest.raises(OverflowError):
        oracle.eg_encode([5000] * 10, cap=8)
    b, bits = oracle.eg_encode([5000] * 10)
    with pytest.raises(ValueError):
        oracle.eg_decode(b[:5], 10)
    empty, bits0 = oracle.eg_encode(np.zeros(0, np.int32))
    assert bits0 == 0 and empty.size == 1  # floor(0/8)+1 bytes (J/Encoder.java:117)


# ---- transform ---------------------------------------------------------------
@pytest.mark.parametrize("cube", [8, 4])
def test_separable_equals_direct(oracle, synth, cube):
    px = synth.natural(16, 8, cube, 3).astype(np.float64)
    d = oracle.dct_direct(px, cube)
    s = oracle.dct_sep(px, cube)
    assert np.abs(d - s).max() <= 1e-10 * max(1.0, np.abs(d).max())
    # orthonormal: energy preserved, DC = sum * s * c0^3
    assert np.isclose((d ** 2).sum(), (px ** 2).sum(), rtol=1e-12)
    scale = np.sqrt(8.0 / cube ** 3) * (1 / np.sqrt(2)) ** 3
    assert np.isclose(d[0, 0, 0], px[:cube, :cube, :cube].sum() * scale, rtol=1e-12)
    back = oracle.idct_direct(d, cube)
    assert np.abs(back - px).max() < 1e-9
    assert np.abs(oracle.idct_sep(d, cube) - px).max() < 1e-9


@pytest.mark.parametrize("cube", [8, 4])
def test_java_structured_port_equals_separable(oracle, synth, cube):
    px = synth.noise(32, 16, 2 * cube, 5).astype(np.float64)
    a = oracle.java_dct(px, cube, threads=3)
    b = oracle.dct_sep(px, cube)
    # grouping equal coefficients to 1e-9 (J/dct/DCT.java:115) changes the result by < 1e-6 absolute
    assert np.abs(a - b).max() < 1e-5
    q = oracle.quantize_planar(b, cube, 0)
    cf = np.zeros_like(b).reshape(-1)
    oracle.lib().orc_dequantize_planar(q.reshape(-1), cf, 32, 16, 2 * cube, cube)
    cf = cf.reshape(b.shape)
    assert np.abs(oracle.java_idct(cf, cube, threads=2) - oracle.idct_sep(cf, cube)).max() < 1e-9


def test_java_plan_statistics(oracle):
    # replaying DCT.initialize()/createSums()
    assert oracle.java_plan_stats(8) == (11567, 2319, 65024)
    assert oracle.java_plan_stats(4) == (333, 169, 2112)


def test_cl_float_restatement_close_to_fp64(oracle, synth):
    clip = synth.natural(16, 16, 8, 4)
    cubes = oracle.frames_to_cubes(clip).astype(np.float32)
    f = oracle.cl_dct_f32(cubes)
    d = oracle.frames_to_cubes(oracle.dct_sep(clip.astype(np.float64)))
    assert np.abs(f - d).max() < 2e-2  # naive 512-term float sums
    back = oracle.cl_idct_f32(f)
    assert np.abs(back - cubes).max() < 1e-2


# ---- quantiser rounding (J/Encoder.java:82 vs C/encoder.c:53) ------------------
def test_rounding_modes(oracle):
    coef = np.zeros((8, 8, 8))
    coef[0, 0, 0] = 2.5
    coef[0, 0, 1] = -12.5   # divisor 5 -> -2.5
    coef[0, 1, 0] = 7.5     # divisor 5 -> 1.5
    coef[1, 1, 1] = -22.5   # divisor 15 -> -1.5
    qj = oracle.quantize_planar(coef, 8, 0).reshape(8, 8, 8)
    qc = oracle.quantize_planar(coef, 8, 1).reshape(8, 8, 8)
    assert (qj[0, 0, 0], qj[0, 0, 1], qj[0, 1, 0], qj[1, 1, 1]) == (3, -2, 2, -1)   # floor(v+0.5)
    assert (qc[0, 0, 0], qc[0, 0, 1], qc[0, 1, 0], qc[1, 1, 1]) == (3, -3, 2, -2)   # half away from zero


def test_c_quantiser_golden(oracle, golden):
    coef = golden["quant_in"].astype(np.float64).reshape(-1, 8, 8, 8)
    planar = oracle.cubes_to_frames(coef, 64, 8)   # 8 cubes side by side
    q = oracle.quantize_planar(planar, 8, 1).reshape(-1)
    assert (q == golden["quant_out"].astype(np.int32)).all()
    cf = np.zeros(q.size)
    oracle.lib().orc_dequantize_planar(q, cf, 64, 8, 8, 8)
    assert (oracle.frames_to_cubes(cf.reshape(8, 8, 64)).reshape(-1) == golden["dequant_out"]).all()


# ---- whole flow vs the reference CLI run (tests/golden/make_golden.py) ---------
def test_reference_flow_golden(oracle, synth, golden):
    clip = synth.natural(64, 48, 24, 1)
    assert hashlib.sha256(clip.tobytes()).digest() == golden["flow_clip_sha"].tobytes()
    # 73 728 values, 84 744 bits, 10 594 bytes, slab offsets 0/28208/56388
    assert int(golden["flow_bits"]) == 84744 and golden["flow_stream"].size == 10594
    assert golden["flow_slab_bit_offsets"].tolist() == [0, 28208, 56388, 84744]
    q = oracle.quantized_cubes(clip, 8, mode=1)
    ref_q = golden["flow_qcubes"].astype(np.int32)
    assert q.reshape(-1, 512)[0][oracle.zigzag(8)][:16].tolist() == [4471, -15, -39, -3, 0, -1, -3, 0, 3, 1, -1, 0, 0,
                                                                     -1, 0, -1]
    flips = int((q != ref_q).sum())
    assert flips <= 4 and np.abs(q - ref_q).max() <= 1   # reference ran float/cosf; we run fp64
    stream, bits = oracle.encode_u8(clip, 8, mode=1)
    if flips == 0:
        assert bits == 84744 and stream.tobytes() == golden["flow_stream"].tobytes()
    # bit-exact Exp-Golomb given identical cubes
    s2, b2 = oracle.eg_encode_cubes(ref_q)
    assert b2 == 84744 and s2.tobytes() == golden["flow_stream"].tobytes()
    # decode of the reference stream vs the reference decoder's pixels: +-1
    dec = oracle.decode_u8(golden["flow_stream"], 64, 48, 24)
    assert np.abs(dec.astype(int) - golden["flow_decoded"].astype(int)).max() <= 1
    err = np.abs(dec.astype(int) - clip.astype(int))
    assert err.max() == 26  # lossy by design


def test_frames_not_multiple_of_cube_are_dropped(oracle, synth):
    clip = synth.natural(16, 16, 11, 9)
    s, bits = oracle.encode_u8(clip, 8, 0)
    s8, bits8 = oracle.encode_u8(clip[:8], 8, 0)
    assert bits == bits8 and s.tobytes() == s8.tobytes()   # J/Encoder.java:39-40
    assert oracle.decode_u8(s, 16, 16, 11).shape == (8, 16, 16)


def test_constant_clip_stream_size(oracle, synth):
    clip = synth.constant(16, 16, 8, 128)
    s, bits = oracle.encode_u8(clip, 8, 0)
    # DC = 128*512/(16*sqrt 2) = 2896.3 -> 2896 -> m = 5792 (13 bits) -> 25-bit code + 511 one-bit zeros
    assert bits == 4 * (25 + 511)
    assert (oracle.decode_u8(s, 16, 16, 8).astype(int) - 128).__abs__().max() <= 1


# ---- oracle vs the reference's compiled code, when present ---------------------
def _need_ref(oracle):
    if oracle.ref() is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")


def test_ref_zigzag_live(oracle):
    _need_ref(oracle)
    for c in (8, 4):
        assert (oracle.ref_zigzag(c) == oracle.zigzag(c)).all()


def test_ref_expgolomb_live(oracle):
    _need_ref(oracle)
    rng = np.random.default_rng(11)
    vals = (rng.integers(-6000, 6000, size=500) * (rng.random(500) < 0.3)).astype(np.int32)
    rb, rbits = oracle.ref_eg_write(vals.tolist())
    b, bits = oracle.eg_encode(vals)
    assert bits == rbits and b.tobytes() == rb.tobytes()
    assert (oracle.ref_eg_read(b, vals.size) == vals).all()


def test_ref_quantise_and_code_live(oracle, synth):
    _need_ref(oracle)
    clip = synth.natural(32, 16, 8, 6)
    q, coef = oracle.quantized_cubes(clip, 8, mode=1, want_coef=True)
    cubes32 = oracle.frames_to_cubes(coef).astype(np.float32)
    rq = oracle.ref_quantize_f32(cubes32).astype(np.int32).reshape(q.shape)
    assert (rq != q).sum() <= 1   # float(coef) vs double(coef) can flip a near-tie
    rb, rbits = oracle.ref_eg_encode_cubes_f32(rq.astype(np.float32))
    b, bits = oracle.eg_encode_cubes(rq)
    assert bits == rbits and b.tobytes() == rb.tobytes()
