"""CPU tests of the kernels' __host__ __device__ arithmetic (csrc/dct_math.h, csrc/eg_bits.h)
compiled for the host by tests/host_harness.cpp, against the oracle.  These are unit tests of
device functions, not a product path: libdct3d.so has no CPU mode."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def hh():
    # DCT3D_HH_DEFS: extra -D flags, to run this file against a compile-time variant of the device arithmetic
    defs = os.environ.get("DCT3D_HH_DEFS", "").split()
    so = os.path.join(HERE, "_host_harness%s.so" % ("_" + "_".join(d.removeprefix("-D").replace("=", "") for d in defs) if defs else ""))
    srcs = [os.path.join(HERE, "host_harness.cpp")] + [
        os.path.join(ROOT, "3ddctvideoencoding_b200", "csrc", f) for f in ("dct_math.h", "eg_bits.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared"] + defs + ["-o", so, srcs[0]])
    L = C.CDLL(so)
    f32 = np.ctypeslib.ndpointer(np.float32, flags="C")
    f64 = np.ctypeslib.ndpointer(np.float64, flags="C")
    i16 = np.ctypeslib.ndpointer(np.int16, flags="C")
    u8 = np.ctypeslib.ndpointer(np.uint8, flags="C")
    u16 = np.ctypeslib.ndpointer(np.uint16, flags="C")
    L.hh_cube_f32.argtypes = [f32, f32, C.c_int, C.c_int]
    L.hh_cube_f64.argtypes = [f64, f64, C.c_int, C.c_int]
    L.hh_cube_scaled_f32.argtypes = [f32, f32, C.c_int, C.c_int]
    L.hh_cube_scaled_f64.argtypes = [f64, f64, C.c_int, C.c_int]
    L.hh_quantize.argtypes = [C.c_float, C.c_int]
    L.hh_quantize.restype = C.c_int
    L.hh_eg_write.argtypes = [i16, C.c_int, C.c_int, C.c_uint64, u8, C.c_size_t]
    L.hh_eg_write.restype = C.c_uint64
    L.hh_eg_parse.argtypes = [u8, C.c_size_t, C.c_uint64, C.c_int, C.c_int, u16, i16]
    L.hh_eg_parse.restype = C.c_uint64
    L.hh_eg_scan.argtypes = [u8, C.c_size_t, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    L.hh_eg_scan.restype = C.c_int
    return L


@pytest.mark.parametrize("n", [8, 4])
def test_butterfly_matches_fp64_oracle(hh, oracle, n):
    rng = np.random.default_rng(3)
    for trial in range(8):
        px = rng.integers(0, 256, size=(n, n, n)).astype(np.float64)
        ref = oracle.dct_direct(px, n)
        out64 = np.zeros_like(px)
        hh.hh_cube_f64(np.ascontiguousarray(px), out64, n, 0)
        assert np.abs(out64 - ref).max() < 1e-10
        out32 = np.zeros((n, n, n), np.float32)
        hh.hh_cube_f32(px.astype(np.float32), out32, n, 0)
        # north-star rule (1): 1e-4 relative to the block's magnitude
        assert np.abs(out32 - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
        assert np.abs(out32 - ref).max() < 2e-3
        back = np.zeros_like(px)
        hh.hh_cube_f64(out64, back, n, 1)
        assert np.abs(back - px).max() < 1e-9
        back32 = np.zeros((n, n, n), np.float32)
        hh.hh_cube_f32(out32, back32, n, 1)
        assert np.abs(back32 - px).max() < 1e-3


@pytest.mark.parametrize("n", [8, 4])
def test_scaled_butterflies_match_fp64_oracle(hh, oracle, n):
    """The normalised x/y butterflies + scaled t butterfly + S[k1] factor equal the plain transform."""
    rng = np.random.default_rng(4)
    for trial in range(8):
        px = rng.integers(0, 256, size=(n, n, n)).astype(np.float64)
        ref = oracle.dct_direct(px, n)
        out64 = np.zeros_like(px)
        hh.hh_cube_scaled_f64(np.ascontiguousarray(px), out64, n, 0)
        assert np.abs(out64 - ref).max() < 1e-10
        out32 = np.zeros((n, n, n), np.float32)
        hh.hh_cube_scaled_f32(px.astype(np.float32), out32, n, 0)
        assert np.abs(out32 - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
        assert np.abs(out32 - ref).max() < 2e-3
        back = np.zeros_like(px)
        hh.hh_cube_scaled_f64(out64, back, n, 1)
        assert np.abs(back - px).max() < 1e-9
        back32 = np.zeros((n, n, n), np.float32)
        hh.hh_cube_scaled_f32(out32, back32, n, 1)
        assert np.abs(back32 - px).max() < 1e-3


def test_quantize_magic_rounding(hh):
    assert hh.hh_quantize(4471.3, 0) == 4471
    assert hh.hh_quantize(-4471.6, 0) == -4472
    assert hh.hh_quantize(12.4, 1) == 2       # /5 = 2.48
    assert hh.hh_quantize(-12.6, 1) == -3     # /5 = -2.52
    assert hh.hh_quantize(0.49, 0) == 0 and hh.hh_quantize(-0.49, 0) == 0
    assert hh.hh_quantize(5770.0, 0) == 5770 and hh.hh_quantize(-5770.0, 0) == -5770


def _rand_cubes(rng, ncubes, cs, density, big=False):
    q = np.zeros((ncubes, cs), np.int16)
    mask = rng.random((ncubes, cs)) < density
    hi = 32767 if big else 300
    q[mask] = rng.integers(-hi - (1 if big else 0), hi + 1, size=int(mask.sum())).astype(np.int16)
    q[:, 0] = rng.integers(-5770, 5771, size=ncubes)
    return q


@pytest.mark.parametrize("cs", [512, 64])
@pytest.mark.parametrize("density,big", [(0.0, False), (0.03, False), (0.5, False), (1.0, True)])
@pytest.mark.parametrize("start", [0, 5, 37])
def test_thread_per_cube_writer_is_bit_exact(hh, oracle, cs, density, big, start):
    rng = np.random.default_rng(cs + int(density * 100) + start)
    q = _rand_cubes(rng, 9, cs, density, big)
    if density == 0.0:
        q[:] = 0
    ref, ref_end = oracle.eg_encode(q.reshape(-1).astype(np.int32), start_bit=start)
    out = np.zeros(ref.size + 16, np.uint8)
    end = hh.hh_eg_write(q, 9, cs, start, out, out.size)
    assert end == ref_end
    assert out[: ref.size].tobytes() == ref.tobytes()
    assert not out[ref.size:].any()
    # parse back (natural order = identity table here)
    izz = np.arange(cs, dtype=np.uint16)
    back = np.zeros_like(q)
    pend = hh.hh_eg_parse(out, out.size, start, 9, cs, izz, back)
    assert pend == ref_end and (back == q).all()


def test_parser_rejects_malformed(hh):
    buf = np.zeros(64, np.uint8)   # all zeros: an endless prefix
    out = np.zeros(512, np.int16)
    assert hh.hh_eg_parse(buf, buf.size, 0, 1, 512, np.arange(512, dtype=np.uint16), out) == 2**64 - 1


def test_segment_scan_counts_and_overhang(hh, oracle):
    rng = np.random.default_rng(5)
    vals = (rng.integers(-3000, 3000, size=4000) * (rng.random(4000) < 0.1)).astype(np.int32)
    stream, end = oracle.eg_encode(vals)
    # code start positions from the oracle
    starts = np.cumsum([0] + [oracle.lib().orc_eg_codelen(int(v)) for v in vals])
    seg = 256
    pos = 0
    total = 0
    nseg = (end + seg - 1) // seg
    for k in range(nseg):
        lim = (k + 1) * seg
        n, nxt = C.c_uint32(), C.c_uint64()
        assert hh.hh_eg_scan(stream, stream.size, pos, lim, C.byref(n), C.byref(nxt)) == 0
        expect = int(((starts[:-1] >= pos) & (starts[:-1] < lim)).sum())
        if k < nseg - 1:
            assert n.value == expect
            assert nxt.value == int(starts[np.searchsorted(starts, lim)])
        else:
            assert n.value >= expect   # padding bits after the last code may parse as codes
        total += n.value
        pos = nxt.value
    assert total >= vals.size


def test_scan_does_not_count_a_code_cut_by_the_end_of_the_stream(hh):
    """bytes FD 80 hold six zeros (1 1 1 1 1 1), then the 3-bit code 0 1 1 across the byte boundary.  Cut to
    one byte the last code is incomplete: it must not be counted, and the scan must not report a position past the
    end of the stream."""
    n, nxt = C.c_uint32(), C.c_uint64()
    two = np.array([0xFD, 0x80], np.uint8)
    assert hh.hh_eg_scan(two, 2, 0, 16, C.byref(n), C.byref(nxt)) == 0
    assert n.value >= 7
    one = np.array([0xFD], np.uint8)
    assert hh.hh_eg_scan(one, 1, 0, 8, C.byref(n), C.byref(nxt)) == 0
    assert n.value == 6 and nxt.value <= 8


def test_scan_rejects_17_bit_code_numbers_beyond_int16(hh, oracle):
    """Only m = 65537 (v = -32768) is a legal 33-bit code: the codec's cubes are int16."""
    def scan(bits):
        bits += "0" * (-len(bits) % 8)
        buf = np.frombuffer(int(bits, 2).to_bytes(len(bits) // 8, "big") + b"\0\0\0\0", np.uint8).copy()
        n, nxt = C.c_uint32(), C.c_uint64()
        rc = hh.hh_eg_scan(buf, buf.size, 0, 64, C.byref(n), C.byref(nxt))
        return rc, n.value, nxt.value
    ok = "111" + "0" * 16 + format(65537, "017b") + "1" * 40
    assert scan(ok)[:2] == (0, 3 + 1 + 28)
    stream, end = oracle.eg_encode(np.array([0, 0, 0, -32768], np.int32))
    assert end == 36 and "".join(format(b, "08b") for b in stream)[:36] == ok[:36]
    for m in (65536, 65538, 0x1FFFF):
        assert scan("111" + "0" * 16 + format(m, "017b") + "1" * 40)[0] == -1


def _bind_round2(hh):
    hh.hh_zero_threshold.argtypes = [C.c_float]
    hh.hh_zero_threshold.restype = C.c_float
    hh.hh_quantize_recip.argtypes = [C.c_float, C.c_float]
    hh.hh_quantize_recip.restype = C.c_int


def test_zero_threshold_is_the_quantisers_zero_set(hh):
    """The fused encoder skips the quantiser of a group of diagonals when every |coefficient| <= zero_threshold(recip).
    For every reciprocal the kernel uses (S[k1] / max(1, 5 (s + k1)), float arithmetic as in encode_kernel) the threshold
    must split the floats exactly where the one-FFMA quantiser switches between zero and non-zero."""
    _bind_round2(hh)
    S = np.array([0.35355339059327376220, 0.49039264020161522456, 0.46193976625564337806, 0.41573480615127261854,
                  0.35355339059327376220, 0.27778511650980111237, 0.46193976625564337806, 0.09754516100806413392], np.float32)
    rng = np.random.default_rng(11)
    for k1 in range(8):
        for s in range(15):
            div = np.float32(1 if s + k1 == 0 else 5 * (s + k1))
            rq = np.float32(S[k1] / div)
            t = np.float32(hh.hh_zero_threshold(float(rq)))
            ti = int(t.view(np.int32))
            # the floats around the threshold, both signs
            for d in range(-4, 5):
                x = np.array([ti + d], np.int32).view(np.float32)[0]
                for v in (x, -x):
                    q = hh.hh_quantize_recip(float(v), float(rq))
                    assert (q == 0) == (d <= 0), (k1, s, d, float(v), q)
            # random magnitudes: the two tests agree everywhere
            xs = (rng.standard_normal(200) * float(t) * 1.5).astype(np.float32)
            for v in xs:
                assert (hh.hh_quantize_recip(float(v), float(rq)) == 0) == (abs(v) <= t)
    # thresholds grow with the diagonal: the bound of diagonal 8 is sufficient for 9..14
    for k1 in range(8):
        ts = [hh.hh_zero_threshold(float(np.float32(S[k1] / np.float32(5 * (s + k1))))) for s in range(8, 15)]
        assert all(a <= b for a, b in zip(ts, ts[1:]))
