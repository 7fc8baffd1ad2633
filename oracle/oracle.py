"""ctypes bindings for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product (libdct3d.so and the
3ddctvideoencoding_b200 package) never does.  See oracle/dct3d_oracle.c for the
reference file:line each function restates.

Two libraries:
  * liboracle.so       -- our restatement (dct3d_oracle.c)
  * _ref/libref_c.so   -- the reference's own C sources compiled unmodified from
                          /root/reference (oracle/Makefile); optional at run time.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
UINT64_MAX = 2**64 - 1


def build(force: bool = False) -> None:
    """Compile liboracle.so and, when /root/reference is present, oracle/_ref."""
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "dct3d_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so", "CC=gcc"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", HERE, "ref", "CC=gcc"], stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_zigzag.argtypes = [C.c_int, C.c_int, C.c_int, i32p]
        L.orc_zigzag.restype = C.c_int
        L.orc_eg_encode.argtypes = [i32p, C.c_size_t, u8p, C.c_size_t, C.c_uint64]
        L.orc_eg_encode.restype = C.c_uint64
        L.orc_eg_codelen.argtypes = [C.c_int32]
        L.orc_eg_codelen.restype = C.c_int
        L.orc_eg_decode.argtypes = [u8p, C.c_size_t, C.c_uint64, C.c_size_t, i32p]
        L.orc_eg_decode.restype = C.c_uint64
        for name in ("orc_dct3d_direct_f64", "orc_idct3d_direct_f64", "orc_dct3d_sep_f64", "orc_idct3d_sep_f64"):
            f = getattr(L, name)
            f.argtypes = [f64p, f64p, C.c_int, C.c_int, C.c_int, C.c_int]
            f.restype = None
        L.orc_quantize_planar.argtypes = [f64p, i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_quantize_planar.restype = None
        L.orc_dequantize_planar.argtypes = [i32p, f64p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_dequantize_planar.restype = None
        L.orc_quantized_cubes_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, C.c_void_p]
        L.orc_quantized_cubes_u8.restype = None
        L.orc_eg_encode_cubes.argtypes = [i32p, C.c_size_t, C.c_int, u8p, C.c_size_t, C.c_uint64]
        L.orc_eg_encode_cubes.restype = C.c_uint64
        L.orc_eg_decode_cubes.argtypes = [u8p, C.c_size_t, C.c_uint64, C.c_size_t, C.c_int, i32p]
        L.orc_eg_decode_cubes.restype = C.c_uint64
        L.orc_encode_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_size_t]
        L.orc_encode_u8.restype = C.c_uint64
        L.orc_reconstruct_u8.argtypes = [i32p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_reconstruct_u8.restype = None
        L.orc_decode_u8.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_decode_u8.restype = C.c_int
        for name in ("orc_cl_dct_f32", "orc_cl_idct_f32"):
            f = getattr(L, name)
            f.argtypes = [f32p, f32p, C.c_size_t, C.c_int]
            f.restype = None
        L.orc_java_plan_stats.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_java_plan_stats.restype = None
        for name in ("orc_java_dct_f64", "orc_java_idct_f64"):
            f = getattr(L, name)
            f.argtypes = [f64p, f64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
            f.restype = None
        L.orc_java_encode_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_size_t]
        L.orc_java_encode_u8.restype = C.c_uint64
        L.orc_java_decode_u8.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_java_decode_u8.restype = C.c_int
        _LIB = L
    return _LIB


# ---------------------------------------------------------------------------
# numpy-level API
# ---------------------------------------------------------------------------
def zigzag(cube: int) -> np.ndarray:
    out = np.zeros(cube ** 3, np.int32)
    n = lib().orc_zigzag(cube, cube, cube, out)
    assert n == cube ** 3
    return out


def eg_encode(values, start_bit: int = 0, cap: int | None = None, buf: np.ndarray | None = None):
    """-> (bytes array of floor(end/8)+1 bytes, end_bit)."""
    v = np.ascontiguousarray(values, np.int32)
    if buf is None:
        cap = cap or (start_bit // 8 + 9 * v.size + 16)
        buf = np.zeros(cap, np.uint8)
    end = lib().orc_eg_encode(v, v.size, buf, buf.size, start_bit)
    if end == UINT64_MAX:
        raise OverflowError("Exp-Golomb buffer overflow")
    return buf[: end // 8 + 1], int(end)


def eg_decode(buf, n: int, start_bit: int = 0):
    b = np.ascontiguousarray(buf, np.uint8)
    out = np.zeros(n, np.int32)
    end = lib().orc_eg_decode(b, b.size, start_bit, n, out)
    if end == UINT64_MAX:
        raise ValueError("Exp-Golomb stream truncated")
    return out, int(end)


def _planar(fn, a, W, H, F, cube):
    a = np.ascontiguousarray(a, np.float64).reshape(-1)
    out = np.zeros_like(a)
    getattr(lib(), fn)(a, out, W, H, F, cube)
    return out.reshape(F, H, W)


def dct_direct(px, cube=8):
    F, H, W = px.shape
    return _planar("orc_dct3d_direct_f64", px, W, H, F, cube)


def idct_direct(cf, cube=8):
    F, H, W = cf.shape
    return _planar("orc_idct3d_direct_f64", cf, W, H, F, cube)


def dct_sep(px, cube=8):
    F, H, W = px.shape
    return _planar("orc_dct3d_sep_f64", px, W, H, F, cube)


def idct_sep(cf, cube=8):
    F, H, W = cf.shape
    return _planar("orc_idct3d_sep_f64", cf, W, H, F, cube)


def quantize_planar(coef, cube=8, mode=0):
    F, H, W = coef.shape
    q = np.zeros(coef.size, np.int32)
    lib().orc_quantize_planar(np.ascontiguousarray(coef, np.float64).reshape(-1), q, W, H, F, cube, mode)
    return q.reshape(-1, cube, cube, cube)


def quantized_cubes(frames, cube=8, mode=0, want_coef=False):
    """u8 frames [F][H][W] -> int32 cubes [ncubes][k0][k1][k2] (+ fp64 planar coefficients)."""
    fr = np.ascontiguousarray(frames, np.uint8)
    F, H, W = fr.shape
    Fe = F - F % cube
    q = np.zeros(W * H * Fe, np.int32)
    coef = np.zeros(W * H * Fe, np.float64) if want_coef else None
    lib().orc_quantized_cubes_u8(fr.reshape(-1), W, H, Fe, cube, mode, q,
                                 coef.ctypes.data_as(C.c_void_p) if want_coef else None)
    q = q.reshape(-1, cube, cube, cube)
    return (q, coef.reshape(Fe, H, W)) if want_coef else q


def eg_encode_cubes(q, cube=8, start_bit=0, cap=None):
    qq = np.ascontiguousarray(q, np.int32).reshape(-1)
    n = qq.size // cube ** 3
    cap = cap or (start_bit // 8 + 4 * qq.size + 16)
    buf = np.zeros(cap, np.uint8)
    end = lib().orc_eg_encode_cubes(qq, n, cube, buf, buf.size, start_bit)
    if end == UINT64_MAX:
        raise OverflowError("Exp-Golomb buffer overflow")
    return buf[: end // 8 + 1], int(end)


def eg_decode_cubes(buf, ncubes, cube=8, start_bit=0):
    b = np.ascontiguousarray(buf, np.uint8)
    q = np.zeros(ncubes * cube ** 3, np.int32)
    end = lib().orc_eg_decode_cubes(b, b.size, start_bit, ncubes, cube, q)
    if end == UINT64_MAX:
        raise ValueError("Exp-Golomb stream truncated")
    return q.reshape(-1, cube, cube, cube), int(end)


def encode_u8(frames, cube=8, mode=0):
    fr = np.ascontiguousarray(frames, np.uint8)
    F, H, W = fr.shape
    buf = np.zeros(4 * fr.size + 16, np.uint8)
    bits = lib().orc_encode_u8(fr.reshape(-1), W, H, F, cube, mode, buf, buf.size)
    if bits == UINT64_MAX:
        raise OverflowError
    return buf[: bits // 8 + 1].copy(), int(bits)


def reconstruct_u8(q, W, H, F, cube=8):
    out = np.zeros(W * H * F, np.uint8)
    lib().orc_reconstruct_u8(np.ascontiguousarray(q, np.int32).reshape(-1), W, H, F, cube, out)
    return out.reshape(F, H, W)


def decode_u8(buf, W, H, F, cube=8):
    b = np.ascontiguousarray(buf, np.uint8)
    Fe = F - F % cube
    out = np.zeros(W * H * Fe, np.uint8)
    rc = lib().orc_decode_u8(b, b.size, W, H, F, cube, out)
    if rc != 0:
        raise ValueError("stream truncated")
    return out.reshape(Fe, H, W)


def cl_dct_f32(cubes, cube=8):
    a = np.ascontiguousarray(cubes, np.float32).reshape(-1)
    out = np.zeros_like(a)
    lib().orc_cl_dct_f32(a, out, a.size // cube ** 3, cube)
    return out.reshape(-1, cube, cube, cube)


def cl_idct_f32(cubes, cube=8):
    a = np.ascontiguousarray(cubes, np.float32).reshape(-1)
    out = np.zeros_like(a)
    lib().orc_cl_idct_f32(a, out, a.size // cube ** 3, cube)
    return out.reshape(-1, cube, cube, cube)


def java_plan_stats(cube=8):
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    lib().orc_java_plan_stats(cube, C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def java_dct(px, cube=8, threads=1):
    F, H, W = px.shape
    a = np.ascontiguousarray(px, np.float64).reshape(-1)
    out = np.zeros_like(a)
    lib().orc_java_dct_f64(a, out, W, H, F, cube, threads)
    return out.reshape(F, H, W)


def java_idct(cf, cube=8, threads=1):
    F, H, W = cf.shape
    a = np.ascontiguousarray(cf, np.float64).reshape(-1)
    out = np.zeros_like(a)
    lib().orc_java_idct_f64(a, out, W, H, F, cube, threads)
    return out.reshape(F, H, W)


def java_encode_u8(frames, cube=8, threads=1):
    fr = np.ascontiguousarray(frames, np.uint8)
    F, H, W = fr.shape
    buf = np.zeros(4 * fr.size + 16, np.uint8)
    bits = lib().orc_java_encode_u8(fr.reshape(-1), W, H, F, cube, threads, buf, buf.size)
    return buf[: bits // 8 + 1].copy(), int(bits)


def java_decode_u8(buf, W, H, F, cube=8, threads=1):
    b = np.ascontiguousarray(buf, np.uint8)
    Fe = F - F % cube
    out = np.zeros(W * H * Fe, np.uint8)
    rc = lib().orc_java_decode_u8(b, b.size, W, H, F, cube, threads, out)
    if rc != 0:
        raise ValueError("stream truncated")
    return out.reshape(Fe, H, W)


def frames_to_cubes(frames, cube=8):
    """u8/float frames [F][H][W] -> cube-major [slab][by][bx][k0][k1][k2] flattened to [ncubes][c][c][c]
    (the layout of C/encoder.c:29-41 readCubes and J/Encoder.java:75-89)."""
    a = np.asarray(frames)
    F, H, W = a.shape
    c = cube
    a = a[: F - F % c].reshape(F // c, c, H // c, c, W // c, c)
    return np.ascontiguousarray(a.transpose(0, 2, 4, 1, 3, 5)).reshape(-1, c, c, c)


def cubes_to_frames(cubes, W, H, cube=8):
    c = cube
    a = np.asarray(cubes).reshape(-1, H // c, W // c, c, c, c)
    S = a.shape[0]
    return np.ascontiguousarray(a.transpose(0, 3, 1, 4, 2, 5)).reshape(S * c, H, W)


# ---------------------------------------------------------------------------
# The reference's own C code (oracle/_ref/libref_c.so), when built.
# ---------------------------------------------------------------------------
class _EGStream(C.Structure):  # mirrors struct ExpGolombStream (C/ExpGolomb.h:4-8)
    _fields_ = [("buffer", C.c_void_p), ("bitPosition", C.c_int), ("bufferPosition", C.c_int)]


class _Coord(C.Structure):  # C/CubeUtils.h:11-15
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("z", C.c_int)]


class _Slices(C.Structure):  # C/CubeUtils.h:17-20
    _fields_ = [("positions", C.POINTER(_Coord)), ("length", C.c_int)]


def ref():
    """The reference's compiled C code, or None if oracle/_ref was never built."""
    global _REF
    if _REF is None:
        so = os.path.join(HERE, "_ref", "libref_c.so")
        if not os.path.exists(so):
            return None
        R = C.CDLL(so)
        R.cubeUtils_diagonalSlices.argtypes = [C.c_int, C.c_int, C.c_int]
        R.cubeUtils_diagonalSlices.restype = C.POINTER(_Slices)
        R.expGolomb_createStream.argtypes = [C.c_void_p]
        R.expGolomb_createStream.restype = C.POINTER(_EGStream)
        R.expGolomb_writeValue.argtypes = [C.POINTER(_EGStream), C.c_int]
        R.expGolomb_writeValue.restype = None
        R.expGolomb_readValue.argtypes = [C.POINTER(_EGStream)]
        R.expGolomb_readValue.restype = C.c_int
        R.applyQuantization.argtypes = [f32p, C.c_size_t]
        R.applyQuantization.restype = None
        R.applyDequantization.argtypes = [f32p, C.c_size_t]
        R.applyDequantization.restype = None
        R.applyExpGolombCoding.argtypes = [f32p, C.c_size_t, C.POINTER(_Slices), C.POINTER(_EGStream)]
        R.applyExpGolombCoding.restype = C.c_int
        R.reorderDctCoeffs.argtypes = [f32p, C.c_size_t, f32p, C.POINTER(_Slices)]
        R.reorderDctCoeffs.restype = None
        R.encode.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
        R.encode.restype = C.c_int
        R.decode.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
        R.decode.restype = C.c_int
        _REF = R
    return _REF


def ref_zigzag(cube: int) -> np.ndarray:
    s = ref().cubeUtils_diagonalSlices(cube, cube, cube).contents
    return np.array([s.positions[i].x + s.positions[i].y * cube + s.positions[i].z * cube * cube
                     for i in range(s.length)], np.int32)


def ref_eg_write(values, cap=None):
    """Run the reference's expGolomb_writeValue over `values`; -> (bytes incl. the partial byte, total bits)."""
    R = ref()
    vals = [int(v) for v in values]
    cap = cap or (9 * len(vals) + 16)
    buf = np.zeros(cap, np.uint8)
    st = R.expGolomb_createStream(buf.ctypes.data)
    for v in vals:
        R.expGolomb_writeValue(st, v)
    pos, bitpos = st.contents.bufferPosition, st.contents.bitPosition
    return buf[: pos + 1].copy(), pos * 8 + (8 - bitpos)


def ref_eg_read(buf, n):
    R = ref()
    b = np.concatenate([np.ascontiguousarray(buf, np.uint8), np.zeros(8, np.uint8)])
    st = R.expGolomb_createStream(b.ctypes.data)
    return np.array([R.expGolomb_readValue(st) for _ in range(n)], np.int32)


def ref_quantize_f32(cubes_f32):
    a = np.ascontiguousarray(cubes_f32, np.float32).reshape(-1).copy()
    ref().applyQuantization(a, a.size)
    return a


def ref_dequantize_f32(cubes_f32):
    a = np.ascontiguousarray(cubes_f32, np.float32).reshape(-1).copy()
    ref().applyDequantization(a, a.size)
    return a


def ref_eg_encode_cubes_f32(qcubes_f32):
    """Reference applyExpGolombCoding on cube-major float data (8^3 only: DCT_BLOCK_* are compile-time)."""
    R = ref()
    a = np.ascontiguousarray(qcubes_f32, np.float32).reshape(-1)
    buf = np.zeros(4 * a.size + 16, np.uint8)
    st = R.expGolomb_createStream(buf.ctypes.data)
    sl = R.cubeUtils_diagonalSlices(8, 8, 8)
    R.applyExpGolombCoding(a, a.size, sl, st)
    pos, bitpos = st.contents.bufferPosition, st.contents.bitPosition
    return buf[: pos + 1].copy(), pos * 8 + (8 - bitpos)
