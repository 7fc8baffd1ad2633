"""Host-side mirror of the reference's entry points over libdct3d.so.

Mirrors (reference paths: J/ = 3d-DCT-video-encoding/src/br/jpiccoli/video/,
C/ = 3d-DCT-video-encoding-OpenCL/):

  * ``DCT`` / ``InverseDCT``  -- ``new DCT(in, out, w, h, cw, ch, cd).run()`` (J/dct/Transform.java:44-65,
    J/Encoder.java:63-64, J/Decoder.java:102-103): planar double arrays, caller-owned output.
  * ``Encoder.main`` / ``Decoder.main`` -- the Java command lines (J/Encoder.java:14-129,
    J/Decoder.java:15-121): raw grayscale in, zlib(Exp-Golomb) out and back, default deflate level.
  * ``codec_main`` -- the C CLI ``codec encode|decode|list_platforms`` (C/main.c:11-49), slab-by-slab
    flow with a carried bit position and Z_BEST_COMPRESSION (C/encoder.c:136-139,203-278).

All arithmetic happens in the CUDA library; this file only moves bytes and wraps zlib, which
the reference also leaves to its host code.  No CPU fallback exists: without a GPU every call
raises ``Dct3dError``.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import zlib

import numpy as np

from . import _lib


class Dct3dError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dct3d error {code}: {msg}")
        self.code = code


def _ptr(a) -> int:
    """Device or host address of a numpy array / torch tensor / int."""
    if a is None:
        return 0
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch tensor


class Codec:
    """One libdct3d context: a GPU, a frame size and a cube edge (8 or 4)."""

    def __init__(self, width: int, height: int, cube: int = 8, device: int = 0):
        self.L = _lib.load()
        self.width, self.height, self.cube, self.device = width, height, cube, device
        h = C.c_void_p()
        rc = self.L.dct3d_create(C.byref(h), device, width, height, cube)
        if rc != _lib.OK:
            raise Dct3dError(rc, (self.L.dct3d_last_error(None) or b"").decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.dct3d_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != _lib.OK:
            raise Dct3dError(rc, (self.L.dct3d_last_error(self.h) or b"").decode())

    # -- options / stats --------------------------------------------------------------------
    def set_option(self, key: str, value: int):
        self._check(self.L.dct3d_set_option(self.h, key.encode(), value))

    def stat(self, key: str) -> int:
        return self.L.dct3d_get_stat(self.h, key.encode())

    @property
    def cube_size(self) -> int:
        return self.cube ** 3

    def _nframes_eff(self, nframes: int) -> int:
        return nframes - nframes % self.cube

    # -- fused path, host buffers -------------------------------------------------------------
    def encode_u8(self, frames: np.ndarray, cap: int | None = None):
        """u8 frames [F][H][W] -> (stream bytes as np.uint8 of floor(bits/8)+1, nbits)."""
        fr = np.ascontiguousarray(frames, np.uint8)
        F = fr.shape[0] if fr.ndim == 3 else fr.size // (self.width * self.height)
        cap = cap or (fr.size // 2 + 4096)
        while True:
            out = np.zeros(cap, np.uint8)
            nbits, nbytes = C.c_uint64(), C.c_size_t()
            rc = self.L.dct3d_encode_u8(self.h, _ptr(fr), F, _ptr(out), cap, C.byref(nbits), C.byref(nbytes))
            if rc == _lib.E_OVERFLOW and cap < 4 * fr.size + 4096:
                cap = min(cap * 4, 4 * fr.size + 4096)
                continue
            self._check(rc)
            return out[: nbytes.value], nbits.value

    def decode_u8(self, stream, nframes: int) -> np.ndarray:
        s = np.ascontiguousarray(stream, np.uint8)
        Fe = self._nframes_eff(nframes)
        out = np.zeros((Fe, self.height, self.width), np.uint8)
        self._check(self.L.dct3d_decode_u8(self.h, _ptr(s), s.size, nframes, _ptr(out)))
        return out

    # -- sharded coding: a slab range per GPU / process, placed into the clip's one stream -------------
    def encode_u8_range(self, frames) -> int:
        """Phase 1: code the frames from bit 0 into the context's device buffer; returns the bit count."""
        fr = frames if not isinstance(frames, np.ndarray) else np.ascontiguousarray(frames, np.uint8)
        F = (fr.size if isinstance(fr, np.ndarray) else fr.numel()) // (self.width * self.height)
        nbits = C.c_uint64()
        self._check(self.L.dct3d_encode_u8_range(self.h, _ptr(fr), F, C.byref(nbits)))
        return nbits.value

    def encode_u8_place(self, start_bit: int, last: bool, stream, cap: int | None = None) -> int:
        """Phase 2: move the coded range to global bit `start_bit` of the host buffer `stream`; returns the range's
        first byte when it is shared with the predecessor (to be OR-ed in by the caller), else 0."""
        fb = C.c_uint8(0)
        cap = cap if cap is not None else (stream.size if isinstance(stream, np.ndarray) else stream.numel())
        self._check(self.L.dct3d_encode_u8_place(self.h, start_bit, 1 if last else 0, _ptr(stream), cap, C.byref(fb)))
        return fb.value

    def decode_u8_range(self, stream, start_bit: int, nframes: int, end_bit_hint: int = 0, out=None, nbytes: int | None = None):
        """Decode `nframes` frames whose first code starts at bit `start_bit` of `stream`; returns (frames, end bit)."""
        s = stream if not isinstance(stream, np.ndarray) else np.ascontiguousarray(stream, np.uint8)
        nbytes = nbytes if nbytes is not None else (s.size if isinstance(s, np.ndarray) else s.numel())
        if out is None:
            out = np.zeros((self._nframes_eff(nframes), self.height, self.width), np.uint8)
        end = C.c_uint64()
        self._check(self.L.dct3d_decode_u8_range(self.h, _ptr(s), nbytes, start_bit, end_bit_hint, nframes, _ptr(out), C.byref(end)))
        return out, end.value

    def stream_shift_dev(self, d_src, nbits: int, phase: int, d_dst, cap: int, stream=0):
        self._check(self.L.dct3d_stream_shift_dev(self.h, _ptr(d_src), nbits, phase, _ptr(d_dst), cap, stream))

    # -- streaming ------------------------------------------------------------------------------
    def stream_begin(self):
        self._check(self.L.dct3d_stream_begin(self.h))

    def stream_encode(self, frames: np.ndarray, last: bool) -> np.ndarray:
        fr = np.ascontiguousarray(frames, np.uint8)
        F = fr.size // (self.width * self.height)
        cap = fr.size // 2 + 4096
        while True:
            out = np.zeros(cap, np.uint8)
            n = C.c_size_t()
            # a failed call leaves the carried state untouched, so retrying with a larger buffer is safe
            rc = self.L.dct3d_stream_encode(self.h, _ptr(fr), F, 1 if last else 0, _ptr(out), cap, C.byref(n))
            if rc == _lib.E_OVERFLOW and cap < 4 * fr.size + 4096:
                cap = min(cap * 4, 4 * fr.size + 4096)
                continue
            self._check(rc)
            return out[: n.value]

    def stream_decode(self, buf, bitpos: int, nframes: int):
        """-> (frames, new bitpos) or None when more input is needed."""
        s = np.ascontiguousarray(buf, np.uint8)
        out = np.zeros((self._nframes_eff(nframes), self.height, self.width), np.uint8)
        bp = C.c_uint64(bitpos)
        rc = self.L.dct3d_stream_decode(self.h, _ptr(s), s.size, C.byref(bp), nframes, _ptr(out))
        if rc == _lib.E_NEED_MORE:
            return None
        self._check(rc)
        return out, bp.value

    # -- transform seams ---------------------------------------------------------------------------
    def forward_f32(self, cubes: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(cubes, np.float32)
        nslabs = a.size // (self.width * self.height * self.cube)
        out = np.zeros_like(a)
        self._check(self.L.dct3d_forward_f32(self.h, _ptr(a), _ptr(out), nslabs))
        return out

    def inverse_f32(self, coef: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(coef, np.float32)
        nslabs = a.size // (self.width * self.height * self.cube)
        out = np.zeros_like(a)
        self._check(self.L.dct3d_inverse_f32(self.h, _ptr(a), _ptr(out), nslabs))
        return out

    def forward_f64(self, planar: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(planar, np.float64)
        F = a.size // (self.width * self.height)
        out = np.zeros_like(a)
        self._check(self.L.dct3d_forward_f64(self.h, _ptr(a), _ptr(out), F))
        return out

    def inverse_f64(self, planar: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(planar, np.float64)
        F = a.size // (self.width * self.height)
        out = np.zeros_like(a)
        self._check(self.L.dct3d_inverse_f64(self.h, _ptr(a), _ptr(out), F))
        return out

    # -- stages --------------------------------------------------------------------------------------
    def quantize_u8(self, frames: np.ndarray) -> np.ndarray:
        fr = np.ascontiguousarray(frames, np.uint8)
        F = fr.size // (self.width * self.height)
        Fe = self._nframes_eff(F)
        q = np.zeros(self.width * self.height * Fe, np.int16)
        self._check(self.L.dct3d_quantize_u8(self.h, _ptr(fr), F, _ptr(q)))
        return q.reshape(-1, self.cube, self.cube, self.cube)

    def reconstruct_i16(self, qcubes: np.ndarray, nframes: int) -> np.ndarray:
        q = np.ascontiguousarray(qcubes, np.int16)
        out = np.zeros((self._nframes_eff(nframes), self.height, self.width), np.uint8)
        self._check(self.L.dct3d_reconstruct_i16(self.h, _ptr(q), nframes, _ptr(out)))
        return out

    def eg_encode_i16(self, qcubes: np.ndarray, start_bit: int = 0, prefix: np.ndarray | None = None, cap: int | None = None):
        """int16 cubes -> (stream bytes, end_bit).  `prefix` supplies the bytes before/at start_bit."""
        q = np.ascontiguousarray(qcubes, np.int16)
        ncubes = q.size // self.cube_size
        cap = cap or (start_bit // 8 + 5 * q.size + 64)
        out = np.zeros(cap, np.uint8)
        if prefix is not None:
            out[: len(prefix)] = prefix
        end = C.c_uint64()
        self._check(self.L.dct3d_eg_encode_i16(self.h, _ptr(q), ncubes, start_bit, _ptr(out), cap, C.byref(end)))
        return out[: end.value // 8 + 1], end.value

    def eg_decode_i16(self, stream, ncubes: int, start_bit: int = 0):
        s = np.ascontiguousarray(stream, np.uint8)
        q = np.zeros(ncubes * self.cube_size, np.int16)
        end = C.c_uint64()
        self._check(self.L.dct3d_eg_decode_i16(self.h, _ptr(s), s.size, start_bit, ncubes, _ptr(q), C.byref(end)))
        return q.reshape(-1, self.cube, self.cube, self.cube), end.value

    def eg_locate(self, stream, ncubes: int, start_bit: int = 0) -> int:
        """Bit position right after the first `ncubes` cubes of the stream (index discovery only, nothing is decoded)."""
        s = np.ascontiguousarray(stream, np.uint8)
        end = C.c_uint64()
        self._check(self.L.dct3d_eg_locate(self.h, _ptr(s), s.size, start_bit, ncubes, C.byref(end)))
        return end.value

    # -- device-resident (torch tensors or raw device addresses) ------------------------------------------
    def encode_u8_dev(self, d_frames, nframes: int, d_stream, cap: int, start_bit: int = 0, stream=0, want_end=True):
        end = C.c_uint64()
        self._check(self.L.dct3d_encode_u8_dev(self.h, _ptr(d_frames), nframes, _ptr(d_stream), cap, start_bit,
                                               C.byref(end) if want_end else None, stream))
        return end.value if want_end else None

    def decode_u8_dev(self, d_stream, nbytes: int, nframes: int, d_frames, start_bit: int = 0, stream=0):
        end = C.c_uint64()
        self._check(self.L.dct3d_decode_u8_dev(self.h, _ptr(d_stream), nbytes, start_bit, nframes, _ptr(d_frames),
                                               C.byref(end), stream))
        return end.value

    def forward_f32_dev(self, d_in, d_out, nslabs: int, stream=0):
        self._check(self.L.dct3d_forward_f32_dev(self.h, _ptr(d_in), _ptr(d_out), nslabs, stream))

    def inverse_f32_dev(self, d_in, d_out, nslabs: int, stream=0):
        self._check(self.L.dct3d_inverse_f32_dev(self.h, _ptr(d_in), _ptr(d_out), nslabs, stream))

    def forward_f64_dev(self, d_in, d_out, nframes: int, stream=0):
        self._check(self.L.dct3d_forward_f64_dev(self.h, _ptr(d_in), _ptr(d_out), nframes, stream))

    def inverse_f64_dev(self, d_in, d_out, nframes: int, stream=0):
        self._check(self.L.dct3d_inverse_f64_dev(self.h, _ptr(d_in), _ptr(d_out), nframes, stream))

    def quantize_u8_dev(self, d_frames, nframes: int, d_q, stream=0):
        self._check(self.L.dct3d_quantize_u8_dev(self.h, _ptr(d_frames), nframes, _ptr(d_q), stream))

    def reconstruct_i16_dev(self, d_q, nframes: int, d_frames, stream=0):
        self._check(self.L.dct3d_reconstruct_i16_dev(self.h, _ptr(d_q), nframes, _ptr(d_frames), stream))


    # -- colour planes (J/RGBUtils.java:39-131) ------------------------------------------------------------
    def rgb_split(self, rgb):
        """Raw RGB24 bytes -> (r, g, b) planes: byte i belongs to plane i % 3 (RGBUtils.split)."""
        a = np.ascontiguousarray(rgb, np.uint8).reshape(-1)
        planes = [np.empty((a.size + 2 - p) // 3, np.uint8) for p in range(3)]
        self._check(self.L.dct3d_rgb_split(self.h, _ptr(a), a.size, _ptr(planes[0]), _ptr(planes[1]), _ptr(planes[2])))
        return planes

    def rgb_mix(self, r, g, b):
        """Three equal planes -> raw RGB24 bytes (RGBUtils.mix)."""
        r, g, b = (np.ascontiguousarray(x, np.uint8).reshape(-1) for x in (r, g, b))
        if not (r.size == g.size == b.size):
            raise ValueError("planes differ in size")
        out = np.empty(3 * r.size, np.uint8)
        self._check(self.L.dct3d_rgb_mix(self.h, _ptr(r), _ptr(g), _ptr(b), r.size, _ptr(out)))
        return out

    def rgb_split_dev(self, d_rgb, nbytes: int, d_r, d_g, d_b, stream=0):
        self._check(self.L.dct3d_rgb_split_dev(self.h, _ptr(d_rgb), nbytes, _ptr(d_r), _ptr(d_g), _ptr(d_b), stream))

    def rgb_mix_dev(self, d_r, d_g, d_b, npixels: int, d_rgb, stream=0):
        self._check(self.L.dct3d_rgb_mix_dev(self.h, _ptr(d_r), _ptr(d_g), _ptr(d_b), npixels, _ptr(d_rgb), stream))


class MultiCodec:
    """Several GPUs in one process (dct3d_multi_*): contiguous slab ranges, one stream."""

    def __init__(self, width: int, height: int, cube: int = 8, devices=None, ndevices: int | None = None):
        self.L = _lib.load()
        self.width, self.height, self.cube = width, height, cube
        devs = list(devices) if devices is not None else list(range(ndevices or 1))
        self.n = len(devs)
        arr = (C.c_int * self.n)(*devs)
        h = C.c_void_p()
        rc = self.L.dct3d_multi_create(C.byref(h), arr, self.n, width, height, cube)
        if rc != _lib.OK:
            raise Dct3dError(rc, (self.L.dct3d_last_error(None) or b"").decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.dct3d_multi_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != _lib.OK:
            raise Dct3dError(rc, (self.L.dct3d_multi_last_error(self.h) or b"").decode())

    def set_option(self, key: str, value: int):
        self._check(self.L.dct3d_multi_set_option(self.h, key.encode(), value))

    def encode_u8(self, frames, out=None, cap: int | None = None):
        """-> (stream bytes, nbits, start bit of every range [n+1])."""
        fr = frames if not isinstance(frames, np.ndarray) else np.ascontiguousarray(frames, np.uint8)
        size = fr.size if isinstance(fr, np.ndarray) else fr.numel()
        F = size // (self.width * self.height)
        if out is None:
            cap = cap or (size // 2 + 4096)
            out = np.zeros(cap, np.uint8)
        cap = cap or (out.size if isinstance(out, np.ndarray) else out.numel())
        nbits, nbytes = C.c_uint64(), C.c_size_t()
        starts = (C.c_uint64 * (self.n + 1))()
        self._check(self.L.dct3d_multi_encode_u8(self.h, _ptr(fr), F, _ptr(out), cap, C.byref(nbits), C.byref(nbytes), starts))
        return out[: nbytes.value], nbits.value, list(starts)

    def set_weights(self, weights=None):
        """Shares of the slabs per GPU (None = equal)."""
        arr = None if weights is None else (C.c_double * self.n)(*[float(w) for w in weights])
        self._check(self.L.dct3d_multi_set_weights(self.h, arr))

    def probe_links(self):
        """-> (h2d GB/s, d2h GB/s, proposed weights) per GPU, all GPUs copying at once."""
        up, down, w = ((C.c_double * self.n)() for _ in range(3))
        self._check(self.L.dct3d_multi_probe_links(self.h, up, down, w))
        return list(up), list(down), list(w)

    def stream_begin(self):
        self._check(self.L.dct3d_multi_stream_begin(self.h))

    def stream_encode(self, frames: np.ndarray, last: bool, cap: int | None = None) -> np.ndarray:
        fr = np.ascontiguousarray(frames, np.uint8)
        F = fr.size // (self.width * self.height)
        cap = cap or (fr.size // 2 + 4096)
        out = np.zeros(cap, np.uint8)
        n = C.c_size_t()
        self._check(self.L.dct3d_multi_stream_encode(self.h, _ptr(fr), F, 1 if last else 0, _ptr(out), cap, C.byref(n)))
        return out[: n.value]

    def stream_decode(self, buf, bitpos: int, nframes: int):
        """-> (frames, new bitpos) or None when more input is needed."""
        s = np.ascontiguousarray(buf, np.uint8)
        out = np.zeros((nframes - nframes % self.cube, self.height, self.width), np.uint8)
        bp = C.c_uint64(bitpos)
        rc = self.L.dct3d_multi_stream_decode(self.h, _ptr(s), s.size, C.byref(bp), nframes, _ptr(out))
        if rc == _lib.E_NEED_MORE:
            return None
        self._check(rc)
        return out, bp.value

    def locate(self, stream, nframes: int):
        s = stream if not isinstance(stream, np.ndarray) else np.ascontiguousarray(stream, np.uint8)
        starts = (C.c_uint64 * (self.n + 1))()
        self._check(self.L.dct3d_multi_locate(self.h, _ptr(s), s.size if isinstance(s, np.ndarray) else s.numel(), nframes, starts))
        return list(starts)

    def decode_u8(self, stream, nframes: int, range_start_bits=None, out=None):
        s = stream if not isinstance(stream, np.ndarray) else np.ascontiguousarray(stream, np.uint8)
        nbytes = s.size if isinstance(s, np.ndarray) else s.numel()
        if out is None:
            out = np.zeros((nframes - nframes % self.cube, self.height, self.width), np.uint8)
        sb = None
        if range_start_bits is not None:
            sb = (C.c_uint64 * (self.n + 1))(*[int(x) for x in range_start_bits])
        self._check(self.L.dct3d_multi_decode_u8(self.h, _ptr(s), nbytes, nframes, _ptr(out), sb))
        return out


def list_devices() -> str:
    L = _lib.load()
    buf = C.create_string_buffer(4096)
    n = L.dct3d_list_devices(buf, 4096)
    if n < 0:
        raise Dct3dError(n, (L.dct3d_last_error(None) or b"").decode())
    return buf.value.decode()


# --------------------------------------------------------------------------------------------------
# Java-shaped transform objects (J/dct/Transform.java:44-65)
# --------------------------------------------------------------------------------------------------
class _Transform:
    _inverse = False

    def __init__(self, input, output, frameWidth, frameHeight, cubeWidth=8, cubeHeight=8, cubeDepth=8, device=0):
        if not (cubeWidth == cubeHeight == cubeDepth):
            raise ValueError("only cubic blocks are supported (8x8x8 or 4x4x4)")
        self.input, self.output = input, output
        self.frameWidth, self.frameHeight, self.cube = frameWidth, frameHeight, cubeWidth
        self.device = device

    def run(self, threads=None):
        """Blocks until the transform of every cube is in ``output`` (J/dct/Transform.java:63-104)."""
        a = np.ascontiguousarray(self.input, np.float64).reshape(-1)
        with Codec(self.frameWidth, self.frameHeight, self.cube, self.device) as c:
            res = c.inverse_f64(a) if self._inverse else c.forward_f64(a)
        np.copyto(np.asarray(self.output).reshape(-1)[: res.size], res.reshape(-1))


class DCT(_Transform):
    """``new DCT(pixels, dctCoeff, w, h, 8, 8, 8).run()`` (J/Encoder.java:63-64)."""


class InverseDCT(_Transform):
    """``new InverseDCT(dct, pixels, w, h, 8, 8, 8).run()`` (J/Decoder.java:102-103); clamps to [0,255]."""
    _inverse = True


# --------------------------------------------------------------------------------------------------
# Command lines
# --------------------------------------------------------------------------------------------------
class Encoder:
    """``java br.jpiccoli.video.Encoder <in> <out> <w> <h> <frames>`` (J/Encoder.java:14-129)."""

    @staticmethod
    def main(args, cube: int = 8, device: int = 0) -> int:
        if len(args) < 5:   # the reference guards on 4 but reads args[4] unconditionally (J/Encoder.java:16,32-33)
            print("Usage: Encoder <input file> <output file> <frame width> <frame height> <number of frames to encode>")
            return -1
        width, height, depth = int(args[2]), int(args[3]), int(args[4])
        depth -= depth % cube                                              # J/Encoder.java:39-40
        raw = np.fromfile(args[0], np.uint8, count=width * height * depth)
        if raw.size < width * height * depth:
            raise EOFError("input file shorter than width*height*frames")  # DataInputStream.readFully
        with Codec(width, height, cube, device) as c:
            stream, _bits = c.encode_u8(raw.reshape(depth, height, width))
        with open(args[1], "wb") as f:
            f.write(zlib.compress(stream.tobytes()))                       # Deflater() default level, :116-123
        print("Finished. Frames encoded: %d" % depth)
        return 0


class Decoder:
    """``java br.jpiccoli.video.Decoder <in> <out> <w> <h> <frames>`` (J/Decoder.java:15-121)."""

    @staticmethod
    def main(args, cube: int = 8, device: int = 0) -> int:
        if len(args) < 5:
            print("Usage: Decoder <input file> <output file> <frame width> <frame height> <number of frames to decode>")
            return -1
        width, height, depth = int(args[2]), int(args[3]), int(args[4])
        depth -= depth % cube                                              # J/Decoder.java:35-36
        stream = np.frombuffer(zlib.decompress(open(args[0], "rb").read()), np.uint8)
        with Codec(width, height, cube, device) as c:
            frames = c.decode_u8(stream, depth)
        frames.tofile(args[1])
        print("Complete!")
        return 0


class RGBUtils:
    """`java RGBUtils split|mix ...` (J/RGBUtils.java:13-37): colour video is coded as three gray streams."""

    @staticmethod
    def main(args, device: int = 0) -> int:
        if len(args) < 3 or args[0] not in ("split", "mix"):
            print("Usage:\n\njava RGBUtils split <input file> <output files names pattern>\n"
                  "java RGBUtils mix <input files names pattern> <output file>")
            return 1
        names = (".red", ".green", ".blue")
        with Codec(8, 8, 8, device) as c:           # the plane kernels do not depend on the frame geometry
            if args[0] == "split":
                for p, ext in zip(c.rgb_split(np.fromfile(args[1], np.uint8)), names):
                    p.tofile(args[2] + ext)
            else:
                planes = [np.fromfile(args[1] + ext, np.uint8) for ext in names]
                n = planes[0].size                  # mix() takes the red file's length as the pixel count (:113-127)
                c.rgb_mix(*(np.resize(p, n) if p.size != n else p for p in planes)).tofile(args[2])
        return 0


def codec_main(argv, cube: int = 8) -> int:
    """``codec list_platforms`` / ``codec encode|decode <in> <out> <w> <h> <frames> [device]`` (C/main.c:5-49).
    The optional last argument is the 1-based GPU index, as the reference's platform index is."""
    usage = ("Usage\n\ncodec list_platforms -> List available CUDA devices\n"
             "codec encode|decode <input file> <output file> <width> <height> <nr of frames to encode/decode> "
             "<device_index (optional)> -> Encode/Decode given file")
    if len(argv) < 2:
        print(usage)
        return 0
    if argv[1] == "list_platforms":
        print(list_devices(), end="")
        return 0
    if len(argv) < 7 or argv[1] not in ("encode", "decode"):
        print(usage)
        return 0
    src, dst, width, height, frames = argv[2], argv[3], int(argv[4]), int(argv[5]), int(argv[6])
    device = int(argv[7]) - 1 if len(argv) > 7 else 0
    slab = width * height * cube
    with Codec(width, height, cube, device) as c:
        if argv[1] == "encode":
            z = zlib.compressobj(9)                                        # Z_BEST_COMPRESSION, C/encoder.c:139
            c.stream_begin()
            done = 0
            with open(src, "rb") as fi, open(dst, "wb") as fo:
                while done < frames:                                       # C/encoder.c:203
                    buf = np.frombuffer(fi.read(slab), np.uint8)
                    if buf.size < slab:                                    # the reference reads garbage here; we pad with zeros
                        buf = np.concatenate([buf, np.zeros(slab - buf.size, np.uint8)])
                    done += cube
                    last = done >= frames
                    out = c.stream_encode(buf, last)
                    fo.write(z.compress(out.tobytes()))
                    if last:
                        fo.write(z.flush())
                    print("Frames processed: %d" % done)
            print("Encoding process completed")
        else:
            d = zlib.decompressobj()
            pending = np.zeros(0, np.uint8)
            bitpos, done = 0, 0
            with open(src, "rb") as fi, open(dst, "wb") as fo:
                eof = False
                while done < frames:                                       # C/decoder.c:207
                    res = c.stream_decode(pending, bitpos, cube) if pending.size else None
                    if res is None:
                        chunk = fi.read(slab) if not eof else b""
                        if not chunk:
                            if eof:
                                raise Dct3dError(_lib.E_STREAM, "input ended before all frames were decoded")
                            eof = True
                            more = d.flush()
                        else:
                            more = d.decompress(chunk)
                        pending = np.concatenate([pending, np.frombuffer(more, np.uint8)])
                        continue
                    out, bitpos = res
                    out.tofile(fo)
                    pending = pending[bitpos // 8:]                        # expGolomb_freeBuffer(..., 0), C/decoder.c:233-235
                    bitpos %= 8
                    done += cube
                    print("Frames processed: %d" % done)
            print("Decoding process completed")
    return 0


if __name__ == "__main__":
    sys.exit(codec_main(sys.argv))
