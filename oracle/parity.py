"""TEST INFRASTRUCTURE (not a product path): the four parity rules of BASELINE.json's north star, evaluated for one
piece of a clip against the CPU oracle.  Used by tests/ and by bench.py's cpu_baseline leg, which has the oracle's
results for its sample slabs at hand anyway.

Rules (reference lines in brackets):
  (1) float DCT coefficients within 1e-4 relative (to the block's magnitude) of the fp64 oracle
      [J/dct/DCT.java:41-59; the C flavour's float kernels C/3dDCT.cl:43-143];
  (2) quantised cubes identical except +-1 flips at rounding ties, counted [J/Encoder.java:82, C/encoder.c:53];
  (3) Exp-Golomb stream bit-exact given identical quantised cubes [J/ExpGolombWriter.java:19-49, C/ExpGolomb.c:32-64];
  (4) decoded pixels within +-1 [J/Decoder.java:107-117, C/decoder.c:29].
"""
from __future__ import annotations

import numpy as np

from . import oracle as O

NEAR_TIE = 2e-3     # |frac(v) - 0.5| below which an fp32 result may round the other way


def classify_flips(q, ref, coef_planar, cube):
    """Every mismatch must be a +-1 flip at a rounding tie of the fp64 value: an EXACT tie (the 4^3 transform has
    rational coefficients, e.g. DC = sum/8; the reference's own Java and C results differ there) or a NEAR tie within
    fp32 error.  Returns (exact, near); raises AssertionError on any other difference."""
    bad = np.argwhere(q != ref)
    if bad.size == 0:
        return 0, 0
    assert np.abs(q - ref).max() <= 1, "a quantised value differs from the oracle by more than one"
    cc = O.frames_to_cubes(coef_planar, cube)
    k = np.indices((cube, cube, cube)).sum(axis=0)
    div = np.maximum(1, 5 * k)[None]
    v = (cc / div)[tuple(bad.T)]
    dist = np.abs(np.abs(v - np.floor(v)) - 0.5)
    assert dist.max() < NEAR_TIE, "a quantised value differs from the oracle away from any rounding tie"
    exact = int((dist < 1e-9).sum())
    return exact, int(bad.shape[0] - exact)


def piece_parity(codec, frames: np.ndarray, cube: int) -> dict:
    """All four rules on `frames` (a whole number of slabs) through the codec object's C-ABI calls."""
    fr = np.ascontiguousarray(frames, np.uint8)
    F, H, W = fr.shape
    ref, coef = O.quantized_cubes(fr, cube, mode=0, want_coef=True)
    # (1) coefficients through the float seam (the reference's own device boundary)
    cubes = O.frames_to_cubes(fr, cube).astype(np.float32)
    got = codec.forward_f32(cubes).reshape(cubes.shape).astype(np.float64)
    want = O.frames_to_cubes(coef, cube)
    blockmax = np.abs(want).reshape(want.shape[0], -1).max(axis=1).reshape(-1, 1, 1, 1)
    coef_rel = float((np.abs(got - want) / np.maximum(blockmax, 1.0)).max())
    # (2) quantised cubes
    q = codec.quantize_u8(fr).astype(np.int32)
    exact, near = classify_flips(q, ref, coef, cube)
    # (3) the stream, given the cubes the GPU produced
    stream, nbits = codec.encode_u8(fr)
    want_stream, want_bits = O.eg_encode_cubes(q, cube, cap=5 * q.size + 64)
    stream_ok = bool(nbits == want_bits and stream.tobytes() == want_stream[: nbits // 8 + 1].tobytes())
    # (4) pixels
    dec = codec.decode_u8(stream, F)
    odec = O.decode_u8(stream, W, H, F, cube)
    pix = int(np.abs(dec.astype(np.int16) - odec.astype(np.int16)).max())
    return {"coefficients": int(q.size), "flips_exact": exact, "flips_near": near, "coef_max_rel": coef_rel,
            "stream_bit_exact": stream_ok, "pixel_max_abs": pix, "pixels_differing": int((dec != odec).sum()), "nbits": int(nbits)}


def merge(parts: list[dict]) -> dict:
    out = {"pieces_checked": len(parts)}
    for k in ("coefficients", "flips_exact", "flips_near", "pixels_differing"):
        out[k] = int(sum(p[k] for p in parts))
    out["coef_max_rel"] = float(max(p["coef_max_rel"] for p in parts))
    out["pixel_max_abs"] = int(max(p["pixel_max_abs"] for p in parts))
    out["stream_bit_exact"] = bool(all(p["stream_bit_exact"] for p in parts))
    return out
