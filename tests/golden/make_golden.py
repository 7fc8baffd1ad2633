"""Generate tests/golden/* from the REFERENCE'S OWN C code (oracle/_ref, built by
oracle/Makefile from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

The reference ships no golden vectors, so these are outputs of its
unmodified CubeUtils.c / ExpGolomb.c / encoder.c / decoder.c (+ main.c for the
whole flow, through the CPU OpenCL shim oracle/ref_shim.c).  The zig-zag and
Exp-Golomb vectors are exact goldens; the whole-flow arrays depend on the shim's
cosf and are used with the tolerances stated in the tests.
"""
import importlib
import json
import os
import subprocess
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

synth = importlib.import_module("3ddctvideoencoding_b200.synth")
HERE = os.path.dirname(os.path.abspath(__file__))


def fnv1a32(a):
    h = 0x811C9DC5
    for b in np.asarray(a, "<i4").tobytes():
        h = ((h ^ b) * 0x01000193) & 0xFFFFFFFF
    return "%08x" % h


def main():
    O.build()
    assert O.ref() is not None, "oracle/_ref not built (needs /root/reference)"
    kat = {}
    for c in (8, 4):
        zz = O.ref_zigzag(c)
        sizes = np.bincount([(i % c) + (i // c) % c + i // (c * c) for i in zz.tolist()]).tolist()
        kat["zigzag%d" % c] = {"linear": zz.tolist(), "fnv1a32": fnv1a32(zz), "slice_sizes": sizes}
    singles = [0, 1, -1, 2, -2, 3, -3, 4, -4, 7, -7, 8, 100, -100, 2900, -2900, 5770, -5770, 32767, -32768]
    kat["eg_single"] = []
    for v in singles:
        b, bits = O.ref_eg_write([v])
        kat["eg_single"].append({"v": v, "bits": bits, "hex": b.tobytes().hex()})
    seq = [363, -12, 5, 0, 0, 1, -1, 0, 0, 0, 2, 0, 0, 0, 0, 0]
    b, bits = O.ref_eg_write(seq)
    kat["eg_sequence"] = {"values": seq, "bits": bits, "hex": b.tobytes().hex()}
    rng = np.random.default_rng(7)
    mag = rng.choice([0, 0, 0, 0, 0, 0, 1, 2, 5, 40, 700, 5770], size=3000)
    vals = (rng.integers(-1, 2, size=3000) * rng.integers(0, mag + 1)).astype(np.int32)
    b, bits = O.ref_eg_write(vals.tolist())
    back = O.ref_eg_read(b, vals.size)
    assert (back == vals).all()
    # quantiser rounding, C flavour (C/encoder.c:47-58, C/decoder.c:48-59), including exact ties
    coef = (rng.normal(0, 300, size=8 * 512)).astype(np.float32)
    coef[:64] = np.array([2.5, -2.5, 0.5, -0.5, 7.5, -7.5, 12.5, -12.5] * 8, np.float32)
    qc = O.ref_quantize_f32(coef)
    dq = O.ref_dequantize_f32(qc)
    json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), indent=1)

    # whole reference C flow on the App. C clip (64x48x24, natural, seed 1) through the reference CLI
    W, H, F = 64, 48, 24
    clip = synth.natural(W, H, F, 1)
    d = tempfile.mkdtemp()
    raw, enc, dec = (os.path.join(d, n) for n in ("in.raw", "out.enc", "out.raw"))
    clip.tofile(raw)
    exe = os.path.join(ROOT, "oracle", "_ref", "codec_ref")
    subprocess.run([exe, "encode", raw, enc, str(W), str(H), str(F)], stdout=subprocess.DEVNULL, check=True)
    subprocess.run([exe, "decode", enc, dec, str(W), str(H), str(F)], stdout=subprocess.DEVNULL, check=True)
    stream = np.frombuffer(zlib.decompress(open(enc, "rb").read()), np.uint8)
    decoded = np.fromfile(dec, np.uint8).reshape(F, H, W)
    ncubes = W * H * F // 512
    q, end = O.eg_decode_cubes(stream, ncubes, 8)
    offs = [0]
    for _ in range(F // 8):
        offs.append(O.eg_decode_cubes(stream, ncubes // (F // 8), 8, offs[-1])[1])
    np.savez_compressed(
        os.path.join(HERE, "ref_golden.npz"),
        eg_values=vals, eg_bytes=b, eg_bits=np.int64(bits),
        quant_in=coef, quant_out=qc, dequant_out=dq,
        flow_clip_sha=np.frombuffer(__import__("hashlib").sha256(clip.tobytes()).digest(), np.uint8),
        flow_stream=stream, flow_bits=np.int64(end), flow_slab_bit_offsets=np.array(offs, np.int64),
        flow_qcubes=q.astype(np.int16), flow_decoded=decoded,
    )
    print("wrote kat.json, ref_golden.npz:", end, "bits,", stream.size, "bytes, slab offsets", offs)


if __name__ == "__main__":
    main()
