"""One process, G GPUs (dct3d_multi_*): end-to-end encode and decode of ONE 1080p clip from page-locked host memory.

    python profiles/tools/multi_time.py [frames] [G G ...]     -> gpurun_out/multi_time.json, one line per G

Reports frames/s of dct3d_multi_encode_u8, of dct3d_multi_decode_u8 with the encoder's range start bits and without
(distributed index discovery), the time of dct3d_multi_locate alone, and the SHA-256 of the stream (the same for every G)."""
import ctypes as C, hashlib, importlib, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
import bench
codec = importlib.import_module('3ddctvideoencoding_b200.codec')
lib = importlib.import_module('3ddctvideoencoding_b200._lib').load()
W, H = 1920, 1080
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
Gs = [int(x) for x in sys.argv[2:]] or [1, 2, 4, 8]
ndev = torch.cuda.device_count()
N = W * H * F


def pinned(n):
    p = lib.dct3d_host_alloc(n)
    assert p
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n,)), p

frames, pf = pinned(N)
out, po = pinned(N)
cap = N // 2 + 4096
stream, ps = pinned(cap)
dev = torch.device('cuda', 0)
for s0 in range(0, F // 8, 16):
    s1 = min(F // 8, s0 + 16)
    t = bench.synth_slabs_torch(W, H, 8, s0, s1, 1, dev)
    frames[s0 * 8 * W * H:s1 * 8 * W * H] = t.cpu().numpy().reshape(-1)
del t
rows = []
for G in Gs:
    if G > ndev:
        continue
    with codec.MultiCodec(W, H, 8, devices=list(range(G))) as m:
        nbits, nbytes = C.c_uint64(), C.c_size_t()
        starts = (C.c_uint64 * (G + 1))()

        def enc():
            t0 = time.perf_counter()
            rc = lib.dct3d_multi_encode_u8(m.h, pf, F, ps, cap, C.byref(nbits), C.byref(nbytes), starts)
            assert rc == 0, lib.dct3d_multi_last_error(m.h)
            return time.perf_counter() - t0

        def dec(side):
            t0 = time.perf_counter()
            rc = lib.dct3d_multi_decode_u8(m.h, ps, nbytes.value, F, po, starts if side else None)
            assert rc == 0, lib.dct3d_multi_last_error(m.h)
            return time.perf_counter() - t0

        def loc():
            found = (C.c_uint64 * (G + 1))()
            t0 = time.perf_counter()
            rc = lib.dct3d_multi_locate(m.h, ps, nbytes.value, F, found)
            assert rc == 0, lib.dct3d_multi_last_error(m.h)
            dt = time.perf_counter() - t0
            assert list(found)[:G] == list(starts)[:G]
            return dt
        enc(); dec(True); dec(False)
        te = min(enc() for _ in range(3))
        sha = hashlib.sha256(stream[:nbytes.value].tobytes()).hexdigest()
        td = min(dec(True) for _ in range(3))
        ref = out.copy() if G == Gs[0] else ref
        same = bool((out == ref).all())
        tn = min(dec(False) for _ in range(3))
        same = same and bool((out == ref).all())
        tl = min(loc() for _ in range(3)) if G > 1 else 0.0
        balanced = None
        if G > 1:
            # link-aware shares: probe every GPU's copy rates (all copying at once), share the slabs out accordingly
            up, down, w = m.probe_links()
            m.set_weights(w)
            enc(); dec(True)
            be, bd = min(enc() for _ in range(3)), min(dec(True) for _ in range(3))
            balanced = {"h2d_gbs": [round(x, 1) for x in up], "d2h_gbs": [round(x, 1) for x in down], "encode_fps": F / be, "decode_fps_side_info": F / bd,
                        "stream_sha256_equal": hashlib.sha256(stream[:nbytes.value].tobytes()).hexdigest() == sha}
            m.set_weights(None)
        rows.append({"gpus": G, "frames": F, "balanced": balanced, "encode_fps": F / te, "decode_fps_side_info": F / td, "decode_fps_discovery": F / tn,
                     "locate_ms": tl * 1e3, "encode_ms": te * 1e3, "decode_ms": td * 1e3, "decode_discovery_ms": tn * 1e3,
                     "stream_bytes": nbytes.value, "stream_sha256": sha, "frames_equal_first_config": same})
        print(json.dumps(rows[-1]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/multi_time.json", "w"), indent=1)
