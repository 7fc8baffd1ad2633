"""Why does a decode beside an encode not overlap?  Per-call wall times of dct3d_encode_u8 / dct3d_decode_u8 alone and
side by side (two contexts, two host threads), plus raw pinned copies in both directions at the same chunk size."""
import ctypes as C, importlib, json, sys, threading, time
import torch
sys.path.insert(0, '.')
import bench
codec = importlib.import_module('3ddctvideoencoding_b200.codec')
W, H, F = 1920, 1080, 256
dev = torch.device('cuda', 0)
frames = bench.synth_slabs_torch(W, H, 8, 0, F // 8, 1, dev)
h_frames = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True); h_frames.copy_(frames)
h_out = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True)
cap = W * H * F // 2 + 4096
h_stream = torch.zeros(cap, dtype=torch.uint8, pin_memory=True)
ce, cd = codec.Codec(W, H, 8), codec.Codec(W, H, 8)
L = ce.L
nb, ny = C.c_uint64(), C.c_size_t()
def enc():
    t = time.perf_counter(); assert L.dct3d_encode_u8(ce.h, h_frames.data_ptr(), F, h_stream.data_ptr(), cap, C.byref(nb), C.byref(ny)) == 0; return t, time.perf_counter()
def dec():
    t = time.perf_counter(); assert L.dct3d_decode_u8(cd.h, h_stream.data_ptr(), ny.value, F, h_out.data_ptr()) == 0; return t, time.perf_counter()
enc(); dec(); enc(); dec()
res = {}
res["enc_alone_ms"] = [round((b - a) * 1e3, 2) for a, b in (enc() for _ in range(3))]
res["dec_alone_ms"] = [round((b - a) * 1e3, 2) for a, b in (dec() for _ in range(3))]
te, td = [], []
t0 = time.perf_counter()
th = [threading.Thread(target=lambda: te.extend(enc() for _ in range(4))), threading.Thread(target=lambda: td.extend(dec() for _ in range(4)))]
for t in th: t.start()
for t in th: t.join()
res["both_total_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
res["enc_beside_ms"] = [(round((a - t0) * 1e3, 1), round((b - t0) * 1e3, 1)) for a, b in te]
res["dec_beside_ms"] = [(round((a - t0) * 1e3, 1), round((b - t0) * 1e3, 1)) for a, b in td]
# raw copies
n = 32 << 20
hs = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
ds = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, reps=16):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): ds[0].copy_(hs[0], non_blocking=True)
        if down:
            with torch.cuda.stream(s2): hs[1].copy_(ds[1], non_blocking=True)
    torch.cuda.synchronize(); return n * reps / (time.perf_counter() - t) / 1e9
run(1, 1)
res["raw_GBps"] = {"h2d": round(run(1, 0), 1), "d2h": round(run(0, 1), 1), "both_each_way": round(run(1, 1), 1)}
# the same from two host threads
def cp(up, out):
    st = torch.cuda.Stream()
    t = time.perf_counter()
    with torch.cuda.stream(st):
        for _ in range(16):
            (ds[0].copy_(hs[0], non_blocking=True) if up else hs[1].copy_(ds[1], non_blocking=True))
    st.synchronize(); out.append(n * 16 / (time.perf_counter() - t) / 1e9)
o1, o2 = [], []
th = [threading.Thread(target=cp, args=(1, o1)), threading.Thread(target=cp, args=(0, o2))]
for t in th: t.start()
for t in th: t.join()
res["raw_two_threads_GBps"] = {"h2d": round(o1[0], 1), "d2h": round(o2[0], 1)}
print(json.dumps(res))
