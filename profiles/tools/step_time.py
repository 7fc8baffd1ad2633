"""Times encode+decode steps of 1920x1080xF (device-resident) for the library named by DCT3D_LIB (default: the in-tree
build) and prints one JSON line: ms per step, per-direction ms, the two transform kernels' device times, and hashes of
the stream and of the decoded frames (variants of one kernel must agree bit for bit)."""
import hashlib, importlib, json, os, sys
import torch
sys.path.insert(0, '.')
import bench
codec = importlib.import_module('3ddctvideoencoding_b200.codec')
F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cube = int(sys.argv[3]) if len(sys.argv) > 3 else 8
kind = sys.argv[4] if len(sys.argv) > 4 else "natural"
precision = int(sys.argv[5]) if len(sys.argv) > 5 else 32
W, H = 1920, 1080
dev = torch.device('cuda', 0)
frames = bench.synth_slabs_torch(W, H, cube, 0, F // cube, 1, dev, kind=kind)
cap = W * H * F * (4 if kind == "noise" else 1) // 2 + 4096
d_stream = torch.zeros(cap, dtype=torch.uint8, device=dev)
d_out = torch.empty_like(frames)
c = codec.Codec(W, H, cube)
c.set_option("reuse_zeroed", 1)
c.set_option("precision", precision)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
for i in range(5):
    end = c.encode_u8_dev(frames, F, d_stream, cap, 0, st)
    c.decode_u8_dev(d_stream, end // 8 + 1, F, d_out, 0, st)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
c.set_option("kernel_times_reset", 1)
te = td = 0.0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
evs = []
for i in range(steps):
    a, b, d = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    end = c.encode_u8_dev(frames, F, d_stream, cap, 0, st)
    b.record()
    c.decode_u8_dev(d_stream, end // 8 + 1, F, d_out, 0, st)
    d.record()
    evs.append((a, b, d))
e1.record()
torch.cuda.synchronize()
te = sum(a.elapsed_time(b) for a, b, d in evs) / steps
td = sum(b.elapsed_time(d) for a, b, d in evs) / steps
print(json.dumps({"lib": os.path.basename(os.environ.get("DCT3D_LIB", "in-tree")),
                  "opts": {k: c.stat(k) for k in ("zero_skip", "tma_store", "pack_sort", "tma_store_used")}, "frames": F, "cube": cube, "kind": kind, "precision": precision,
                  "ms_per_step": e0.elapsed_time(e1) / steps, "encode_ms": te, "decode_ms": td,
                  "encode_kernel_ms": c.stat("ns_encode_kernel_avg") * 1e-6, "reconstruct_kernel_ms": c.stat("ns_reconstruct_kernel_avg") * 1e-6,
                  "fps": F * steps / (e0.elapsed_time(e1) * 1e-3), "bits": end,
                  "stream_sha": hashlib.sha256(d_stream[: end // 8 + 1].cpu().numpy().tobytes()).hexdigest()[:16],
                  "frames_sha": hashlib.sha256(d_out.cpu().numpy().tobytes()).hexdigest()[:16]}))
