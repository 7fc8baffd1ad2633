"""A/B of the pipelined decoder: stream uploaded whole and parsed once (piece_bytes >= stream) vs uploaded and parsed piece by
piece (default 8 MiB pieces, 2 MiB, 32 MiB).  ms per dct3d_decode_u8 of 1920x1080xF from page-locked host memory."""
import ctypes as C, importlib, json, sys, time
import torch
sys.path.insert(0, '.')
import bench
codec = importlib.import_module('3ddctvideoencoding_b200.codec')
W, H, F = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device('cuda', 0)
frames = bench.synth_slabs_torch(W, H, 8, 0, F // 8, 1, dev)
h_frames = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True); h_frames.copy_(frames)
h_out = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True)
cap = W * H * F // 2 + 4096
h_stream = torch.zeros(cap, dtype=torch.uint8, pin_memory=True)
c = codec.Codec(W, H, 8); L = c.L
nb, ny = C.c_uint64(), C.c_size_t()
assert L.dct3d_encode_u8(c.h, h_frames.data_ptr(), F, h_stream.data_ptr(), cap, C.byref(nb), C.byref(ny)) == 0
res = {}
for label, piece in (("whole", 1 << 40), ("8MiB", 8 << 20), ("2MiB", 2 << 20), ("32MiB", 32 << 20), ("whole_again", 1 << 40)):
    c.set_option("piece_bytes", piece)
    for _ in range(2):
        assert L.dct3d_decode_u8(c.h, h_stream.data_ptr(), ny.value, F, h_out.data_ptr()) == 0
    t = time.perf_counter()
    for _ in range(8):
        assert L.dct3d_decode_u8(c.h, h_stream.data_ptr(), ny.value, F, h_out.data_ptr()) == 0
    res[label] = round((time.perf_counter() - t) / 8 * 1e3, 3)
print(json.dumps(res))
