"""Where does the placement of a coded range spend its time?  dct3d_encode_u8_place into a cudaHostAlloc buffer, into a
cudaHostRegister-ed /dev/shm mapping, and the host memcpy from the former into an unregistered /dev/shm mapping."""
import ctypes as C, importlib, json, mmap, os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import bench
codec = importlib.import_module('3ddctvideoencoding_b200.codec')
W, H, F = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device('cuda', 0)
frames = bench.synth_slabs_torch(W, H, 8, 0, F // 8, 1, dev)
h_frames = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True); h_frames.copy_(frames)
c = codec.Codec(W, H, 8); L = c.L
nb = C.c_uint64(); fb = C.c_uint8()
assert L.dct3d_encode_u8_range(c.h, h_frames.data_ptr(), F, C.byref(nb)) == 0
n = nb.value // 8 + 2
cap = n + 4096
pinned = torch.zeros(cap, dtype=torch.uint8, pin_memory=True)
fd = os.open("/dev/shm/t_place", os.O_RDWR | os.O_CREAT, 0o600); os.ftruncate(fd, cap)
mm = mmap.mmap(fd, cap); shm = np.frombuffer(mm, np.uint8)
shm[:] = 0
def t(fn, reps=5):
    fn(); a = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - a) / reps * 1e3
res = {"range_bytes": n}
for phase in (0, 3):
    res[f"place_pinned_phase{phase}_ms"] = t(lambda: L.dct3d_encode_u8_place(c.h, phase, 1, pinned.data_ptr(), cap, C.byref(fb)))
# destinations whose alignment differs from the source's (the range's first byte lands at an arbitrary byte of the stream)
big = torch.zeros(cap + (1 << 20), dtype=torch.uint8, pin_memory=True)
for off in (0, 4, 1, 2, 3, 12345, 12346, 12347, 12348):
    sb = off * 8 + 3
    res[f"place_pinned_byte0={off}_ms"] = t(lambda: L.dct3d_encode_u8_place(c.h, sb, 1, big.data_ptr(), cap + (1 << 20), C.byref(fb)))
pn = pinned.numpy()
def cp(): shm[:n] = pn[:n]
res["memcpy_pinned_to_shm_ms"] = t(cp)
def cp2(): pn[:n] = shm[:n]
res["memcpy_shm_to_pinned_ms"] = t(cp2)
assert L.dct3d_host_register(shm.ctypes.data, cap) == 0
res["place_registered_shm_phase3_ms"] = t(lambda: L.dct3d_encode_u8_place(c.h, 3, 1, shm.ctypes.data, cap, C.byref(fb)))
L.dct3d_host_unregister(shm.ctypes.data)
os.unlink("/dev/shm/t_place")
print(json.dumps(res))
