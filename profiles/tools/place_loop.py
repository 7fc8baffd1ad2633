"""Bisecting why placement is slow inside bench.py's multi-rank e2e loop: N processes run range -> exchange -> place [-> decode]
in a loop, with the shared stream sized / laid out as bench.py does it.  python profiles/tools/place_loop.py N"""
import ctypes as C, importlib, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def worker(rank, world, tag):
    import numpy as np, torch
    import bench
    codec = importlib.import_module('3ddctvideoencoding_b200.codec')
    sh = importlib.import_module('3ddctvideoencoding_b200.sharding')
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    W, H, F = 1920, 1080, 256
    frames = bench.synth_slabs_torch(W, H, 8, rank * 32, rank * 32 + 32, 1, dev)
    h_frames = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True); h_frames.copy_(frames)
    h_out = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True)
    c = codec.Codec(W, H, 8, device=rank); L = c.L
    scap = W * H * F * world // 2 + 4096
    xch = sh.ShmExchange("dct3d_pl_x_%s" % tag, world, rank, create=False)
    shm = sh.SharedStream("dct3d_pl_s_%s" % tag, scap, create=False)
    assert L.dct3d_host_register(shm.array.ctypes.data, scap) == 0
    ptr = shm.array.ctypes.data
    nb = C.c_uint64(); fb = C.c_uint8(); end = C.c_uint64()
    out = {}
    for variant in ("range+place", "range+place+decode", "range+sleep+place+decode"):
        tp, tr, td = [], [], []
        for it in range(5):
            t0 = time.perf_counter()
            assert L.dct3d_encode_u8_range(c.h, h_frames.data_ptr(), F, C.byref(nb)) == 0
            t1 = time.perf_counter()
            o = sh.bit_offsets(xch.all_gather(nb.value))
            if "sleep" in variant:
                time.sleep(0.005)
            t2 = time.perf_counter()
            assert L.dct3d_encode_u8_place(c.h, o[rank], 1 if rank == world - 1 else 0, ptr, scap, C.byref(fb)) == 0
            t3 = time.perf_counter()
            xch.signal()
            if o[rank] % 8:
                xch.wait_for(rank - 1)
                shm.array[o[rank] // 8] |= fb.value
            if "decode" in variant:
                assert L.dct3d_decode_u8_range(c.h, ptr, o[-1] // 8 + 1, o[rank], o[rank + 1], F, h_out.data_ptr(), C.byref(end)) == 0
            t4 = time.perf_counter()
            xch.all_gather(0)
            tr.append(t1 - t0); tp.append(t3 - t2); td.append(t4 - t3)
        out[variant] = {"range_ms": round(float(np.median(tr[1:])) * 1e3, 2), "place_ms": round(float(np.median(tp[1:])) * 1e3, 2), "rest_ms": round(float(np.median(td[1:])) * 1e3, 2)}
    L.dct3d_host_unregister(ptr)
    print("RESULT " + json.dumps({"rank": rank, **out}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        worker(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]); sys.exit(0)
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    sh = importlib.import_module('3ddctvideoencoding_b200.sharding')
    tag = str(os.getpid())
    xch = sh.ShmExchange("dct3d_pl_x_%s" % tag, world, 0, create=True)
    shm = sh.SharedStream("dct3d_pl_s_%s" % tag, 1920 * 1080 * 256 * world // 2 + 4096, create=True)
    ps = [subprocess.Popen([sys.executable, __file__, "--worker", str(r), str(world), tag], stdout=subprocess.PIPE, text=True) for r in range(world)]
    for p in ps:
        out, _ = p.communicate(timeout=600)
        for l in out.splitlines():
            if l.startswith("RESULT "): print(l[7:])
    xch.unlink(); shm.unlink()
