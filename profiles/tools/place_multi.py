"""N processes, one GPU each: every process codes the same 128-frame range, then all place their ranges AT THE SAME TIME into
(a) a private cudaHostAlloc buffer, (b) one cudaHostRegister-ed /dev/shm mapping (disjoint offsets), (c) as (b) but one rank
after the other.  Prints ms per placement and rank.   python profiles/tools/place_multi.py N"""
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
    W, H, F = 1920, 1080, 128
    frames = bench.synth_slabs_torch(W, H, 8, 0, F // 8, 1, dev)
    h_frames = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True); h_frames.copy_(frames)
    c = codec.Codec(W, H, 8, device=rank); L = c.L
    nb = C.c_uint64(); fb = C.c_uint8()
    assert L.dct3d_encode_u8_range(c.h, h_frames.data_ptr(), F, C.byref(nb)) == 0
    n = nb.value // 8 + 2
    slot = (n + 4096 + 4095) & ~4095
    xch = sh.ShmExchange("dct3d_pm_x_%s" % tag, world, rank, create=False)
    shm = sh.SharedStream("dct3d_pm_s_%s" % tag, world * slot + 4096, create=False)
    assert L.dct3d_host_register(shm.array.ctypes.data, shm.nbytes) == 0
    private = torch.zeros(slot + 4096, dtype=torch.uint8, pin_memory=True)
    start = (rank * slot + 5) * 8 + 3                       # an odd byte, phase 3

    def timed(dst_ptr, cap, sbit, serial=False):
        ts = []
        for it in range(6):
            xch.all_gather(it)
            if serial:
                for r in range(world):
                    if r == rank:
                        t = time.perf_counter()
                        assert L.dct3d_encode_u8_place(c.h, sbit, 0, dst_ptr, cap, C.byref(fb)) == 0
                        ts.append(time.perf_counter() - t)
                    xch.all_gather(100 + r)
            else:
                t = time.perf_counter()
                assert L.dct3d_encode_u8_place(c.h, sbit, 0, dst_ptr, cap, C.byref(fb)) == 0
                ts.append(time.perf_counter() - t)
        return round(float(np.median(ts[1:])) * 1e3, 3)
    res = {"rank": rank, "bytes": n,
           "private_together_ms": timed(private.data_ptr(), slot + 4096, 5 * 8 + 3),
           "shm_together_ms": timed(shm.array.ctypes.data, shm.nbytes, start),
           "shm_one_after_the_other_ms": timed(shm.array.ctypes.data, shm.nbytes, start, serial=True)}
    # plain copies of the same size, together: private pinned and registered shm
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    hp = private[:n]
    hs = torch.from_numpy(shm.array[rank * slot:rank * slot + n])
    for name, dst in (("copy_private_together_ms", hp), ("copy_shm_together_ms", hs)):
        ts = []
        for it in range(6):
            xch.all_gather(it)
            torch.cuda.synchronize(); t = time.perf_counter(); dst.copy_(d, non_blocking=True); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
        res[name] = round(float(np.median(ts[1:])) * 1e3, 3)
    L.dct3d_host_unregister(shm.array.ctypes.data)
    print("RESULT " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        worker(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]); sys.exit(0)
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    sh = importlib.import_module('3ddctvideoencoding_b200.sharding')
    tag = str(os.getpid())
    xch = sh.ShmExchange("dct3d_pm_x_%s" % tag, world, 0, create=True)
    shm = sh.SharedStream("dct3d_pm_s_%s" % tag, world * (40 << 20) + 4096, create=True)
    ps = [subprocess.Popen([sys.executable, __file__, "--worker", str(r), str(world), tag], stdout=subprocess.PIPE, text=True) for r in range(world)]
    for p in ps:
        out, _ = p.communicate(timeout=600)
        for l in out.splitlines():
            if l.startswith("RESULT "): print(l[7:])
    xch.unlink(); shm.unlink()
