"""Do N processes writing disjoint ranges of ONE /dev/shm mapping slow each other down?  (No GPU involved.)
python profiles/tools/shm_contention.py N [MB per process]"""
import mmap, os, subprocess, sys, time
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "--worker":
    rank, world, mb, path = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    n = mb << 20
    fd = os.open(path, os.O_RDWR); mm = mmap.mmap(fd, world * n + 4096); shm = np.frombuffer(mm, np.uint8)
    src = np.random.randint(0, 255, n, dtype=np.uint8)
    private = np.empty(n + 64, np.uint8)
    flag = np.frombuffer(mm, np.int64, count=world, offset=world * n)   # crude barrier in the tail page
    def barrier(k):
        flag[rank] = k
        while (flag < k).any(): pass
    res = []
    for name, dst in (("private", private), ("shm_aligned", shm[rank * n:(rank + 1) * n]), ("shm_unaligned", shm[rank * n + 3:(rank + 1) * n])):
        m = min(n, dst.size) - 8
        dst[:m] = src[:m]
        ts = []
        for it in range(5):
            barrier(10 * (len(res) + 1) + it)
            t = time.perf_counter(); dst[:m] = src[:m]; ts.append(time.perf_counter() - t)
        res.append((name, m / np.median(ts) / 1e9))
    print(rank, " ".join(f"{a} {b:.1f} GB/s" for a, b in res), flush=True)
    sys.exit(0)
world = int(sys.argv[1]); mb = int(sys.argv[2]) if len(sys.argv) > 2 else 19
path = "/dev/shm/dct3d_contention"
fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o600); os.ftruncate(fd, world * (mb << 20) + 4096); os.close(fd)
ps = [subprocess.Popen([sys.executable, __file__, "--worker", str(r), str(world), str(mb), path]) for r in range(world)]
for p in ps: p.wait()
os.unlink(path)
