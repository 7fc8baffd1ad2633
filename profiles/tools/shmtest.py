import mmap, os, time, numpy as np
n = 32 << 20
src = np.random.randint(0, 255, n, dtype=np.uint8)
def bw(dst, label, reps=10):
    dst[:n] = src
    t = time.perf_counter()
    for _ in range(reps): dst[:n] = src
    dt = (time.perf_counter() - t) / reps
    t = time.perf_counter()
    for _ in range(reps): src[:] = dst[:n]
    dr = (time.perf_counter() - t) / reps
    print(f"{label:24s} write {n/dt/1e9:6.2f} GB/s  read {n/dr/1e9:6.2f} GB/s")
bw(np.empty(n, np.uint8), "numpy heap")
fd = os.open("/dev/shm/t_bw", os.O_RDWR | os.O_CREAT, 0o600); os.ftruncate(fd, 1 << 30)
mm = mmap.mmap(fd, 1 << 30); a = np.frombuffer(mm, np.uint8)
bw(a, "/dev/shm mmap (1 GiB file)")
bw(a[512 << 20:], "/dev/shm mmap offset 512M")
am = mmap.mmap(-1, 1 << 30); bw(np.frombuffer(am, np.uint8), "anonymous shared mmap")
os.unlink("/dev/shm/t_bw")
import subprocess; print(subprocess.run("df -h /dev/shm; cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/shmem_enabled", shell=True, capture_output=True, text=True).stdout)
