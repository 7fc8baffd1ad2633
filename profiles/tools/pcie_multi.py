"""Pure pinned host<->device copy ceiling of the box with N processes, one GPU each, no kernels:

    python profiles/tools/pcie_multi.py N [bind]      -> profiles/r2_pcie_nN.json

Every process copies 32 MiB chunks (the library's pipeline chunk size at 1080p) H2D only, D2H only and both at once; the
processes start each phase together (shared-memory all_gather).  `bind` pins every process to the CPUs of its GPU's NUMA
node before it allocates its pinned buffers (first touch), which is what bench.py does for its e2e path."""
import importlib, json, os, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def gpu_cpus(index):
    out = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    bdf = out.lower()
    if bdf.count(":") == 2 and len(bdf.split(":")[0]) == 8:
        bdf = bdf[4:]
    p = f"/sys/bus/pci/devices/{bdf}/local_cpulist"
    if not os.path.exists(p):
        return None, bdf
    cpus = set()
    for part in open(p).read().strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    try:
        node = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()
    except OSError:
        node = "?"
    return cpus, f"{bdf} numa {node}"


def worker(rank, world, bind, tag):
    import torch
    sh = importlib.import_module("3ddctvideoencoding_b200.sharding")
    torch.cuda.set_device(rank)
    info = {"rank": rank}
    if bind:
        cpus, bdf = gpu_cpus(rank)
        info["bdf"] = bdf
        if cpus:
            try:
                os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or os.sched_getaffinity(0))
                info["cpus"] = len(os.sched_getaffinity(0))
            except OSError:
                pass
    xch = sh.ShmExchange("dct3d_pcie_%s" % tag, world, rank, create=False)
    dev = torch.device("cuda", rank)
    n, reps = 32 << 20, 32
    hs = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    for h in hs:
        h.fill_(rank)
    ds = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, down):
        torch.cuda.synchronize()
        xch.all_gather(0)
        t = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s1):
                    ds[0].copy_(hs[0], non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    hs[1].copy_(ds[1], non_blocking=True)
        torch.cuda.synchronize()
        return n * reps / (time.perf_counter() - t) / 1e9
    run(1, 1)
    info["h2d"], info["d2h"], info["both_each_way"] = run(1, 0), run(0, 1), run(1, 1)
    print("RESULT " + json.dumps(info), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[1] == "--worker":
        worker(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4] == "1", sys.argv[5])
        sys.exit(0)
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    bind = len(sys.argv) > 2 and sys.argv[2] == "bind"
    sh = importlib.import_module("3ddctvideoencoding_b200.sharding")
    tag = str(os.getpid())
    xch = sh.ShmExchange("dct3d_pcie_%s" % tag, world, 0, create=True)
    procs = [subprocess.Popen([sys.executable, __file__, "--worker", str(r), str(world), "1" if bind else "0", tag], stdout=subprocess.PIPE, text=True)
             for r in range(world)]
    rows = []
    for p in procs:
        out, _ = p.communicate(timeout=600)
        rows += [json.loads(l[7:]) for l in out.splitlines() if l.startswith("RESULT ")]
    xch.unlink()
    res = {"processes": world, "numa_bound": bind, "chunk_bytes": 32 << 20, "per_rank": sorted(rows, key=lambda r: r["rank"]),
           "aggregate_GBps": {k: sum(r[k] for r in rows) for k in ("h2d", "d2h", "both_each_way")}}
    name = os.path.join(ROOT, "profiles", "r2_pcie_n%d%s.json" % (world, "_bound" if bind else ""))
    json.dump(res, open(name, "w"), indent=1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", os.path.basename(name)), "w"), indent=1)
    print(json.dumps(res["aggregate_GBps"]), "bound" if bind else "unbound", "N =", world)
