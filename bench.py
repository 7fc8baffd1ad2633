#!/usr/bin/env python
"""bench.py -- encode+decode throughput of the 3D-DCT codec hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--config c1|c2|c3|c4|c5]   # our CUDA path (libdct3d.so)
    python bench.py --impl reference --steps K --warmup W                     # the reference's CPU path (oracle port)

One step = one pass of the hot path (u8 frames -> Exp-Golomb stream -> u8 frames) over the configuration's clip:

    c2 (default) 1920x1080 gray, 256 frames PER GPU, 8x8x8 cubes   BASELINE configs[1]; weak scaling over N
    c1           640x480, 64 frames                                 configs[0], the reference's own CPU-runnable case
    c3           3840x2160, 1024 frames, sharded over the N GPUs    configs[2]; strong scaling
    c4           c2 with 4x4x4 cubes                                configs[3]
    c5           1920x1080, 4096 frames, sharded over the N GPUs    configs[4]; strong scaling

With N > 1 (torchrun, one process per GPU) the ranks code ONE clip: rank g owns the contiguous slab range g of it
(slabs are independent key-frame groups), codes it from bit 0, the N bit counts are all-gathered and prefix-summed, and
every rank moves its bits to its phase of the clip's ONE stream (stream_shift_kernel) and decodes its range from that
global bit position.  No data-path collective.  `value` is whole-job frames/s with the frames resident in HBM (CUDA
events on the launching stream, max over ranks; the all_gather of the N scalars and the shift are inside the timed
region); `e2e` is the same through the host-buffer C ABI with pinned host memory, H2D/D2H inside the timed region: the
ranks place their ranges straight into one shared-memory stream (dct3d_encode_u8_range / _place) and decode their
ranges from it (dct3d_decode_u8_range); at N = 1 these are dct3d_encode_u8 / dct3d_decode_u8.  The SHA-256 of that
stream is printed: it is the same for every N of a strong-scaling configuration.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "3ddctvideoencoding_b200"
UNIT = "frames/s"

CONFIGS = {
    #      W     H     frames cube  scaling   BASELINE.json configs[] index
    "c1": (640, 480, 64, 8, "strong", 0),
    "c2": (1920, 1080, 256, 8, "weak", 1),
    "c3": (3840, 2160, 1024, 8, "strong", 2),
    "c4": (1920, 1080, 256, 4, "weak", 3),
    "c5": (1920, 1080, 4096, 8, "strong", 4),
}


def metric_name(W, H):
    return "1080p gray encode+decode frames/s" if (W, H) == (1920, 1080) else f"{W}x{H} gray encode+decode frames/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json, written by
    profiles/summarize.py from the --set full export); None when no capture of that kernel is committed."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    for k, v in d.get("kernels", {}).items():
        if kernel.startswith(k):
            return float(v["dram_bytes"]), d.get("source")
    return None, d.get("source")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def synth_slabs_torch(W, H, cube, slab_lo, slab_hi, seed, device, kind="natural", out=None):
    """The SURVEY.md 8d 'natural' generator on the GPU, seeded PER SLAB: any rank regenerates any slab of the clip
    bit for bit, so the clip (and its stream) does not depend on how many ranks share it."""
    import torch
    n = slab_hi - slab_lo
    if out is None:
        out = torch.empty((n * cube, H, W), dtype=torch.uint8, device=device)
    y = torch.arange(H, device=device, dtype=torch.float32)[None, :, None]
    x = torch.arange(W, device=device, dtype=torch.float32)[None, None, :]
    g = torch.Generator(device=device)
    for s in range(slab_lo, slab_hi):
        g.manual_seed(seed * 1000003 + s)
        o = out[(s - slab_lo) * cube:(s - slab_lo + 1) * cube]
        if kind == "noise":
            o.copy_(torch.randint(0, 256, o.shape, generator=g, device=device, dtype=torch.uint8))
            continue
        t = torch.arange(s * cube, (s + 1) * cube, device=device, dtype=torch.float32)[:, None, None]
        v = 128.0 + 60.0 * torch.sin((x + 3.0 * t) / 37.0) + 50.0 * torch.cos((y - 2.0 * t) / 23.0)
        v = v + 6.0 * torch.randn(v.shape, generator=g, device=device)
        o.copy_(v.round().clamp(0, 255).to(torch.uint8))
    return out


def synth_clip_torch(W, H, F, seed, device, cube=8):
    """F frames of the natural clip on the GPU (whole slabs of `cube` frames)."""
    return synth_slabs_torch(W, H, cube, 0, F // cube, seed, device)


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's Java algorithm (the one place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------------------
def cpu_piece(clip, cube, threads):
    """Times the oracle's Java-structured port (oracle/dct3d_oracle.c) on `clip`; returns (t_enc, t_dec, stream, frames)."""
    from oracle import oracle as O
    F, H, W = clip.shape
    t0 = time.perf_counter()
    stream, bits = O.java_encode_u8(clip, cube, threads)
    t1 = time.perf_counter()
    dec = O.java_decode_u8(stream, W, H, F, cube, threads)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, stream, dec


def sample_slabs(nslabs: int, want: int = 4):
    """First, two in the middle, last (at least 4 slabs of the headline configuration)."""
    if nslabs <= want:
        return list(range(nslabs))
    return sorted({0, nslabs // 3, (2 * nslabs) // 3, nslabs - 1})


def ref_c_host_sample(W, H):
    """Times the reference's own C host code (encoder.c / decoder.c / ExpGolomb.c / CubeUtils.c compiled unmodified
    into oracle/_ref/codec_ref, its four OpenCL kernels executed by the CPU shim oracle/ref_shim.c) on one slab."""
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "codec_ref")
    if not os.path.exists(exe):
        return None
    synth = importlib.import_module(PKG + ".synth")
    with tempfile.TemporaryDirectory() as td:
        raw, dct, out = (os.path.join(td, n) for n in ("a.raw", "a.dct", "a.out"))
        synth.natural(W, H, 8, 1).tofile(raw)
        t0 = time.perf_counter()
        r1 = subprocess.run([exe, "encode", raw, dct, str(W), str(H), "8", "1"], capture_output=True, cwd=os.path.dirname(exe))
        t1 = time.perf_counter()
        r2 = subprocess.run([exe, "decode", dct, out, str(W), str(H), "8", "1"], capture_output=True, cwd=os.path.dirname(exe))
        t2 = time.perf_counter()
        if r1.returncode or r2.returncode or not os.path.exists(out):
            return None
    return {"value": 8 / (t2 - t0), "unit": UNIT, "kind": "reference", "cores": os.cpu_count() or 1,
            "encode_s_per_slab": t1 - t0, "decode_s_per_slab": t2 - t1,
            "sample": "8 frames (1 slab), the reference C codec CLI end to end (zlib level 9 included), its O(512^2)-per-cube "
                      "OpenCL kernels run by a pthreads CPU shim"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O
    O.build()
    synth = importlib.import_module(PKG + ".synth")
    cores = os.cpu_count() or 1
    W, H, F, cube, scaling, idx = config_of(args)
    # a bounded sample of the workload per step: SAMPLE slabs spread over the clip (slabs are independent, the whole clip
    # scales linearly); config c1 is small enough to run in full
    nslabs = F // cube
    slabs = list(range(nslabs)) if args.config == "c1" else sample_slabs(nslabs, args.ref_slabs)
    pieces = [slab_numpy(W, H, cube, s, 1) for s in slabs]   # the clip is seeded per slab: any slab can be made alone
    vals = []
    for i in range(args.warmup + args.steps):
        te = td = 0.0
        for p in pieces:
            a, b, _, _ = cpu_piece(p, cube, cores)
            te += a
            td += b
        if i >= args.warmup:
            vals.append((te, td))
    te = float(np.mean([v[0] for v in vals]))
    td = float(np.mean([v[1] for v in vals]))
    frames = len(slabs) * cube
    fps = frames / (te + td)
    line = {
        "impl": "reference", "metric": metric_name(W, H), "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * (te + td), "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),
        "note": "no JVM in the image: the Java Encoder/Decoder is timed as the oracle's C restatement of its "
                "algorithm (grouped-coefficient DCT on all cores, single-threaded quantise/Exp-Golomb)",
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{frames} frames ({len(slabs)} slabs: {slabs}) of the workload per step, encode+decode"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "encode_s_per_slab": te / len(slabs), "decode_s_per_slab": td / len(slabs),
        # the other flavour of the reference, for the record: slower than the Java algorithm, so the port above is the
        # conservative denominator
        "ref_c_host": ref_c_host_sample(W, H) if not args.no_ref_c else None,
    }
    print(json.dumps(line))
    return 0


def slab_numpy(W, H, cube, slab, seed):
    """One slab of the SURVEY.md 8d natural clip on the CPU (numpy RNG, seeded per slab)."""
    return importlib.import_module(PKG + ".synth").natural_slab(W, H, cube, slab, seed)


def l2_note(bytes_per_gpu):
    if bytes_per_gpu > 126e6:
        return "inputs (%.0f MB per GPU) larger than the 126 MB L2, no flush needed" % (bytes_per_gpu / 1e6)
    return "inputs (%.0f MB per GPU) FIT the 126 MB L2: a parity configuration, not a bench line" % (bytes_per_gpu / 1e6)


def config_of(args):
    W, H, F, cube, scaling, idx = CONFIGS[args.config]
    if args.width:
        W = args.width
    if args.height:
        H = args.height
    if args.frames:
        F = args.frames
    return W, H, F, cube, scaling, idx


def workload_config(args, world):
    W, H, F, cube, scaling, idx = config_of(args)
    total = F * world if scaling == "weak" else F
    per = f"{F} frames per GPU" if scaling == "weak" else f"{F} frames sharded over the GPUs"
    return {"workload": f"{W}x{H} gray, {per}, {cube}x{cube}x{cube} cubes (BASELINE configs[{idx}])", "name": args.config,
            "cube": cube, "total_frames": total,
            "l2": l2_note(W * H * (F if scaling == "weak" else F / world))}


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libdct3d has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    gloo = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        gloo = dist.new_group(backend="gloo")                   # host-side scalars of the e2e path

    importlib.import_module(PKG + ".build").build()
    codec = importlib.import_module(PKG + ".codec")
    sh = importlib.import_module(PKG + ".sharding")
    # the rank's host threads and, by first touch, its page-locked buffers stay on the NUMA node of its GPU
    numa = sh.bind_to_gpu_numa(local) if world > 1 and not os.environ.get("DCT3D_NO_NUMA_BIND") else None
    W, H, F, cube, scaling, idx = config_of(args)
    total_frames = F * world if scaling == "weak" else F
    nslabs = total_frames // cube
    lo, hi = sh.slab_range(nslabs, rank, world)
    Fr = (hi - lo) * cube                                       # this rank's frames
    N = W * H * Fr
    c = codec.Codec(W, H, cube, device=local)
    if args.tma is not None:
        c.set_option("tma", args.tma)
    c.set_option("reuse_zeroed", 1)     # the stream buffer is reused every step: wipe only what the last step wrote
    frames = synth_slabs_torch(W, H, cube, lo, hi, 1, dev)
    cap = N // 2 + 4096
    d_part = torch.zeros(cap, dtype=torch.uint8, device=dev)    # this rank's bits, coded from bit 0
    d_placed = torch.zeros(cap + 64, dtype=torch.uint8, device=dev) if world > 1 else None   # ... moved to their phase
    d_out = torch.empty_like(frames)
    # a real (non-NULL) stream: NULL would mean "the context's own stream" to libdct3d, invisible to torch events
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    st = tstream.cuda_stream
    assert st != 0
    xch = None
    if world > 1:
        # the one exchange step: N scalars through the host, here a shared-memory table of the node's ranks
        xname = "dct3d_xch_%s" % os.environ.get("MASTER_PORT", "0")
        if rank == 0:
            xch = sh.ShmExchange(xname, world, rank, create=True)
        dist.barrier()
        if rank != 0:
            xch = sh.ShmExchange(xname, world, rank, create=False)
        dist.barrier()

    def gather_counts(end):
        return xch.all_gather(end)

    def step(events=None, i=0):
        if events:
            events[3 * i].record()
        end = c.encode_u8_dev(frames, Fr, d_part, cap, 0, st)
        if world > 1:
            offs = sh.bit_offsets(gather_counts(end))
            phase = offs[rank] % 8
            c.stream_shift_dev(d_part, end, phase, d_placed, cap + 64, st)
            src, nbytes = d_placed, (phase + end) // 8 + 1
        else:
            offs, phase, src, nbytes = [0, end], 0, d_part, end // 8 + 1
        if events:
            events[3 * i + 1].record()
        c.decode_u8_dev(src, nbytes, Fr, d_out, phase, st)
        if events:
            events[3 * i + 2].record()
        return end, offs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled from the first warm-up step to the end of the timed region; warm-up runs for at
    # least 1.5 s of the same load so that nvidia-smi (100 ms period) sees the steady state
    sampler = ClockSampler(local) if rank == 0 else None
    t_w = time.perf_counter()
    nwarm = 0
    while True:
        nbits, offs = step()
        nwarm += 1
        go = torch.tensor([1 if (nwarm < max(args.warmup, 3) or time.perf_counter() - t_w < 1.5) else 0], device=dev)
        if world > 1:
            dist.all_reduce(go, op=dist.ReduceOp.MAX)            # the ranks must agree on the number of warm-up steps
        if not int(go.item()):
            break
    launches0 = c.stat("launches")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps)]
    k_enc, k_rec = [], []
    barrier()
    t_wall0 = time.perf_counter()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    c.set_option("kernel_times_reset", 1)
    for i in range(args.steps):
        step(ev, i)
    e_end.record()
    barrier()
    # device time of the two transform kernels over the timed region: libdct3d records CUDA events around every launch
    # on the launching stream (a ring of the last 32); read here, after the region, so that no step waits for them
    k_enc.append(c.stat("ns_encode_kernel_avg"))
    k_rec.append(c.stat("ns_reconstruct_kernel_avg"))
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    total_ms = e_start.elapsed_time(e_end)
    enc_ms = float(np.mean([ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(args.steps)]))
    dec_ms = float(np.mean([ev[3 * i + 1].elapsed_time(ev[3 * i + 2]) for i in range(args.steps)]))
    launches = c.stat("launches") - launches0
    kenc_ms, krec_ms = float(np.mean(k_enc)) * 1e-6, float(np.mean(k_rec)) * 1e-6
    device_roundtrip_ok = True

    t = torch.tensor([total_ms, enc_ms, dec_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, enc_ms_max, dec_ms_max = [float(v) for v in t.tolist()]
    ms_per_step = total_ms / args.steps
    value = total_frames / (ms_per_step * 1e-3)
    total_bits = offs[-1]
    S_total = total_bits // 8 + 1

    # the HBM-bound entry points (the reference's own float device boundary): GB/s of forward/inverse_f32
    seam = None
    if rank == 0 and not args.quick:
        ns = max(1, (64 * 1920 * 1080) // (W * H * cube))
        a = torch.empty(W * H * cube * ns, dtype=torch.float32, device=dev).uniform_(0, 255)
        b = torch.empty_like(a)
        res = {}
        for name, fn in (("forward_f32", c.forward_f32_dev), ("inverse_f32", c.inverse_f32_dev)):
            for _ in range(3):
                fn(a, b, ns, st)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn(a, b, ns, st)
            e1.record()
            torch.cuda.synchronize()
            res[name] = a.numel() * 8 / (e0.elapsed_time(e1) / 10 * 1e-3) / 1e9
        seam = res
        del a, b

    # ---- end to end through the host-buffer C ABI (pinned host memory; H2D/D2H inside the timed region) ----------
    # One call pair per step: the clip goes in as host frames and comes back as ONE host stream, which goes in again and
    # comes back as host frames.  The chunked H2D / kernels / D2H overlap lives inside the library calls.  With N ranks
    # every rank places its range into the same shared-memory stream and decodes its range from it.
    lib = c.L
    # Link-aware shares for the end-to-end path.  The GPUs of one box do not all get the same share of the host links (this
    # pool, 8 GPUs copying at once: 24 / 12 GB/s up / down for GPUs 0-3, 39 / 20 for GPUs 4-7), and a host-buffer call is
    # bound by its rank's link, so equal slab ranges leave the fast links idle while the slow ones finish.  Every rank
    # measures its own rates while all ranks copy at once; the clip's slabs are then shared out in proportion to
    # 1 / (1/up + 1/down).  The stream does not depend on the shares (same SHA-256); the device-resident measurement above
    # keeps equal ranges.  --balance equal switches this off.
    e_lo, e_hi, balance = lo, hi, None
    if world > 1 and args.balance == "auto":
        nprobe, reps = 32 << 20, 6
        hp = torch.empty(nprobe, dtype=torch.uint8, pin_memory=True)
        dp = torch.empty(nprobe, dtype=torch.uint8, device=dev)

        def rate(up):
            torch.cuda.synchronize()
            xch.all_gather(0)
            t0 = time.perf_counter()
            for _ in range(reps):
                (dp.copy_(hp, non_blocking=True) if up else hp.copy_(dp, non_blocking=True))
            torch.cuda.synchronize()
            return nprobe * reps / (time.perf_counter() - t0) / 1e9
        rate(True)
        up_gbs, down_gbs = rate(True), rate(False)
        ws = xch.all_gather(int(1e6 / (1.0 / up_gbs + 1.0 / down_gbs)))
        ups, downs = xch.all_gather(int(up_gbs * 1000)), xch.all_gather(int(down_gbs * 1000))
        balance = {"h2d_gbs": [u / 1000 for u in ups], "d2h_gbs": [d / 1000 for d in downs], "applied": False}
        if os.environ.get("DCT3D_BALANCE_WEIGHTS"):             # testing: force the shares
            ws = [int(1000 * float(x)) for x in os.environ["DCT3D_BALANCE_WEIGHTS"].split(",")][:world]
        if max(ws) > 1.15 * min(ws):                            # uneven links: follow them
            cum = [0.0]
            for w in ws:
                cum.append(cum[-1] + w)
            bounds = [0] + [int(nslabs * cum[g + 1] / cum[-1] + 0.5) for g in range(world - 1)] + [nslabs]
            e_lo, e_hi = bounds[rank], bounds[rank + 1]
            balance.update(applied=True, frames_per_rank=[(bounds[g + 1] - bounds[g]) * cube for g in range(world)])
        del hp, dp
    Fe = (e_hi - e_lo) * cube                                   # this rank's frames of the end-to-end path
    frames_e = frames if (e_lo, e_hi) == (lo, hi) else synth_slabs_torch(W, H, cube, e_lo, e_hi, 1, dev)
    h_frames = torch.empty((Fe, H, W), dtype=torch.uint8, pin_memory=True)
    h_frames.copy_(frames_e)
    h_out = torch.empty((Fe, H, W), dtype=torch.uint8, pin_memory=True)
    if frames_e is not frames:
        # what the device-resident path decodes for this rank's share, to compare the host-buffer result with
        cap_e = W * H * Fe // 2 + 4096
        d_tmp = torch.zeros(cap_e, dtype=torch.uint8, device=dev)
        d_out_e = torch.empty_like(frames_e)
        end_e = c.encode_u8_dev(frames_e, Fe, d_tmp, cap_e, 0, st)
        c.decode_u8_dev(d_tmp, end_e // 8 + 1, Fe, d_out_e, 0, st)
        torch.cuda.synchronize()
        del d_tmp
    else:
        d_out_e = d_out
    scap = (W * H * total_frames) // 2 + 4096
    shm = None
    if world > 1:
        name = "dct3d_bench_%s" % os.environ.get("MASTER_PORT", "0")
        if rank == 0:
            shm = sh.SharedStream(name, scap, create=True)
        dist.barrier()
        if rank != 0:
            shm = sh.SharedStream(name, scap, create=False)
        # every rank's GPU copies its placed bytes straight into the shared mapping, page-locked in every process
        # (measured on this pool: the same 0.7 ms per 38 MB as into a cudaHostAlloc buffer when a rank runs alone; a detour
        # through a private page-locked buffer and a host memcpy into the mapping was 2.4x slower with four ranks)
        assert lib.dct3d_host_register(shm.array.ctypes.data, scap) == 0
        h_stream_np = shm.array
        stream_ptr = shm.array.ctypes.data
    else:
        h_stream = torch.zeros(scap, dtype=torch.uint8, pin_memory=True)
        h_stream_np = h_stream.numpy()
        stream_ptr = h_stream.data_ptr()
    e2e_state = {}

    def e2e_encode():
        if world == 1:
            nb, ny = C.c_uint64(), C.c_size_t()
            rc = lib.dct3d_encode_u8(c.h, h_frames.data_ptr(), Fe, stream_ptr, scap, C.byref(nb), C.byref(ny))
            assert rc == 0, lib.dct3d_last_error(c.h)
            e2e_state["offs"] = [0, nb.value]
            return 0.0
        nb = C.c_uint64()
        ta = time.perf_counter()
        rc = lib.dct3d_encode_u8_range(c.h, h_frames.data_ptr(), Fe, C.byref(nb))
        assert rc == 0, lib.dct3d_last_error(c.h)
        t0 = time.perf_counter()                                # from here on: the concatenation
        o = sh.bit_offsets(xch.all_gather(nb.value))            # (waits for the slowest rank's range)
        t1 = time.perf_counter()
        fb = C.c_uint8(0)
        rc = lib.dct3d_encode_u8_place(c.h, o[rank], 1 if rank == world - 1 else 0, stream_ptr, scap, C.byref(fb))
        assert rc == 0, lib.dct3d_last_error(c.h)
        t2 = time.perf_counter()
        xch.signal()                                            # this rank's bytes have landed
        if o[rank] % 8:
            xch.wait_for(rank - 1)                              # ... and so have the predecessor's: OR the shared byte
            h_stream_np[o[rank] // 8] |= fb.value
        t3 = time.perf_counter()
        e2e_state["offs"] = o
        e2e_state["phases"] = [t0 - ta, t1 - t0, t2 - t1, t3 - t2]   # range, wait for all counts, place, boundary byte
        return t3 - t1

    def e2e_decode():
        o = e2e_state["offs"]
        end = C.c_uint64()
        rc = lib.dct3d_decode_u8_range(c.h, stream_ptr, o[-1] // 8 + 1, o[rank], o[rank + 1], Fe, h_out.data_ptr(), C.byref(end))
        assert rc == 0, lib.dct3d_last_error(c.h)
        assert end.value == o[rank + 1]

    def host_barrier():
        if world > 1:
            dist.barrier(group=gloo)

    e2e_encode()
    e2e_decode()
    e2e_steps = max(1, min(args.steps, 5))
    host_barrier()
    t0 = time.perf_counter()
    concat_s = te_s = 0.0
    phases = np.zeros(5)
    for _ in range(e2e_steps):
        ta = time.perf_counter()
        concat_s += e2e_encode()
        tb = time.perf_counter()
        te_s += tb - ta
        e2e_decode()
        phases += np.array(e2e_state.get("phases", [tb - ta, 0, 0, 0]) + [time.perf_counter() - tb])
        if xch is not None:
            # the ranks start every step together: on this pool's hosts the links carry 110 GB/s when every GPU copies the same
            # way but 51 GB/s each way when some upload while others download, so a rank that ran ahead into its next encode
            # would slow everybody's decode (measured: 17.8 k -> 21 k frames/s at 2 GPUs)
            xch.all_gather(0)
    host_barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te2 = torch.tensor([e2e_s, te_s / e2e_steps, concat_s / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te2, op=dist.ReduceOp.MAX)
    e2e_s, e2e_enc_s, concat_s = [float(v) for v in te2.tolist()]
    # per-rank phase times (ms): range coding, waiting for every rank's bit count, placement, boundary byte, decode
    ph = torch.tensor(phases / e2e_steps * 1e3, dtype=torch.float64, device=dev)
    ph_all = [torch.zeros_like(ph) for _ in range(world)]
    if world > 1:
        dist.all_gather(ph_all, ph)
    else:
        ph_all = [ph]
    phase_table = [[round(float(v), 2) for v in t.tolist()] for t in ph_all]
    e2e_value = total_frames / e2e_s
    chunks_per_call = c.stat("chunks")
    roundtrip_ok = bool((h_out.to(dev) == d_out_e).all().item())
    ok_t = torch.tensor([1 if roundtrip_ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
    roundtrip_ok = bool(int(ok_t.item()))
    sha = hashlib.sha256(h_stream_np[:S_total].tobytes()).hexdigest() if rank == 0 else None
    e2e_h2d = W * H * total_frames + S_total
    # two clips in flight at N = 1: a second context decodes clip k while the first encodes clip k+1 (PCIe is full
    # duplex; every call is still the product call and every clip one stream)
    duplex_value = duplex_trace = None
    if world == 1 and not args.quick:
        c2 = codec.Codec(W, H, cube, device=local)
        h_stream2 = [h_stream, torch.zeros(scap, dtype=torch.uint8, pin_memory=True)]
        sizes = [0, 0]
        nsteps = 2 * e2e_steps + 2
        trace = {"enc": [], "dec": []}
        ready = [threading.Semaphore(0), threading.Semaphore(0)]
        free = [threading.Semaphore(1), threading.Semaphore(1)]
        errs = []

        def enc_thread():
            nb, ny = C.c_uint64(), C.c_size_t()
            for k in range(nsteps):
                free[k & 1].acquire()
                ta_ = time.perf_counter()
                if lib.dct3d_encode_u8(c.h, h_frames.data_ptr(), Fr, h_stream2[k & 1].data_ptr(), scap, C.byref(nb), C.byref(ny)) != 0:
                    errs.append(lib.dct3d_last_error(c.h))
                trace["enc"].append((ta_, time.perf_counter()))
                sizes[k & 1] = ny.value
                ready[k & 1].release()

        def dec_thread():
            for k in range(nsteps):
                ready[k & 1].acquire()
                ta_ = time.perf_counter()
                if lib.dct3d_decode_u8(c2.h, h_stream2[k & 1].data_ptr(), sizes[k & 1], Fr, h_out.data_ptr()) != 0:
                    errs.append(lib.dct3d_last_error(c2.h))
                trace["dec"].append((ta_, time.perf_counter()))
                free[k & 1].release()

        assert lib.dct3d_decode_u8(c2.h, stream_ptr, e2e_state["offs"][-1] // 8 + 1, Fr, h_out.data_ptr()) == 0   # warm-up: allocations
        ths = [threading.Thread(target=enc_thread), threading.Thread(target=dec_thread)]
        t0 = time.perf_counter()
        for t_ in ths:
            t_.start()
        for t_ in ths:
            t_.join()
        duplex_value = nsteps * Fr / (time.perf_counter() - t0)
        assert not errs, errs
        duplex_trace = {k: [(round((a - t0) * 1e3, 1), round((b - t0) * 1e3, 1)) for a, b in v] for k, v in trace.items()}
        c2.close()

    # ---- extra device-resident lines at N = 1: worst-case content and the 4^3 variant ------------------------------
    extra = {}
    if world == 1 and rank == 0 and not args.quick and args.config == "c2":
        def quick_fps(cc, fr, nfr, reps=5):
            cp = fr.numel() // 2 + 4096              # 4 bit/sample: noise codes at 3.3
            ds = torch.zeros(cp, dtype=torch.uint8, device=dev)
            do = torch.empty_like(fr)
            for _ in range(3):
                e_ = cc.encode_u8_dev(fr, nfr, ds, cp, 0, st)
                cc.decode_u8_dev(ds, e_ // 8 + 1, nfr, do, 0, st)
            evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(reps)]
            for a_, b_, c_ in evs:
                a_.record()
                e_ = cc.encode_u8_dev(fr, nfr, ds, cp, 0, st)
                b_.record()
                cc.decode_u8_dev(ds, e_ // 8 + 1, nfr, do, 0, st)
                c_.record()
            torch.cuda.synchronize()
            te_ = sum(a_.elapsed_time(b_) for a_, b_, c_ in evs)
            td_ = sum(b_.elapsed_time(c_) for a_, b_, c_ in evs)
            return {"frames": nfr, "encode_fps": nfr * reps / (te_ * 1e-3), "decode_fps": nfr * reps / (td_ * 1e-3),
                    "bits_per_sample": e_ / fr.numel()}
        noise = synth_slabs_torch(W, H, cube, 0, 8, 2, dev, kind="noise")
        extra["noise_content"] = quick_fps(c, noise, 64)
        del noise
        with codec.Codec(W, H, 4, device=local) as c4:
            f4 = synth_slabs_torch(W, H, 4, 0, 64, 1, dev)
            extra["cube4"] = quick_fps(c4, f4, 256)
            del f4
        c.set_option("precision", 64)
        f64 = frames[:64].contiguous()
        extra["fp64_mode"] = quick_fps(c, f64, 64, reps=3)
        c.set_option("precision", 32)

    # ---- the C codec command line (host/codec: file -> batches -> GPU -> parallel deflate -> file, and back) ------------
    cli = None
    if world == 1 and not args.quick and not args.no_cli and os.path.exists(os.path.join(ROOT, "host", "codec")):
        import tempfile
        exe = os.path.join(ROOT, "host", "codec")
        d = tempfile.mkdtemp(prefix="dct3d_cli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            raw, enc, dec = (os.path.join(d, n) for n in ("clip.raw", "clip.dct", "clip.out"))
            h_frames.numpy().tofile(raw)
            cli = {"clip": f"{W}x{H}x{Fr} in tmpfs", "cores": os.cpu_count()}
            def loop_s(err):                                   # "dct3d-cli ...: setup a s, loop b s, total c s" (DCT3D_CLI_TIMING)
                for line in err.splitlines():
                    if line.startswith("dct3d-cli"):
                        return float(line.split("loop")[1].split("s")[0])
                return None
            for level in (9, 1):
                env = dict(os.environ, DCT3D_ZLIB_LEVEL=str(level), DCT3D_CLI_TIMING="1")
                t0 = time.perf_counter()
                r1 = subprocess.run([exe, "encode", raw, enc, str(W), str(H), str(Fr), str(local + 1)], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
                t1 = time.perf_counter()
                r2 = subprocess.run([exe, "decode", enc, dec, str(W), str(H), str(Fr), str(local + 1)], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
                t2 = time.perf_counter()
                ok = r1.returncode == 0 and r2.returncode == 0 and bool((np.fromfile(dec, np.uint8) == h_out.numpy().reshape(-1)).all())
                le, ld = loop_s(r1.stderr), loop_s(r2.stderr)
                cli[f"zlib{level}"] = {"encode_fps": Fr / le if le else None, "decode_fps": Fr / ld if ld else None,
                                      "encode_process_s": t1 - t0, "decode_process_s": t2 - t1, "file_bytes": os.path.getsize(enc),
                                      "decoded_equals_library_decode": ok}
            cli["note"] = ("*_fps: the codec's batch loop (file read -> GPU -> deflate -> file write and back), after context creation and "
                           "buffer allocation; *_process_s: wall clock of the whole process; deflate runs on all cores "
                           "(host/pdeflate.c), inflate is one serial zlib stream on its own thread")
        finally:
            import shutil
            shutil.rmtree(d, ignore_errors=True)

    # ---- CPU baseline beside it + the four parity rules on the same sample slabs (rank 0, N = 1) -----------------
    cpu = parity = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        from oracle import parity as PAR
        O.build()
        cores = os.cpu_count() or 1
        slabs = list(range(nslabs)) if nslabs <= 8 else sample_slabs(nslabs, 4)
        te_ = td_ = 0.0
        parts = []
        for s in slabs:
            piece = frames[s * cube:(s + 1) * cube].cpu().numpy()
            a_, b_, ostream, odec = cpu_piece(piece, cube, cores)
            te_ += a_
            td_ += b_
            p = PAR.piece_parity(c, piece, cube)
            # the benchmark's own device-resident decode of this slab (from the whole clip's stream) is the decode of the
            # slab's own stream, which piece_parity has just compared with the oracle's decode of the same bits
            own, _ = c.encode_u8(piece)
            p["bench_decode_equal"] = bool((d_out[s * cube:(s + 1) * cube].cpu().numpy() == c.decode_u8(own, cube)).all())
            parts.append(p)
        cpu = {"value": len(slabs) * cube / (te_ + td_), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{len(slabs) * cube} frames (slabs {slabs} of the workload), encode+decode, oracle Java-structured port",
               "encode_s_per_slab": te_ / len(slabs), "decode_s_per_slab": td_ / len(slabs)}
        parity = PAR.merge(parts)
        parity["slabs_checked"] = slabs
        parity["bench_decode_equal"] = all(p["bench_decode_equal"] for p in parts)
        parity["rule"] = "coef <= 1e-4 rel; cubes equal except counted +-1 tie flips; stream bit-exact given cubes; pixels +-1"

    if rank == 0:
        peak, peak_src = peaks()
        S_rank = nbits // 8 + 1
        alg_bytes = N + S_rank   # algorithmic bytes of encode_u8 and of decode_u8 per launch: pixels + stream
        # dominant kernel of the step = the longer of the two transform kernels, timed live by CUDA events
        # recorded around it on the launching stream inside libdct3d
        dom = ((f"reconstruct_coo_kernel<{cube}>", krec_ms) if krec_ms >= kenc_ms else (f"encode_kernel<{cube},MODE_ZZ>", kenc_ms))
        achieved = alg_bytes / (dom[1] * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(dom[0].split("<")[0]) if args.config == "c2" else (None, None)
        line = {
            "metric": metric_name(W, H), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(args, world),
            "detail": {"frames_per_gpu": Fr, "warmup_steps_run": nwarm, "numa_bind": numa, "tma": c.stat("tma"), "stream_bytes": int(S_total),
                       "bits_per_sample": total_bits / (W * H * total_frames),
                       "parallelism": f"slab-range x{world}, one clip, one stream; N bit counts through a host table, no data-path collective"},
            "encode_fps": total_frames / (enc_ms_max * 1e-3), "decode_fps": total_frames / (dec_ms_max * 1e-3),
            "encode_ms": enc_ms_max, "decode_ms": dec_ms_max,
            "roofline": {"bound": "hbm", "kernel": dom[0], "kernel_ms": dom[1], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes": int(alg_bytes),
                         "traffic_source": traffic_src,
                         "step_frac": 2 * alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                         "note": "the fused u8<->bitstream kernels are issue-bound, not HBM-bound (DESIGN.md 4); "
                                 "the HBM-bound float seam is in roofline_f32_seam"},
            "kernels_ms": {"encode_kernel": kenc_ms, "reconstruct_coo_kernel": krec_ms},
            "roofline_f32_seam": None if seam is None else {
                "bound": "hbm", "unit": "GB/s", "peak": peak, "algorithmic_bytes_per_sample": 8,
                "forward_f32": seam["forward_f32"], "inverse_f32": seam["inverse_f32"],
                "frac": min(seam["forward_f32"], seam["inverse_f32"]) / peak},
            "cpu_baseline": cpu,
            "parity": parity,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(e2e_h2d), "d2h_bytes_per_step": int(e2e_h2d),
                    "steps": e2e_steps, "matches_device_path": roundtrip_ok, "encode_ms": e2e_enc_s * 1e3,
                    "decode_ms": (e2e_s - e2e_enc_s) * 1e3, "concat_ms": concat_s * 1e3,
                    "per_rank_ms[range,wait_counts,place,boundary,decode]": phase_table,
                    "stream_sha256": sha, "balance": balance, "duplex_value": duplex_value, "duplex_trace_ms": duplex_trace, "chunks_per_call": chunks_per_call,
                    "how": ("one dct3d_encode_u8 + one dct3d_decode_u8 per step on pinned host buffers" if world == 1 else
                            f"per rank: dct3d_encode_u8_range, {world} bit counts exchanged through a shared-memory table, dct3d_encode_u8_place "
                            "straight into the page-locked shared-memory stream at byte B_g/8, boundary byte OR-ed once the predecessor has landed (concat_ms = all of that, max over ranks), then "
                            "dct3d_decode_u8_range from the rank's global start bit") +
                           "; the chunked H2D / kernel / D2H overlap is inside the calls; duplex_value = a second context decodes "
                           "clip k while clip k+1 is encoded",
                    "bound": "PCIe: %.2f GB each way per step" % (e2e_h2d / 1e9)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "wall_s": t_wall,
            "cli": cli,
            "extra": extra,
        }
        print(json.dumps(line))
    if shm is not None:
        lib.dct3d_host_unregister(shm.array.ctypes.data)
        dist.barrier()
        if rank == 0:
            shm.unlink()
            xch.unlink()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", choices=sorted(CONFIGS), default="c2")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--tma", type=int, default=None)
    ap.add_argument("--ref-slabs", type=int, default=4, help="reference arm: sample slabs of the workload per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-c", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip timing the C codec command line")
    ap.add_argument("--balance", choices=["auto", "equal"], default="auto",
                    help="N > 1, end-to-end path: share the slabs out by measured host-link rates (auto) or equally")
    ap.add_argument("--quick", action="store_true", help="skip the seam, duplex and extra measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
