#!/usr/bin/env python
"""bench.py -- 1080p grayscale encode+decode throughput of the 3D-DCT codec hot path.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (libdct3d.so)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

One step = one pass of the hot path (u8 frames -> Exp-Golomb stream -> u8 frames) over
BASELINE.json configs[1]: 1920x1080 grayscale, 256 frames, 8x8x8 cubes, synthetic "natural"
clip (SURVEY.md 8d generator).  `value` is whole-job frames/s with the frames resident in HBM
(CUDA events on the launching stream, max over ranks); `e2e` is the same metric through the
host-buffer C ABI (dct3d_encode_u8 / dct3d_decode_u8) with pinned host memory, H2D/D2H inside the
timed region.  Multi-GPU: one process per GPU (torchrun), each rank codes its own 256-frame slab
range (slabs are independent key-frame groups), no data-path collective -> weak scaling.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
METRIC = "1080p gray encode+decode frames/s"
UNIT = "frames/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def synth_clip_torch(W, H, F, seed, device):
    """The SURVEY.md 8d 'natural' generator, evaluated on the GPU (same formula, torch RNG)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    t = torch.arange(F, device=device, dtype=torch.float32)[:, None, None]
    y = torch.arange(H, device=device, dtype=torch.float32)[None, :, None]
    x = torch.arange(W, device=device, dtype=torch.float32)[None, None, :]
    out = torch.empty((F, H, W), dtype=torch.uint8, device=device)
    step = 32
    for f0 in range(0, F, step):
        tt = t[f0:f0 + step]
        v = 128.0 + 60.0 * torch.sin((x + 3.0 * tt) / 37.0) + 50.0 * torch.cos((y - 2.0 * tt) / 23.0)
        v = v + 6.0 * torch.randn(v.shape, generator=g, device=device)
        out[f0:f0 + step] = v.round().clamp(0, 255).to(torch.uint8)
    return out


def cpu_baseline_sample(W, H, frames, threads):
    """Times the oracle's Java-structured port (oracle/dct3d_oracle.c) on `frames` frames."""
    from oracle import oracle as O
    synth = importlib.import_module(PKG + ".synth")
    clip = synth.natural(W, H, frames, 1)
    t0 = time.perf_counter()
    stream, bits = O.java_encode_u8(clip, 8, threads)
    t1 = time.perf_counter()
    O.java_decode_u8(stream, W, H, frames, 8, threads)
    t2 = time.perf_counter()
    return frames / (t2 - t0), t1 - t0, t2 - t1, bits


def ref_c_host_sample(W, H):
    """Times the reference's own C host code (encoder.c / decoder.c / ExpGolomb.c / CubeUtils.c compiled unmodified
    into oracle/_ref/codec_ref, its four OpenCL kernels executed by the CPU shim oracle/ref_shim.c) on one slab."""
    import subprocess
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
    cores = os.cpu_count() or 1
    W, H = args.width, args.height
    sample = 8
    vals = []
    for i in range(args.warmup + args.steps):
        fps, te, td, _ = cpu_baseline_sample(W, H, sample, cores)
        if i >= args.warmup:
            vals.append((fps, te, td))
    fps = float(np.mean([v[0] for v in vals]))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sample / fps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{W}x{H} gray, 256 frames, 8x8x8 cubes (BASELINE configs[1])", "cube": 8,
                   "note": "no JVM in the image: the Java Encoder/Decoder is timed as the oracle's C restatement of its "
                           "algorithm (grouped-coefficient DCT on all cores, single-threaded quantise/Exp-Golomb)"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} frames (1 slab) of the workload per step, encode+decode"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "encode_s_per_slab": float(np.mean([v[1] for v in vals])), "decode_s_per_slab": float(np.mean([v[2] for v in vals])),
        # the other flavour of the reference, for the record: slower than the Java algorithm, so the port above is the
        # conservative denominator
        "ref_c_host": ref_c_host_sample(W, H),
    }
    print(json.dumps(line))
    return 0


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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    importlib.import_module(PKG + ".build").build()
    codec = importlib.import_module(PKG + ".codec")
    W, H, F, cube = args.width, args.height, args.frames, 8
    N = W * H * F
    c = codec.Codec(W, H, cube, device=local)
    if args.tma is not None:
        c.set_option("tma", args.tma)
    c.set_option("reuse_zeroed", 1)     # the stream buffer is reused every step: wipe only what the last step wrote
    frames = synth_clip_torch(W, H, F, 1 + rank, dev)
    cap = N // 2 + 4096
    d_stream = torch.zeros(cap, dtype=torch.uint8, device=dev)
    d_out = torch.empty_like(frames)
    # a real (non-NULL) stream: NULL would mean "the context's own stream" to libdct3d, invisible to torch events
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    st = tstream.cuda_stream
    assert st != 0

    def step():
        end = c.encode_u8_dev(frames, F, d_stream, cap, 0, st)
        nbytes = end // 8 + 1
        c.decode_u8_dev(d_stream, nbytes, F, d_out, 0, st)
        return end

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled from the first warm-up step to the end of the timed region; warm-up runs for at
    # least 1.5 s of the same load so that nvidia-smi (100 ms period) sees the steady state
    sampler = ClockSampler(local) if rank == 0 else None
    t_w = time.perf_counter()
    nwarm = 0
    while nwarm < max(args.warmup, 3) or time.perf_counter() - t_w < 1.5:
        nbits = step()
        nwarm += 1
    S = nbits // 8 + 1
    launches0 = c.stat("launches")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps)]
    k_enc, k_rec = [], []
    barrier()
    t_wall0 = time.perf_counter()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    for i in range(args.steps):
        ev[3 * i].record()
        end = c.encode_u8_dev(frames, F, d_stream, cap, 0, st)
        ev[3 * i + 1].record()
        c.decode_u8_dev(d_stream, end // 8 + 1, F, d_out, 0, st)
        ev[3 * i + 2].record()
        k_enc.append(c.stat("ns_encode_kernel"))     # the step has already synchronised (decode returns its end bit)
        k_rec.append(c.stat("ns_reconstruct_kernel"))
    e_end.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    total_ms = e_start.elapsed_time(e_end)
    enc_ms = float(np.mean([ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(args.steps)]))
    dec_ms = float(np.mean([ev[3 * i + 1].elapsed_time(ev[3 * i + 2]) for i in range(args.steps)]))
    launches = c.stat("launches") - launches0
    kenc_ms, krec_ms = float(np.mean(k_enc)) * 1e-6, float(np.mean(k_rec)) * 1e-6

    # the HBM-bound entry points (the reference's own float device boundary): GB/s of forward/inverse_f32
    seam = None
    if rank == 0:
        ns = 8
        a = torch.empty(W * H * 8 * ns, dtype=torch.float32, device=dev).uniform_(0, 255)
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

    t = torch.tensor([total_ms, enc_ms, dec_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, enc_ms_max, dec_ms_max = [float(v) for v in t.tolist()]
    ms_per_step = total_ms / args.steps
    value = world * F / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer C ABI (pinned host memory) ---------------------------
    # The clip is streamed the way the reference's C codec streams it (slab ranges in a loop): an
    # encoder context on one host thread and a decoder context on another, so that the H2D copy of
    # range i overlaps the D2H copy of range i-1 (PCIe is full duplex and is what bounds this number).
    # Every range is a self-contained stream coded from bit 0, exactly like a multi-GPU slab range.
    import queue
    import threading as th
    h_frames = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True)
    h_frames.copy_(frames)
    h_out = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True)
    nchunk = int(os.environ.get("DCT3D_E2E_RANGES", 16)) if F % 128 == 0 else 1
    nthr = int(os.environ.get("DCT3D_E2E_THREADS", 2)) if nchunk > 1 else 1   # contexts (= host threads) per direction
    cf = F // nchunk                                   # frames per range
    ccap = W * H * cf // 2 + 4096
    h_streams = [torch.zeros(ccap, dtype=torch.uint8, pin_memory=True) for _ in range(nchunk)]
    enc_ctxs = [c] + [codec.Codec(W, H, cube, device=local) for _ in range(nthr - 1)]
    dec_ctxs = [codec.Codec(W, H, cube, device=local) for _ in range(nthr)]
    L = c.L
    fsz = W * H * cf
    sizes = [0] * nchunk

    def e2e_step():
        q = queue.Queue()
        err = []

        def enc(k):
            nb, ny = C.c_uint64(), C.c_size_t()
            h = enc_ctxs[k].h
            for i in range(k, nchunk, nthr):
                rc = L.dct3d_encode_u8(h, h_frames.data_ptr() + i * fsz, cf, h_streams[i].data_ptr(), ccap, C.byref(nb), C.byref(ny))
                if rc != 0:
                    err.append(L.dct3d_last_error(h))
                sizes[i] = ny.value
                q.put(i)

        def dec(k):
            h = dec_ctxs[k].h
            while True:
                i = q.get()
                if i < 0:
                    return
                rc = L.dct3d_decode_u8(h, h_streams[i].data_ptr(), sizes[i], cf, h_out.data_ptr() + i * fsz)
                if rc != 0:
                    err.append(L.dct3d_last_error(h))

        te_ = [th.Thread(target=enc, args=(k,)) for k in range(nthr)]
        td_ = [th.Thread(target=dec, args=(k,)) for k in range(nthr)]
        for t_ in te_ + td_:
            t_.start()
        for t_ in te_:
            t_.join()
        for _ in td_:
            q.put(-1)
        for t_ in td_:
            t_.join()
        assert not err, err

    def e2e_single():
        nb, ny = C.c_uint64(), C.c_size_t()
        big = h_streams[0] if nchunk == 1 else torch.zeros(cap, dtype=torch.uint8, pin_memory=True)
        t0 = time.perf_counter()
        rc = L.dct3d_encode_u8(c.h, h_frames.data_ptr(), F, big.data_ptr(), big.numel(), C.byref(nb), C.byref(ny))
        assert rc == 0, L.dct3d_last_error(c.h)
        rc = L.dct3d_decode_u8(c.h, big.data_ptr(), ny.value, F, h_out.data_ptr())
        assert rc == 0, L.dct3d_last_error(c.h)
        return time.perf_counter() - t0

    e2e_step()
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * F / float(te.item())
    roundtrip_ok = bool((h_out.to(dev) == d_out).all().item())
    e2e_bytes = N + sum(sizes)
    e2e_single()
    single_s = e2e_single()
    for x in enc_ctxs[1:] + dec_ctxs:
        x.close()

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = N + S      # algorithmic bytes of encode_u8 and of decode_u8: pixels + stream
        # dominant kernel of the step = the longer of the two transform kernels, timed live by CUDA events
        # recorded around it on the launching stream inside libdct3d
        dom = ("reconstruct_coo_kernel<8>", krec_ms) if krec_ms >= kenc_ms else ("encode_kernel<8,MODE_ZZ>", kenc_ms)
        achieved = alg_bytes / (dom[1] * 1e-3) / 1e9
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            fps, te_, td_, _ = cpu_baseline_sample(W, H, 8, cores)
            cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "8 frames (1 slab) of the workload, encode+decode, oracle Java-structured port"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{W}x{H} gray, {F} frames per GPU, 8x8x8 cubes (BASELINE configs[1])", "cube": 8,
                       "frames_per_gpu": F, "l2": "inputs (%.0f MB) larger than L2" % (N / 1e6),
                       "tma": c.stat("tma"), "stream_bytes": int(S), "bits_per_sample": nbits / N,
                       "parallelism": f"slab-range x{world}, no collective"},
            "encode_fps": world * F / (enc_ms_max * 1e-3), "decode_fps": world * F / (dec_ms_max * 1e-3),
            "encode_ms": enc_ms_max, "decode_ms": dec_ms_max,
            "roofline": {"bound": "hbm", "kernel": dom[0], "kernel_ms": dom[1], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": 543.0e6 if dom[0].startswith("recon") else 635.0e6,
                         "peak_source": peak_src, "algorithmic_bytes": int(alg_bytes),
                         "traffic_source": "ncu dram__bytes_read+write of that kernel, profiles/r1_summary.md",
                         "note": "the fused u8<->bitstream kernels are issue-bound (SM 62-82%, DRAM 13-21%), not HBM-bound: "
                                 "DESIGN.md 4; the HBM-bound float seam is in roofline_f32_seam"},
            "kernels_ms": {"encode_kernel": kenc_ms, "reconstruct_coo_kernel": krec_ms},
            "roofline_f32_seam": None if seam is None else {
                "bound": "hbm", "unit": "GB/s", "peak": peak, "algorithmic_bytes_per_sample": 8,
                "forward_f32": seam["forward_f32"], "inverse_f32": seam["inverse_f32"],
                "frac": min(seam["forward_f32"], seam["inverse_f32"]) / peak},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(e2e_bytes), "d2h_bytes_per_step": int(e2e_bytes),
                    "steps": e2e_steps, "matches_device_path": roundtrip_ok,
                    "how": f"dct3d_encode_u8 / dct3d_decode_u8 on pinned host buffers, {nchunk} slab ranges of {cf} frames streamed "
                           f"through {nthr} encoder and {nthr} decoder contexts, one host thread each (PCIe is full duplex: the H2D of one "
                           "range overlaps the compute and the D2H of others)",
                    "single_call_value": world * F / single_s,
                    "bound": "PCIe: %.2f GB each way per step" % (e2e_bytes / 1e9)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "wall_s": t_wall,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--tma", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
