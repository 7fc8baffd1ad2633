#!/usr/bin/env python3
"""Warp-state (stall reason) distribution per kernel from an ncu report with source counters:

    python profiles/tools/stalls.py gpurun_out/r1_full.ncu-rep > profiles/r1_stalls.md
"""
import collections
import csv
import io
import subprocess
import sys

KERNELS = ["encode_kernel", "eg_pack_(sorted_)?kernel", "seg_scan_kernel", "seg_emit_kernel", "reconstruct_coo_kernel"]


def page(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    print("# Warp states per kernel (ncu source counters, sampled; one launch each at 1920x1080x256)\n")
    print("`selected` = the warp issued; everything else is why a resident warp did not. Shares of all warp samples.\n")
    table = {}
    reasons = collections.Counter()
    for k in KERNELS:
        hdr, body = page(rep, k)
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        n_i = hdr.index("# Samples")
        seen = set()
        tot = collections.Counter()
        for r in body:
            if len(r) <= n_i or not r[n_i].strip().isdigit():
                continue
            key = (r[hdr.index("Address")] if "Address" in hdr else None, r[hdr.index("Source")])
            if key in seen:                      # some exports list every instruction twice
                continue
            seen.add(key)
            for s in stalls:
                v = r[hdr.index(s)]
                if v.strip().isdigit():
                    tot[s] += int(v)
        total = sum(tot.values()) or 1
        table[k] = {s: 100.0 * v / total for s, v in tot.items()}
        for s, v in tot.items():
            reasons[s] += v
    cols = [s for s, _ in reasons.most_common(10)]
    print("| kernel | " + " | ".join(c.replace("stall_", "") for c in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for k in KERNELS:
        print(f"| {k} | " + " | ".join(f"{table[k].get(c, 0):.1f}%" for c in cols) + " |")


if __name__ == "__main__":
    main()
