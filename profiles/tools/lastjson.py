"""Prints selected keys of the last JSON line on stdin (bench.py output may be preceded by library banners)."""
import json, sys
lines = [l for l in sys.stdin.read().splitlines() if l.startswith("{")]
d = json.loads(lines[-1])
keys = sys.argv[1:] or ["value", "ms_per_step", "encode_ms", "decode_ms", "e2e"]
def get(d, k):
    for part in k.split("."):
        d = d[part]
    return d
print({k: get(d, k) for k in keys})
