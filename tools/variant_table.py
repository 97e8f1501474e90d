"""Prints one line per variant of a variant_bench.py log: the kernel times and the image checksums."""
import json
import sys

for ln in open(sys.argv[1]):
    try:
        d = json.loads(ln)
    except Exception:
        print(ln[:300].rstrip())
        continue
    print(d["tag"], {k: round(v, 1) for k, v in d.items() if k.endswith("_ms")}, {k: v for k, v in d.items() if k.endswith("_checksum") and not k.startswith("c3")})
