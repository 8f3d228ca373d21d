#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import csv
import sys
from collections import defaultdict


def main(path, title=""):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = defaultdict(float); cnt = defaultdict(int)
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki]
        for cut in ("(",):
            name = name.split(cut)[0]
        name = name.replace("void ", "").strip()
        tot[name] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-6)
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# {title}")
    print("# per-launch times are cold-cache and serialised; compare SHARES")
    print(f"# total kernel time {total:.1f} ms over {sum(cnt.values())} launches")
    for name, t in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{t:12.3f} ms  {100 * t / total:6.2f}%  x{cnt[name]:<4d} {name[-90:]}")


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
