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


def per_launch(path, pattern):
    """one line per launch whose kernel name contains `pattern`: time and DRAM bytes when the capture has them"""
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ii, ki, mi, vi, ui = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    by = defaultdict(dict)
    names = {}
    for r in rows[1:]:
        if pattern not in r[ki]:
            continue
        names[r[ii]] = r[ki].split("(")[0].replace("void ", "")[-60:]
        by[r[ii]][r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    for i in sorted(by, key=int):
        d = by[i]
        t = d.get("gpu__time_duration.sum", 0.0)
        rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
        gbs = (rd + wr) / (t * 1e-3) / 1e9 if t else 0.0
        print(f"{int(i):5d} {t:10.3f} ms  read {rd / 1e9:8.3f} GB  write {wr / 1e9:8.3f} GB  {gbs:8.0f} GB/s  {names[i]}")


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[2] == "--per-launch":
        per_launch(sys.argv[1], sys.argv[3])
    else:
        main(sys.argv[1], " ".join(sys.argv[2:]))
