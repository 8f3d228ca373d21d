#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export into one line per launch with the metrics the roofline needs."""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("smsp__cycles_active.avg", "cyc"),
]


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1]
    name_i = hdr.index("Kernel Name")
    cols = {h: i for i, h in enumerate(hdr)}
    print("# " + path)
    for r in rows[2:]:
        out = [r[name_i][:34].ljust(34)]
        for key, short in KEYS:
            match = [h for h in cols if h == key or h.startswith(key)]
            if not match:
                continue
            i = cols[match[0]]
            v = num(r[i])
            if v is None:
                continue
            u = units[i]
            if short in ("dram_rd", "dram_wr", "l2_bytes"):
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                out.append(f"{short}={v * scale / 1e6:.1f}MB")
            elif short == "time":
                scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "second": 1e3, "nsecond": 1e-6}.get(u, 1)
                out.append(f"time={v * scale:.3f}ms")
            else:
                out.append(f"{short}={v:.1f}")
        print("  ".join(out))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
