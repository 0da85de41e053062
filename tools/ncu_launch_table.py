#!/usr/bin/env python
"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list
per kernel: launches, total device time and share, DRAM bytes per launch."""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    agg = defaultdict(lambda: {"n": set(), "t": 0.0, "rd": 0.0, "wr": 0.0})
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("void ", "").replace("swc::", "").replace("<unnamed>::", "")
        a = agg[name]
        a["n"].add(r[ci["ID"]])
        v = float(r[ci["Metric Value"]].replace(",", ""))
        u = r[ci["Metric Unit"]]
        m = r[ci["Metric Name"]]
        if m == "gpu__time_duration.sum":
            a["t"] += v * scale.get(u, 1.0)
        elif m == "dram__bytes_read.sum":
            a["rd"] += v * scale.get(u, 1.0)
        elif m == "dram__bytes_write.sum":
            a["wr"] += v * scale.get(u, 1.0)
    tot = sum(a["t"] for a in agg.values()) or 1.0
    print("| kernel | launches | total us | share | DRAM read MB/launch | DRAM write MB/launch |")
    print("|---|---|---|---|---|---|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        n = len(a["n"])
        print(f"| `{name[:90]}` | {n} | {a['t']:.0f} | {a['t'] / tot:.3f} | {a['rd'] / n / 1e6:.1f} | {a['wr'] / n / 1e6:.1f} |")
    g = [a for k, a in agg.items() if "gemm_tc" in k]
    if g:
        n = sum(len(a["n"]) for a in g)
        print(f"\ntcgen05 GEMM launches: {n}, DRAM bytes per launch (read+write): {sum(a['rd'] + a['wr'] for a in g) / n / 1e6:.1f} MB, "
              f"total {sum(a['rd'] + a['wr'] for a in g) / 1e9:.2f} GB")


if __name__ == "__main__":
    main()
