#!/usr/bin/env python
"""Print the SASS instructions with the most warp-stall samples per kernel from `ncu --page source --csv`."""
import csv
import subprocess
import sys


def main():
    rep, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 20
    only = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    kernels, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
        elif r and r[0] == "Address":
            hdr = {h: i for i, h in enumerate(r)}
        elif cur is not None and hdr and len(r) > 5:
            cur["rows"].append(r)
    for ki, k in enumerate(kernels):
        if only is not None and str(ki) != only:
            continue
        data = k["rows"]
        s = [int(r[hdr["# Samples"]]) for r in data]
        tot = sum(s) or 1
        print(f"== kernel {ki}: {k['name'][:110]}  samples={tot}")
        top = sorted(range(len(data)), key=lambda i: -s[i])[:topn]
        for i in sorted(top):
            r = data[i]
            print(f"  [{i:5d}] {100.0 * s[i] / tot:5.1f}%  exec={r[hdr['Instructions Executed']]:>9}  {r[hdr['Source']].strip()[:110]}")


if __name__ == "__main__":
    main()
