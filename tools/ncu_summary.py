#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --page raw --csv) into a markdown table of the metrics the roofline argument uses."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time_us"),
    ("sm__cycles_elapsed.max.per_second", "sm_ghz"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_active_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_thru_%"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l2_to_sm"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lsu_smem_%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_%"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("smsp__inst_executed.sum", "inst"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    names = [k for k, _ in KEYS if k in col]
    print("| kernel | grid | " + " | ".join(dict(KEYS)[k] for k in names) + " |")
    print("|---|---|" + "---|" * len(names))
    for r in data:
        kn = r[col["Kernel Name"]]
        kn = kn.split("(")[0].replace("void ", "").replace("unnamed>::", "")[-70:]
        vals = []
        for k in names:
            v, u = r[col[k]], units[col[k]]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            vals.append(f"{v} {u}".strip())
        print(f"| `{kn}` | {r[col['Grid Size']]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
