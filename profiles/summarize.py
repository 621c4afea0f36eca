#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/r1_launches.csv profiles/r1_launches_summary.md "title"
    python profiles/summarize.py full     gpurun_out/r1_prof.ncu-rep profiles/r1_ncu_full_summary.md "title"
"""
import collections
import csv
import subprocess
import sys


def launches(src, dst, title):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0][:70]
        t = float(r[vi].replace(",", ""))
        if r[ui] == "ns":
            t /= 1e3
        elif r[ui] == "ms":
            t *= 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nPer-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {t:.0f} | {t / c:.1f} | {100 * t / tot:.1f}% |\n")
        f.write(f"\ntotal {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches\n")


WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"),
]


def full(src, dst, title):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    idx = [(hdr.index(m), lab) for m, lab in WANT if m in hdr]
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on`; one row per captured launch.\n\n")
        f.write("| kernel | " + " | ".join(f"{lab} [{units[i]}]" if units[i] else lab for i, lab in idx) + " |\n")
        f.write("|---|" + "---|" * len(idx) + "\n")
        for r in rows[2:]:
            name = r[ki].split("(")[0][:60]
            vals = []
            for i, _ in idx:
                try:
                    vals.append(f"{float(r[i].replace(',', '')):.4g}")
                except ValueError:
                    vals.append(r[i])
            f.write(f"| `{name}` | " + " | ".join(vals) + " |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:5])
