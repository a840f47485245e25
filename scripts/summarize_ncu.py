#!/usr/bin/env python
"""Turn `ncu -i x.ncu-rep --page raw --csv` exports into the markdown tables kept under profiles/.

    python scripts/summarize_ncu.py raw.csv [--launches launches.csv] > profiles/summary.md
"""
import argparse
import csv
import re
import sys
from collections import OrderedDict, defaultdict

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1TEX %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def short(name):
    name = re.sub(r"void |\(anonymous namespace\)::|<unnamed>::", "", name)
    return re.sub(r"\(.*$", "", name)


def fmt(v, unit):
    try:
        f = float(v.replace(",", ""))
    except ValueError:
        return v
    if unit in ("byte", "Kbyte", "Mbyte", "Gbyte", "Tbyte"):
        f *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
        return f"{f / 1e9:.3f} GB" if f >= 1e8 else f"{f / 1e6:.2f} MB"
    if unit in ("ns", "us", "ms", "s"):
        f *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit]
        return f"{f:.4f} ms"
    if unit == "%":
        return f"{f:.1f}"
    return f"{f:g}" + (f" {unit}" if unit and unit not in ("register/thread",) else "")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("--launches")
    args = ap.parse_args()
    rows = list(csv.reader(open(args.raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for r in rows[2:]:
        cells = [fmt(r[idx[c]], units[idx[c]]) if c in idx else "-" for c, _ in COLS]
        print(f"| `{short(r[idx['Kernel Name']])}` | " + " | ".join(cells) + " |")
    if args.launches:
        lr = [r for r in csv.reader(open(args.launches)) if len(r) > 5]
        h = lr[0]
        kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
        agg = defaultdict(lambda: [0, 0.0])
        order = OrderedDict()
        for r in lr[1:]:
            t = float(r[mv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu], 1e-6)
            k = short(r[kn])
            agg[k][0] += 1
            agg[k][1] += t
            order.setdefault(k, None)
        tot = sum(v[1] for v in agg.values())
        print("\n| kernel (all launches of the command) | launches | total ms | mean ms | share |")
        print("|---|---|---|---|---|")
        for k in sorted(agg, key=lambda k: -agg[k][1]):
            n, t = agg[k]
            print(f"| `{k}` | {n} | {t:.3f} | {t / n:.4f} | {100 * t / tot:.1f} % |")


if __name__ == "__main__":
    main()
