"""Turn an `ncu --page raw --csv` dump into a markdown table (one row per profiled launch).

    python tools/ncu_summary.py gpurun_out/x_raw.csv "title" > profiles/x.md
"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 (LTS) %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1TEX %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
]
WITH_UNIT = {"gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"}


def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {k: hdr.index(k) for k, _ in COLS if k in hdr}
    kn = hdr.index("Kernel Name")
    print(f"# {title}\n")
    print("Source: `ncu --set full --clock-control none` on one B200 (cold-cache, serialised launches; read the")
    print("percentages and bytes, not the absolute times -- timing lives in bench.py / tools/bench_all.py).\n")
    print("| kernel | " + " | ".join(n for k, n in COLS if k in ix) + " |")
    print("|---|" + "---|" * len(ix))
    for d in data:
        name = d[kn].replace("<unnamed>::", "").replace("void ", "").split("(")[0]
        cells = []
        for k, _ in COLS:
            if k not in ix:
                continue
            v, u = d[ix[k]], units[ix[k]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.3f}" if f < 100 else f"{f:.0f}"
            except ValueError:
                pass
            cells.append(f"{v} {u}" if k in WITH_UNIT else v)
        print(f"| `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
