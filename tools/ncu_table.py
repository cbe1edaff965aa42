"""Markdown table of the per-kernel figures quoted in profiles/*_ncu_summary.md from an `ncu --set full` report:

    ncu -i gpurun_out/r2_prof_final.ncu-rep --page raw --csv > gpurun_out/r2_prof_final_raw.csv
    python tools/ncu_table.py gpurun_out/r2_prof_final_raw.csv
"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "time us", 1e-3),          # ns -> us when the unit row says nsecond
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", 1),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("dram__bytes_read.sum", "DRAM read MB", 1),
    ("dram__bytes_write.sum", "DRAM write MB", 1),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | grid x block | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("fb::", "")
        cells = []
        for key, _, _ in COLS:
            if key not in idx:
                cells.append("n/a")
                continue
            v, u = float(r[idx[key]].replace(",", "")), units[idx[key]]
            if u == "nsecond" or u == "ns":
                v /= 1e3
            elif u == "msecond":
                v *= 1e3
            if u == "byte":
                v /= 1e6
            elif u == "Kbyte":
                v /= 1e3
            elif u == "Gbyte":
                v *= 1e3
            cells.append(f"{v:.1f}" if v < 1000 else f"{v:.0f}")
        print(f"| `{name}` | {r[idx['launch__grid_size']]} x {r[idx['launch__block_size']]} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
