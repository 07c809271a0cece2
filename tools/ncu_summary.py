"""Markdown table from an `ncu --set full` report (runs `ncu -i <rep> --page raw --csv`; no GPU needed).
    python tools/ncu_summary.py gpurun_out/r02_kernels.ncu-rep > profiles/r02_ncu_kernels.md"""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time µs", 1e-3), ("dram__bytes_read.sum", "DRAM read MB", 1e-6), ("dram__bytes_write.sum", "DRAM write MB", 1e-6),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak", 1), ("lts__t_sector_hit_rate.pct", "L2 hit %", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %", 1),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr. %", 1), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %", 1),
        ("launch__registers_per_thread", "regs", 1), ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1)]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {c: i for i, c in enumerate(hdr)}
    unit_scale = {"ns": 1.0, "us": 1e3, "usecond": 1e3, "nsecond": 1.0, "ms": 1e6, "msecond": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    print("| kernel | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|" + "---:|" * len(COLS))
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("nextou::", "")
        cells = []
        for key, _, scale in COLS:
            if key not in ix or r[ix[key]] == "":
                cells.append("-")
                continue
            v = float(r[ix[key]].replace(",", ""))
            if scale != 1:
                v *= unit_scale.get(units[ix[key]], 1.0) * scale
            cells.append(f"{v:.1f}" if abs(v) < 1e5 and v != int(v) else f"{int(v)}")
        print(f"| `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
