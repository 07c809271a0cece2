"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list of one
training step (tools/ncu_step.py) into a markdown table + per-kernel-family DRAM traffic (profiles/roofline_traffic.json).
    python tools/summarize_launches.py gpurun_out/launches.csv profiles/r01_launch_summary.md [profiles/roofline_traffic.json]
"""
import collections
import csv
import gzip
import json
import re
import sys

FAMILIES = {  # bench.py roofline family -> kernel-name regex
    "conv_halo_tcgen05": r"conv_halo_tcgen05_kernel", "wgrad_halo_tcgen05": r"wgrad_halo_tcgen05_kernel",
    "gemm_tcgen05": r"gemm_pers_tcgen05_kernel", "conv_pertap_tcgen05": r"nextou::gemm_tcgen05_kernel",
    "wgrad_tcgen05": r"nextou::wgrad_tcgen05_kernel", "knn_topk": r"knn_topk_kernel",
    "wgrad_planes_tcgen05": r"wgrad_planes_tcgen05_kernel", "conv_small": r"conv_small_(fwd|wgrad)_kernel",
}


def main(src, dst_md, dst_json=None):
    op = gzip.open if src.endswith(".gz") else open
    rows = [r for r in csv.reader(op(src, "rt", errors="replace")) if r]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    ix = {c: i for i, c in enumerate(h)}
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    units = {}
    for r in rows[hi + 1:]:
        if len(r) != len(h):
            continue
        name, metric, val = r[ix["Kernel Name"]], r[ix["Metric Name"]], r[ix["Metric Value"]]
        units[metric] = r[ix["Metric Unit"]]
        key = (r[ix["ID"]], name)
        try:
            per[key][metric] = float(val.replace(",", ""))
        except ValueError:
            pass

    def to_ms(v):
        u = units.get("gpu__time_duration.sum", "ns")
        return v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(u, 1e-6)

    def to_mb(metric, v):
        u = units.get(metric, "byte")
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)

    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for (_, name), m in per.items():
        short = re.sub(r"^void ", "", name)
        short = re.split(r"[<(]", short)[0]
        a = agg[short]
        a[0] += 1
        a[1] += to_ms(m.get("gpu__time_duration.sum", 0.0))
        a[2] += to_mb("dram__bytes_read.sum", m.get("dram__bytes_read.sum", 0.0))
        a[3] += to_mb("dram__bytes_write.sum", m.get("dram__bytes_write.sum", 0.0))
    total = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    own = sum(a[1] for k, a in agg.items() if k.startswith("nextou::"))
    lines = [f"{n} launches, {total:.1f} ms summed kernel time (per-launch times under ncu are cold-cache and serialised: compare SHARES).",
             "", "| kernel | launches | sum ms | share | DRAM read MB/launch | DRAM write MB/launch |", "|---|---:|---:|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        lines.append(f"| `{k[:70]}` | {a[0]} | {a[1]:.2f} | {100 * a[1] / total:.1f}% | {a[2] / a[0]:.1f} | {a[3] / a[0]:.1f} |")
    lines += ["", f"Own kernels (`nextou::*`): {own:.1f} ms = {100 * own / total:.1f}% of the summed kernel time; the rest are ATen "
                  "element-wise / reduction kernels of autograd glue (gradient accumulation, concat, dtype casts) and the optimizer."]
    open(dst_md, "a").write("\n".join(lines) + "\n")
    if dst_json:
        traffic = {}
        for fam, rx in FAMILIES.items():
            sel = [m for (_, name), m in per.items() if re.search(rx, name)]
            if sel:
                traffic[fam] = sum(m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0) for m in sel) / len(sel) \
                    * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units.get("dram__bytes_read.sum", "byte"), 1.0)
        json.dump(traffic, open(dst_json, "w"), indent=1)
    print("\n".join(lines[:14]))


if __name__ == "__main__":
    main(*sys.argv[1:4])
