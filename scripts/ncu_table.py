"""Turn an `ncu --csv` log (metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum,
dram__throughput.avg.pct_of_peak_sustained_elapsed, optionally sm__pipe_tensor_cycles_active...) into a markdown table:
one row per kernel name with launches, total / median duration, DRAM bytes per launch and achieved DRAM GB/s.

    python scripts/ncu_table.py gpurun_out/streaming_ncu.csv [--filter lit::] > profiles/...md
"""
import collections
import csv
import statistics
import sys


def main():
    path = sys.argv[1]
    filt = sys.argv[sys.argv.index("--filter") + 1] if "--filter" in sys.argv else ""
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10 and r[0].isdigit()]
    per = collections.OrderedDict()
    for r in rows:
        name, metric, unit, val = r[4], r[-3], r[-2], r[-1].replace(",", "")
        if filt and filt not in name:
            continue
        short = name.split("(")[0].replace("void ", "")[:90]
        key = (r[0], short)
        d = per.setdefault(key, {})
        try:
            v = float(val)
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
                 "%": 1.0}.get(unit, 1.0)
        d[metric] = v * scale
    agg = collections.OrderedDict()
    for (_, short), d in per.items():
        agg.setdefault(short, []).append(d)
    print("| kernel | launches | total ms | median us | DRAM MB / launch (read + write) | DRAM GB/s | dram % of peak (ncu) | tensor pipe % |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|")
    for short, ds in sorted(agg.items(), key=lambda kv: -sum(d.get("gpu__time_duration.sum", 0) for d in kv[1])):
        t = [d.get("gpu__time_duration.sum", 0.0) for d in ds]
        by = [d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in ds]
        pct = [d["dram__throughput.avg.pct_of_peak_sustained_elapsed"] for d in ds
               if "dram__throughput.avg.pct_of_peak_sustained_elapsed" in d]
        tp = [d[k] for d in ds for k in d if k.startswith("sm__pipe_tensor_cycles_active") or k.startswith("sm__pipe_tensor_op")]
        gbs = sum(by) / max(sum(t), 1e-9) / 1e3
        print(f"| `{short}` | {len(ds)} | {sum(t) / 1e3:.3f} | {statistics.median(t):.1f} | {statistics.median(by) / 1e6:.1f} | "
              f"{gbs:.0f} | {statistics.median(pct) if pct else float('nan'):.1f} | "
              f"{statistics.median(tp) if tp else float('nan'):.1f} |")


if __name__ == "__main__":
    main()
