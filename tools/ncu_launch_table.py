"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active...]`
log: per kernel name launches, total / average duration, share, and (when present) DRAM MB per launch and the
duration-weighted tensor-pipe activity."""
import csv
import sys
from collections import OrderedDict


def main(path, skip=0, take=None):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    per = OrderedDict()
    for r in rd:
        per.setdefault(r["ID"], {"name": r["Kernel Name"]})[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * (
            {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1.0) if "time" in r["Metric Name"] else
            {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(r["Metric Unit"], 1.0))
    launches = list(per.values())[skip: (skip + take) if take else None]
    agg = OrderedDict()
    for l in launches:
        a = agg.setdefault(l["name"].split("(")[0], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "tw": 0.0})
        a["n"] += 1
        a["us"] += l.get("gpu__time_duration.sum", 0.0)
        a["rd"] += l.get("dram__bytes_read.sum", 0.0)
        a["wr"] += l.get("dram__bytes_write.sum", 0.0)
        a["tw"] += l.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * l.get("gpu__time_duration.sum", 0.0)
    tot = sum(a["us"] for a in agg.values())
    print("| share | total us | launches | avg us | DRAM MB / launch | tensor pipe % (time-weighted) | kernel |")
    print("|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        print(f"| {100 * a['us'] / tot:.2f}% | {a['us']:.1f} | {a['n']} | {a['us'] / a['n']:.1f} | {(a['rd'] + a['wr']) / a['n']:.1f} | "
              f"{a['tw'] / a['us'] if a['us'] else 0:.1f} | `{k}` |")
    tw = sum(a["tw"] for a in agg.values())
    print(f"\ntotal {tot:.1f} us over {len(launches)} launches; duration-weighted tensor-pipe activity {tw / tot if tot else 0:.1f} %")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else None)
