"""Developer tool: turn an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` log of
bench.py into profiles/roofline_traffic.json (average DRAM bytes per launch of the dominant kernel) and a launch list."""
import collections, csv, json, re, sys

src, variant, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
ki, mi, vi, ui, idi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
per = collections.defaultdict(dict)
names = {}
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
    per[r[idi]][r[mi]] = v * scale
    names[r[idi]] = re.sub(r"\(.*", "", r[ki])
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for i, m in per.items():
    a = agg[names[i]]
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0)
    a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print("| share | total us | launches | avg us | avg DRAM MB / launch | kernel |\n|---|---|---|---|---|---|")
for k, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {100 * t / tot:.2f}% | {t:.1f} | {n} | {t / n:.1f} | {b / n / 1e6:.1f} | `{k}` |")
dom = [k for k in agg if re.search(r"gemm_bf16_tcgen05_kernel<(\(int\))?256, (\(bool\))?(1|true)>", k)]
if dom:
    n, t, b = agg[dom[0]]
    json.dump({"variant": variant, "batch": batch, "kernel": dom[0], "launches": n, "dram_bytes_per_launch": b / n,
               "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, average over the launches of one bench run"},
              open("profiles/roofline_traffic.json", "w"), indent=1)
