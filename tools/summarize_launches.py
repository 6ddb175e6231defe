"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table (markdown)."""
import csv, collections, sys
src, out, title = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
skip = int(sys.argv[4]) if len(sys.argv) > 4 else -1  # launches to skip (warm-up); -1 = first half
with open(src) as f:
    rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
rows = rows[len(rows) // 2:] if skip < 0 else rows[skip:]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r["Kernel Name"].split("(")[0].replace("<unnamed>::", "")
    if "gemm_tn" in name:
        name = "gemm_tn_kernel<128x128, 8+4 warps>" if "384" in r["Block Size"] else "gemm_tn_kernel<64x64, 4+4 warps>"
    v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r["Metric Unit"]]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
with open(out, "w") as f:
    f.write(f"# {title}\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none` (per-launch times are cold-cache and serialised: compare shares).\n\n")
    f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% | {1e3 * v[1] / v[0]:.1f} |\n")
    f.write(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.3f} | 100% | |\n")
print(open(out).read())
