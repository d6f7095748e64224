"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (profiles/)."""
import collections
import csv
import sys


def main(path, out, title):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0, None, None])
    for row in csv.DictReader(lines):
        try:
            val = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row.get("Metric Unit", "ns")
        us = val / 1000.0 if unit == "ns" else (val * 1000.0 if unit == "ms" else val)
        a = agg[row["Kernel Name"]]
        a[0] += 1
        a[1] += us
        a[2], a[3] = row["Block Size"], row["Grid Size"]
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as o:
        o.write(f"# {title}\n\nSource: `{path}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised "
                f"launches: compare SHARES, not absolutes).  {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms total.\n\n")
        o.write("| kernel | launches | total us | avg us | share | block | grid (last) |\n|---|---:|---:|---:|---:|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write(f"| `{k.split('(')[0][:70]}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.2f} | {v[1] / tot:.3f} | {v[2]} | {v[3]} |\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "ncu launch list")
