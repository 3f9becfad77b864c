"""Summarise an ncu --csv launch list (gpu__time_duration.sum) by kernel and by (kernel, grid)."""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"].split("(")[0], r.get("Grid Size", ""), r.get("Block Size", ""), ns))
    total = sum(r[3] for r in rows)
    print(f"{len(rows)} launches, total {total / 1e6:.3f} ms")
    by = defaultdict(lambda: [0, 0.0])
    for k, g, b, ns in rows:
        by[k][0] += 1
        by[k][1] += ns
    print("\n== by kernel")
    for k, (n, ns) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print(f"{ns / 1e6:9.3f} ms {100 * ns / total:5.1f}%  x{n:4d}  {k[:90]}")
    byg = defaultdict(lambda: [0, 0.0])
    for k, g, b, ns in rows:
        byg[(k, g)][0] += 1
        byg[(k, g)][1] += ns
    print("\n== by (kernel, grid), top", top)
    for (k, g), (n, ns) in sorted(byg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{ns / 1e6:9.3f} ms {100 * ns / total:5.1f}%  x{n:3d} avg {ns / n / 1e3:8.1f} us  grid {g:18s} {k[:60]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
