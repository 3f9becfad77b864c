"""Per-kernel table from an ncu launch list taken with
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
(one step under tools/ncu_step.py): time, share, launches, DRAM MB per launch, DRAM TB/s, time-weighted tensor-pipe %.
Optionally writes the GEMM-class DRAM traffic per launch (bench.py's roofline.traffic) to a JSON file.

    python tools/summarize_ncu_metrics.py gpurun_out/launches.csv [traffic.json key]"""
import csv
import json
import os
import sys
from collections import defaultdict


def main(path, out_json=None, key=None):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    per = defaultdict(dict)  # launch id -> {metric: value, "name": ...}
    for r in csv.DictReader(lines):
        i = int(r["ID"])
        per[i]["name"] = r["Kernel Name"].split("(")[0][:60]
        v = float(r["Metric Value"].replace(",", ""))
        u = r.get("Metric Unit", "")
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)  # -> us
        elif m.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        per[i][m] = v
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])  # n, us, bytes, tensor% * us
    for d in per.values():
        a = agg[d["name"]]
        us = d.get("gpu__time_duration.sum", 0.0)
        a[0] += 1
        a[1] += us
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a[3] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * us
    total = sum(a[1] for a in agg.values())
    print(f"{len(per)} launches in one step, serialised by ncu (cold caches): total {total / 1e3:.3f} ms")
    print("   ms     share   n   avg us   DRAM MB/launch  DRAM TB/s  tensor-pipe% (time-weighted)  kernel")
    for k, (n, us, by, tp) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us / 1e3:7.3f}  {100 * us / total:5.1f}%  {n:3d}  {us / n:7.1f}  {by / n / 1e6:11.1f}  {by / us / 1e6 if us else 0:9.2f}"
              f"  {tp / us if us else 0:9.1f}   {k}")
    gem = [(n, us, by) for k, (n, us, by, _tp) in agg.items() if "dm_tapgemm" in k]
    n = sum(g[0] for g in gem)
    by = sum(g[2] for g in gem)
    if n:
        print(f"GEMM-class kernel: {n} launches, DRAM traffic {by / 1e6:.1f} MB per step = {by / n / 1e6:.2f} MB per launch")
    if out_json and key and n:
        doc = json.load(open(out_json)) if os.path.exists(out_json) else {}
        doc[key] = {"bytes_per_launch": round(by / n), "launches": n,
                    "source": f"profiles/{os.path.basename(path)} (ncu, one step, dram__bytes_read.sum + dram__bytes_write.sum)"}
        json.dump(doc, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else None)
