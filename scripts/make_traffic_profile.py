#!/usr/bin/env python
"""Turn an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch, CSV)
of `bench.py --timesteps 3` into the per-diffusion-step DRAM traffic summary that bench.py reports as
roofline.traffic.   usage: make_traffic_profile.py <launches.csv> <out.json> [batch]"""
import collections
import csv
import json
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    lines = open(src).read().splitlines()
    k0 = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    by = collections.OrderedDict()
    for r in csv.DictReader(lines[k0:]):
        if r.get("Metric Value") is None:
            continue
        by.setdefault(r["ID"], {"name": r["Kernel Name"]})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    ids = list(by)
    names = [by[i]["name"] for i in ids]
    starts = [k for k, n in enumerate(names) if "begin_step" in n]
    a, b = starts[0], starts[1]
    agg = collections.OrderedDict()
    for k in range(a, b):
        d = by[ids[k]]
        n = d["name"].split("(")[0].replace("void ", "").replace("ds::", "")
        g = agg.setdefault(n, {"launches": 0, "time_ns": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        g["launches"] += 1
        g["time_ns"] += d["gpu__time_duration.sum"]
        g["dram_read_bytes"] += d["dram__bytes_read.sum"]
        g["dram_write_bytes"] += d["dram__bytes_write.sum"]
    out = {
        "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                  "python bench.py --steps 1 --warmup 1 --batch %d --timesteps 3 --no-cpu-baseline (%s)" % (batch, src),
        "note": "one diffusion step (launches %d..%d of the capture); ncu serialises launches, so times are cold-cache "
                "and without PDL overlap" % (a, b - 1),
        "batch": batch, "launches_per_step": b - a,
        "step_time_ns": sum(g["time_ns"] for g in agg.values()),
        "step_dram_bytes": sum(g["dram_read_bytes"] + g["dram_write_bytes"] for g in agg.values()),
        "kernels": agg,
    }
    json.dump(out, open(dst, "w"), indent=1)
    print("step: %d launches, %.3f ms (serialised), %.2f GB DRAM" % (b - a, out["step_time_ns"] / 1e6, out["step_dram_bytes"] / 1e9))
    for n, g in sorted(agg.items(), key=lambda x: -x[1]["time_ns"]):
        print("  %-42s n=%3d  %8.1f us  %8.1f MB" % (n[:42], g["launches"], g["time_ns"] / 1e3,
                                                     (g["dram_read_bytes"] + g["dram_write_bytes"]) / 1e6))


if __name__ == "__main__":
    main()
