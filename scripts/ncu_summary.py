"""Summarise an `ncu --set full` report (read with `ncu -i ... --page raw --csv`): the key throughput / stall metrics per
captured launch, plus the hottest SASS lines by stall samples.   usage: ncu_summary.py <report.ncu-rep> [n_lines]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_active.avg", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "derived__lts__lts2xbar_bytes.sum.per_second", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("=" * 100)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("%-75s %s %s" % (w, r[i], units[i]))
    st = sorted(((float(r[hdr.index(h)] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)
    print("stall cycles per issued instruction:", ", ".join("%s %.2f" % (n, v) for v, n in st[:9]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"rows": []}
        kern.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
if kern:
    k = kern[0]
    h = k["hdr"]
    si, ni, ei = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    tot = sum(int(r[ni]) for r in k["rows"])
    print("=" * 100)
    print("hottest SASS lines of launch 0 (stall samples, %d total; executed count)" % tot)
    top = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][ni]))[:nl]
    for i in sorted(top):
        r = k["rows"][i]
        print("%6d %5d %9s  %s" % (i, int(r[ni]), r[ei], r[si][:90]))
