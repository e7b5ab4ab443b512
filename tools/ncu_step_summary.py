"""ncu --csv launch list (gpu__time_duration.sum [+ dram bytes, tensor pipe]) of tools/run_step.py -> per-kernel summary.
  python tools/ncu_step_summary.py launches.csv > summary.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 8]
hdr = rows[0]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
launch = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    d = launch.setdefault(r[ii], {"name": r[ki].split("(")[0].replace("tise::", "").replace("void ", "")})
    d[r[mi]] = v
agg = collections.OrderedDict()
for d in launch.values():
    a = agg.setdefault(d["name"], {"n": 0, "us": 0.0, "dram": 0.0, "pipe_w": 0.0})
    t = d.get("gpu__time_duration.sum", 0.0) / 1e3
    a["n"] += 1
    a["us"] += t
    a["dram"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a["pipe_w"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
tot = sum(a["us"] for a in agg.values())
print("%-44s %5s %11s %7s %10s %9s %7s" % ("kernel", "n", "total_us", "share", "dram_MB", "GB/s", "pipe%"))
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    print("%-44s %5d %11.1f %7.3f %10.1f %9.0f %7.1f" % (n[:44], a["n"], a["us"], a["us"] / tot, a["dram"] / 1e6,
                                                       a["dram"] / max(a["us"], 1e-9) / 1e3, a["pipe_w"] / max(a["us"], 1e-9)))
print("total_us %.1f  launches %d" % (tot, sum(a["n"] for a in agg.values())))
