"""Roll-up of an ncu launch list (csv, one row per launch and metric) by kernel name:
   python scripts/launch_summary.py gpurun_out/launches.csv > profiles/launches_rNN.txt"""
import collections
import csv
import sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = list(csv.reader(rows))
h = r[0]
ki, mi, vi, ui, ii = (h.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
launches = collections.OrderedDict()
for x in r[1:]:
    d = launches.setdefault(x[ii], {"name": x[ki]})
    v = float(x[vi].replace(",", ""))
    if x[mi] == "gpu__time_duration.sum":
        v = {"ns": v / 1e6, "us": v / 1e3, "ms": v, "s": v * 1e3, "nsecond": v / 1e6, "usecond": v / 1e3, "msecond": v, "second": v * 1e3}[x[ui]]
    if x[mi].startswith("dram__bytes"):
        v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(x[ui], 1)
    d[x[mi]] = v
agg = collections.defaultdict(lambda: collections.defaultdict(float))
for d in launches.values():
    n = d["name"].split("(")[0][:58]
    a = agg[n]
    a["n"] += 1
    for k, v in d.items():
        if k != "name":
            a[k] += v
T = sum(a["gpu__time_duration.sum"] for a in agg.values())
F = sum(a["sm__inst_executed_pipe_fmaheavy.sum"] for a in agg.values())
print("# ncu launch list of ONE proof (witness batch of 1 + prove) on a single lane; kernels serialised and cold-cache, so the")
print("# SHARE of the proof is what carries over to the live stage_ms of bench.py, not the absolute times.")
print("# fmaheavy / alu / fmalite = warp instructions executed on the integer-multiply (IMAD.WIDE), ALU and FMA-lite pipes")
print("%-58s %4s %8s %6s %11s %6s %11s %11s %11s %8s" % ("kernel", "n", "ms", "%t", "fmaheavy", "%f", "alu", "fmalite", "warp inst", "dram GB"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    dram = a.get("dram__bytes_read.sum", 0) + a.get("dram__bytes_write.sum", 0)
    print("%-58s %4d %8.3f %5.1f%% %11.0f %5.1f%% %11.0f %11.0f %11.0f %8.2f" % (
        k, a["n"], a["gpu__time_duration.sum"], 100 * a["gpu__time_duration.sum"] / T, a["sm__inst_executed_pipe_fmaheavy.sum"],
        100 * a["sm__inst_executed_pipe_fmaheavy.sum"] / F if F else 0, a.get("sm__inst_executed_pipe_alu.sum", 0),
        a.get("sm__inst_executed_pipe_fmalite.sum", 0), a.get("smsp__inst_executed.sum", 0), dram / 1e9))
print("TOTAL %.3f ms, fmaheavy %.0f" % (T, F))
