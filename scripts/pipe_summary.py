"""Per-kernel totals of one proof from an ncu launch list taken with
   --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio
(kernels after the last witness_kernel launch = the single-lane profiling step of bench.py --batch 1)."""
import csv, collections, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = list(csv.reader(rows)); h = r[0]
ki, mi, vi, ui, ii = (h.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
launches = collections.OrderedDict()
for x in r[1:]:
    d = launches.setdefault(x[ii], {"name": x[ki]})
    v = float(x[vi].replace(",", ""))
    if x[mi] == "gpu__time_duration.sum":
        v = {"ns": v / 1e6, "us": v / 1e3, "ms": v, "s": v * 1e3}[x[ui]]
    d[x[mi]] = v
L = list(launches.values())
idx = [i for i, d in enumerate(L) if d["name"].startswith("witness_kernel")]
seg = L[idx[-1]:]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
for d in seg:
    n = d["name"].split("(")[0][:60]
    a = agg[n]; a[0] += 1; a[1] += d["gpu__time_duration.sum"]; a[2] += d.get("sm__inst_executed_pipe_fmaheavy.sum", 0); a[3] += d.get("smsp__inst_executed.sum", 0)
    a[4] += d.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0) * d.get("smsp__inst_executed.sum", 0)
T = sum(a[1] for a in agg.values()); F = sum(a[2] for a in agg.values())
print("# one email (witness batch of 1 + prove, single lane), kernels serialised by ncu; fmaheavy = warp instructions on the")
print("# integer-multiply pipe (sm__inst_executed_pipe_fmaheavy.sum), the resource that bounds the engine (DESIGN.md section 5);")
print("# lanes = active threads per executed warp instruction (32 = no divergence)")
print("%-62s %4s %9s %6s %13s %6s %13s %6s" % ("kernel", "n", "ms", "%t", "fmaheavy", "%f", "warp inst", "lanes"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print("%-62s %4d %9.3f %5.1f%% %13.0f %5.1f%% %13.0f %6.1f" % (k, a[0], a[1], 100 * a[1] / T, a[2], 100 * a[2] / F, a[3], a[4] / a[3] if a[3] else 0))
print("TOTAL %.3f ms, fmaheavy %.0f" % (T, F))
