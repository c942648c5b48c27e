"""Per-kernel HBM table from an ncu summary (scripts/ncu_summary.py output): measured DRAM traffic / duration = achieved
GB/s against the measured copy peak, next to the multiplier-pipe utilisation - the per-kernel figure BASELINE.json's
north_star asks for ("every kernel ships with an ncu capture reporting achieved HBM GB/s against the B200 roofline").
   python scripts/kernel_hbm_table.py profiles/ncu_r02_full_summary.txt [peak GB/s] > profiles/kernel_hbm_table_r02.txt"""
import json, os, re, sys
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
path = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
if peak is None:
    mp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    peak = float(json.load(open(mp))["hbm_gbs"]) if os.path.exists(mp) else 6568.0     # the pool's measured copy bandwidth (bench.py's roofline.peak)
rows, cur = [], None
for line in open(path):
    if line.startswith("kernel:"):
        cur = {"name": re.sub(r"\(.*", "", line[len("kernel:"):].strip()).replace("void ", "")}
        rows.append(cur)
    elif cur is not None and line.startswith("  "):
        parts = line.split()
        key, val = parts[0], parts[-1]
        unit = parts[1] if len(parts) == 3 else ""
        try:
            v = float(val.replace(",", ""))
        except ValueError:
            continue
        cur[key] = v * UNIT.get(unit, 1.0) if "dram__bytes" in key else v
print("# achieved HBM bandwidth per kernel = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration.sum of the ncu --set full")
print("# capture in %s; peak = %.0f GB/s (measured copy bandwidth of this pool's B200s)." % (os.path.basename(path), peak))
print("# Every heavy kernel of this path is bound by the integer-multiply pipe or by dependency latency, not by HBM (DESIGN.md section 5).")
print("%-46s %9s %10s %9s %8s %10s %8s" % ("kernel (longest launch of one proof)", "ms", "DRAM MB", "GB/s", "of peak", "fmaheavy %", "issue %"))
for r in rows:
    ms = r.get("gpu__time_duration.sum", 0.0)
    b = r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)
    gbs = b / (ms * 1e-3) / 1e9 if ms else 0.0
    print("%-46s %9.3f %10.1f %9.1f %7.1f%% %10.1f %8.1f" % (r["name"][:46], ms, b / 1e6, gbs, 100.0 * gbs / peak,
          r.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0), r.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)))
