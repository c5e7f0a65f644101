"""Turn the raw ncu outputs under gpurun_out/ into the committed summaries under profiles/."""
import csv, json, os, subprocess, sys
RND = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
out, traffic = [], {}
for tag in ("emit", "gate", "heavy", "dualrc", "dualloop", "bbsearch"):
    rep = os.path.join(G, "prof_%s_%s.ncu-rep" % (tag, RND))
    if not os.path.isfile(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u, v = rows[0], rows[1], rows[-1]
    out.append("== prof_%s_%s.ncu-rep (ncu --set full --clock-control none, one steady-state launch)" % (tag, RND))
    out.append("  kernel: " + v[h.index("Kernel Name")])
    vals = {}
    for w in WANT:
        if w in h:
            out.append("  %-76s %s %s" % (w, v[h.index(w)], u[h.index(w)]))
            vals[w] = (float(v[h.index(w)].replace(",", "")), u[h.index(w)])
    def nbytes(key):
        x, unit = vals[key]
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    traffic[tag] = nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")
open(os.path.join(P, "ncu_%s_summary.txt" % RND), "w").write("\n".join(out) + "\n")
print("\n".join(out))
json.dump(traffic, open(os.path.join(G, "traffic_raw_%s.json" % RND), "w"))
print(traffic)
