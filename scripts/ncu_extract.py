"""Key metrics of every kernel of an `ncu --set full` report as a small CSV (one column per kernel).
usage: python scripts/ncu_extract.py REPORT.ncu-rep OUT.csv"""
import csv
import io
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keep = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"] + sorted(h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio"))
seen, cols = set(), []
for r in rows[2:]:
    name = re.match(r"(?:void )?([\w<>, ]+?)\(", r[idx["Kernel Name"]])
    name = name.group(1) if name else r[idx["Kernel Name"]][:40]
    if name in seen:
        continue
    seen.add(name)
    cols.append((name, r))
with open(out, "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["metric", "unit"] + [c[0] for c in cols])
    for h in keep:
        if h in idx:
            w.writerow([h.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "_per_issue"), units[idx[h]]] + [c[1][idx[h]] for c in cols])
print("wrote", out, [c[0] for c in cols])
