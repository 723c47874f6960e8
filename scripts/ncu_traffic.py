"""Writes / updates profiles/ncu_traffic.json from an `ncu --set full` report: DRAM bytes read + written per launch
of every kernel in it, keyed by kernel function, workload, items and the hash of the kernel sources the report was
taken on (bench.py reports `roofline.traffic` from this table only when the hash matches the build it runs).

usage: python scripts/ncu_traffic.py REPORT.ncu-rep --workload nested --items 151552 [--copy-to profiles/NAME.ncu-rep]"""
import argparse
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--workload", required=True)
    ap.add_argument("--items", type=int, required=True)
    ap.add_argument("--csv-out", default="")
    a = ap.parse_args()
    import bench
    sha = bench.source_sha()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    table = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    entries = json.load(open(table)) if os.path.exists(table) else []
    seen = {}
    keep = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"] + [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
    extract = [["metric", "unit"]]
    kernels = []
    for r in rows[2:]:
        name = re.match(r"(?:void )?(\w+)", r[idx["Kernel Name"]]).group(1)
        if name in seen:
            continue
        seen[name] = r
        kernels.append(name)

        def val(col):
            v, u = float(r[idx[col]].replace(",", "")), rows[1][idx[col]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        e = {"kernel": name, "workload": a.workload, "items": a.items, "dram_read": val("dram__bytes_read.sum"),
             "dram_write": val("dram__bytes_write.sum"), "source_sha": sha, "report": os.path.basename(a.csv_out or a.report)}
        entries = [x for x in entries if not (x["kernel"] == name and x["workload"] == a.workload and x["items"] == a.items)] + [e]
    json.dump(entries, open(table, "w"), indent=1)
    if a.csv_out:
        for h in keep:
            if h in idx:
                extract.append([h, rows[1][idx[h]]] + [seen[k][idx[h]] for k in kernels])
        extract[0] += kernels
        with open(a.csv_out, "w", newline="") as fh:
            csv.writer(fh).writerows(extract)
    print("updated %s: %d kernels, source %s" % (table, len(kernels), sha))


if __name__ == "__main__":
    main()
