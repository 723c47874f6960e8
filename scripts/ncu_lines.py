"""Per-source-line instruction counts and stall samples of one kernel from an ncu report.

ncu's CSV source page carries the metrics per SASS instruction only; this joins it (by instruction
order) with `nvdisasm -gi` of the object the kernel came from and sums per line of --file (the
outermost inlined frame in that file), so a hot line of a header shows up with everything it inlines.

usage: python scripts/ncu_lines.py REPORT.ncu-rep OBJECT.o KERNEL_SUBSTRING --file ggr_walk.cuh [--items N] [--launch K]
"""
import argparse
import collections
import csv
import glob
import io
import os
import re
import subprocess
import tempfile


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("obj")
    ap.add_argument("kernel")
    ap.add_argument("--file", default="")
    ap.add_argument("--items", type=int, default=1)
    ap.add_argument("--top", type=int, default=60)
    ap.add_argument("--section", default="", help="substring of the mangled name (object side) when KERNEL is a regex or matches several instances")
    ap.add_argument("--inner", action="store_true", help="attribute to the innermost frame (the line itself) instead of the outermost one in --file")
    a = ap.parse_args()
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(a.obj)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    # instructions of the kernel with their line chains
    insts = []
    inside = False
    chain = []
    fresh = False
    for ln in dis:
        if ln.startswith("\t.section") or ln.startswith("//-----"):
            inside = (".text." in ln) and ((a.section or a.kernel) in ln)
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            if not fresh:
                chain = []
                fresh = True
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            insts.append((int(m.group(1), 16), m.group(2).strip(), list(chain)))
            fresh = False
    out = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--kernel-name", "regex:" + a.kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    launches = []
    cur = None
    hdr = None
    for r in rows:
        if len(r) >= 2 and r[0] == "Kernel Name":
            cur = []
            launches.append(cur)
        elif len(r) > 5 and r[0] == "Address":
            hdr = r
        elif len(r) > 5 and cur is not None:
            cur.append(r)
    sass = launches[0]
    ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
    te = hdr.index("Thread Instructions Executed")
    if len(sass) != len(insts):
        print("warning: %d SASS rows in the report, %d instructions in the object" % (len(sass), len(insts)))
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r, (_, _, ch) in zip(sass, insts):
        key = None
        for f, l in (ch if a.inner else reversed(ch)):  # chain: innermost first
            if not a.file or f == a.file:
                key = (f, l)
                break
        if key is None:
            key = ch[0] if ch else ("?", 0)
        v = (int(r[ie]), int(r[sm]), int(r[te]))
        for k in range(3):
            agg[key][k] += v[k]
            tot[k] += v[k]
    print("total: %.1f warp instructions per item, %d samples, %.1f threads per instruction" % (tot[0] / a.items, tot[1], tot[2] / max(tot[0], 1)))
    src = {}
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: a.top]:
        if f not in src:
            for root in ("ggrmcp_b200/csrc", "."):
                p = os.path.join(root, f)
                if os.path.exists(p):
                    src[f] = open(p, errors="replace").read().splitlines()
                    break
            else:
                src[f] = []
        text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
        print("%8.1f inst/item %6d smp %5.1f thr  %s:%d  %s" % (v[0] / a.items, v[1], v[2] / max(v[0], 1), f, l, text))


if __name__ == "__main__":
    main()
