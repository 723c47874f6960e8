"""Static SASS instruction count per source line of --file for one kernel of an object (code size attribution).
usage: python scripts/sass_lines.py OBJECT.o KERNEL_SUBSTRING --file ggr_walk.cuh"""
import argparse, collections, glob, os, re, subprocess, tempfile
ap = argparse.ArgumentParser()
ap.add_argument("obj"); ap.add_argument("kernel"); ap.add_argument("--file", default=""); ap.add_argument("--top", type=int, default=40)
a = ap.parse_args()
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(a.obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
agg = collections.Counter(); inside = False; chain = []; fresh = False; total = 0
for ln in dis:
    if ln.startswith("\t.section") or ln.startswith("//-----"):
        inside = (".text." in ln) and (a.kernel in ln); continue
    if not inside: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        if not fresh: chain = []; fresh = True
        chain.append((os.path.basename(m.group(1)), int(m.group(2)))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+", ln):
        key = None
        for f, l in reversed(chain):
            if not a.file or f == a.file: key = (f, l); break
        if key is None: key = chain[0] if chain else ("?", 0)
        agg[key] += 1; total += 1; fresh = False
print("total instructions:", total, "=", total * 16, "bytes")
for (f, l), n in agg.most_common(a.top):
    text = ""
    for root in ("ggrmcp_b200/csrc", "."):
        p = os.path.join(root, f)
        if os.path.exists(p):
            L = open(p, errors="replace").read().splitlines()
            if 0 < l <= len(L): text = L[l - 1].strip()[:100]
            break
    print("%6d  %s:%d  %s" % (n, f, l, text))
