"""CPU arm on the GPU box's host: how the oracle port scales with threads (what `cpu_baseline.scaling_eff` summarises),
next to what the container is allowed to use (cgroup quota, affinity)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, benchgen, orc
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us", "/proc/loadavg"):
    try: print(f, open(f).read().strip())
    except Exception as e: print(f, "-")
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try: print([l.strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0])
except Exception: pass
O = orc.Schema(open(os.path.join(ROOT, "tests", "golden", "schemas.binpb"), "rb").read())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
wl = benchgen.nested(n, O.msg)
base = None
for th in (1, 4, 16, 32, 64, 128):
    m = n if th > 1 else 4096
    t0 = time.perf_counter()
    O.encode_batch(wl.req_msg[:m], wl.req_json, wl.req_off[:m + 1], threads=th)
    O.decode_batch(wl.rep_msg[:m], wl.rep_wire, wl.rep_off[:m + 1], threads=th)
    dt = time.perf_counter() - t0
    r = m / dt
    base = base or r
    print("threads %3d: %8.0f transcodes/s  x%.1f" % (th, r, r / base))
