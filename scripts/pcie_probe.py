"""D2H / H2D copy rate into pinned host memory first-touched on each NUMA node (is the buffer's node the bound?)."""
import glob, os, time
import torch

def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out

bus = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), "pci_bus_id") else None
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
print("nodes:", [os.path.basename(n) for n in nodes])
for p in glob.glob("/sys/bus/pci/devices/*/numa_node"):
    try:
        cls = open(os.path.join(os.path.dirname(p), "class")).read().strip()
    except OSError:
        continue
    if cls.startswith("0x0302") or cls.startswith("0x0300"):
        print(p, open(p).read().strip())
n = 512 << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda")
all_cpus = os.sched_getaffinity(0)
for nd in nodes:
    cpus = set(cpulist(open(nd + "/cpulist").read())) & all_cpus
    if not cpus:
        continue
    os.sched_setaffinity(0, cpus)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.zero_()
    os.sched_setaffinity(0, all_cpus)
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            fn()
        torch.cuda.synchronize()
        print(os.path.basename(nd), name, "%.1f GB/s" % (4 * n / (time.perf_counter() - t0) / 1e9))
    del h

# both directions at once (two streams): is the link full duplex for large copies?
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_in.zero_()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out.zero_()
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for chunk in (n, 32 << 20, 8 << 20):
    def both():
        for o in range(0, n, chunk):
            with torch.cuda.stream(s1):
                d[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[o:o + chunk].copy_(d2[o:o + chunk], non_blocking=True)
    both(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        both()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("bidirectional, %d MB pieces: %.1f GB/s each way, %.1f GB/s combined" % (chunk >> 20, 4 * n / dt / 1e9, 8 * n / dt / 1e9))
