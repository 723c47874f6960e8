"""Does the alignment of the HOST address of a chunk copy matter?  24 MB copies both ways at once (the shape of the host
pipeline's payload copies) from / to page-aligned host addresses and from / to odd ones."""
import time, torch
n = 24 << 20
reps = 40
h_in = torch.empty(n * 2 + 4096, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n * 2 + 4096, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n + 4096, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n + 4096, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, ho, do in (("aligned", 0, 0), ("host +13", 13, 0), ("host +13, device +13", 13, 13), ("host +2061", 2061, 13), ("aligned again", 0, 0)):
    for timed in (False, True):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for r in range(reps if timed else 3):
            with torch.cuda.stream(s1):
                d_a[do:do + n].copy_(h_in[ho:ho + n], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[ho:ho + n].copy_(d_b[do:do + n], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print("%-24s %.1f GB/s each way" % (name, reps * n / dt / 1e9))
