import sys, time, random
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cases, hostsim, orc, wiremut
import test_hostsim as T
fds = open(os.path.join(ROOT, 'tests', 'golden', 'schemas.binpb'), 'rb').read()
O = orc.Schema(fds); H = hostsim.Schema(fds)
t0=time.time(); n=0
seed_base = int(sys.argv[1]) if len(sys.argv)>1 else 2000000
budget = float(sys.argv[2]) if len(sys.argv)>2 else 600
bad=0; rnd=0
while time.time()-t0 < budget and bad < 4:
    s0 = seed_base + rnd*1000; rnd += 1
    rng = random.Random(s0)
    for i,(name,js) in enumerate(cases.random_encode_cases(60, seed0=s0)):
        for k in range(3):
            b = cases.mutate_json(js, rng)
            if k == 2: b = cases.mutate_json(b, rng)
            try:
                ost, ow, _ = O.encode(name, b)
                est, ew = H.encode(name, b, i % 16, (i*3) % 16)
                assert (ost == 0) == (est == 0), ('enc-status', name, b, ost, est)
                if ost == 0: assert ow == ew, ('enc-bytes', name, b)
                else: pass
                T._check_coop_encode(H, name, b, i, k & 1); T._check_walk(H, name, b, i)
                n += 3
            except AssertionError as e:
                bad += 1; print('MISMATCH', s0, str(e)[:500])
    # damaged wire
    muts = [getattr(wiremut, m) for m in dir(wiremut) if callable(getattr(wiremut, m)) and m.startswith(('mut','shuffle','dup','trunc','inject','split_sub','reorder'))]
    for i,(name,w) in enumerate(cases.random_decode_cases(60, seed0=s0+500)):
        for k in range(3):
            b = bytearray(w)
            if len(b) and k == 0:
                b[rng.randrange(len(b))] = rng.randrange(256)
            elif len(b) and k == 1:
                del b[rng.randrange(len(b))]
            elif k == 2 and len(b) > 2:
                p = rng.randrange(len(b)); b[p:p] = bytes([rng.randrange(256)])
            b = bytes(b)
            try:
                T._check_decode(O, H, name, b, i, k & 1); T._check_coop_decode(H, name, b, i, k & 1); n += 2
            except AssertionError as e:
                bad += 1; print('MISMATCH', s0, str(e)[:500])
print('cases', n, 'rounds', rnd, 'mismatches', bad, 'secs', round(time.time()-t0))
