import sys, time, random
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cases, hostsim, orc, wiremut
import test_hostsim as T
fds = open(os.path.join(ROOT, 'tests', 'golden', 'schemas.binpb'), 'rb').read()
O = orc.Schema(fds); H = hostsim.Schema(fds)
t0=time.time(); n=0
seed_base = int(sys.argv[1]) if len(sys.argv)>1 else 900000
budget = float(sys.argv[2]) if len(sys.argv)>2 else 600
bad=0
rnd = 0
while time.time()-t0 < budget:
    s0 = seed_base + rnd*1000; rnd += 1
    try:
        for i,(name,js) in enumerate(cases.random_encode_cases(80, seed0=s0)):
            T._check_encode(O,H,name,js,i); T._check_coop_encode(H,name,js,i); T._check_walk(H,name,js,i); n+=3
        for i,(name,w) in enumerate(cases.random_decode_cases(80, seed0=s0+500)):
            T._check_decode(O,H,name,w,i); T._check_coop_decode(H,name,w,i); n+=2
    except AssertionError as e:
        bad+=1; print('MISMATCH seed', s0, str(e)[:600]); 
        if bad>3: break
print('cases', n, 'rounds', rnd, 'mismatches', bad, 'secs', round(time.time()-t0))
