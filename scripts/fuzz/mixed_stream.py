import sys, time
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import benchgen, hostsim, orc
import test_hostsim as T
fds = open(os.path.join(ROOT, 'tests', 'golden', 'schemas.binpb'), 'rb').read()
O = orc.Schema(fds); H = hostsim.Schema(fds)
names={}
def mi(n):
    names[H.msg(n)] = n; return H.msg(n)
t0=time.time(); bad=0; n=0
budget=float(sys.argv[1]) if len(sys.argv)>1 else 300
first=200000
while time.time()-t0 < budget and bad < 4:
    wl = benchgen.mixed(600, mi, first=first); first += 600
    jb, wb = wl.req_json.tobytes(), wl.rep_wire.tobytes()
    for i in range(wl.n):
        js = jb[int(wl.req_off[i]):int(wl.req_off[i+1])]; w = wb[int(wl.rep_off[i]):int(wl.rep_off[i+1])]
        if len(js) > 30000 or len(w) > 30000: continue
        rn, pn = names[int(wl.req_msg[i])], names[int(wl.rep_msg[i])]
        try:
            rc, ow, _ = O.encode(rn, js); st, ew = H.encode(rn, js, i % 16, (i*3) % 16)
            assert rc == 0 and st == 0 and ew == ow, ('enc', rn, js[:200])
            T._check_walk(H, rn, js, i)
            rc, oj, _ = O.decode(pn, w); st, ej = H.decode(pn, w, 0, i % 16, (i*7) % 16)
            assert rc == 0 and st == 0 and ej == oj, ('dec', pn, w.hex()[:200])
            rc2, out = H.decode_coop(pn, w, 0, i % 16, (i*5) % 16)
            assert rc2 in (0, 200), ('coop rc', pn, rc2)
            if rc2 == 0: assert out == oj, ('coop', pn, w.hex()[:200])
            n += 1
        except AssertionError as e:
            bad += 1; print('MISMATCH first', first, str(e)[:400])
print('items', n, 'mismatches', bad, 'secs', round(time.time()-t0))
