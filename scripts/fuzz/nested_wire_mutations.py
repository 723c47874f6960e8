import sys, time, random
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cases, hostsim, orc, wiremut, pbgen
import test_hostsim as T
fds = open(os.path.join(ROOT, 'tests', 'golden', 'schemas.binpb'), 'rb').read()
O = orc.Schema(fds); H = hostsim.Schema(fds)
MUTS=(wiremut.shuffle, wiremut.duplicate_some, wiremut.inject_unknown, wiremut.truncate, wiremut.corrupt)
def deep_mutate(w, rng, depth=0):
    try: fields = wiremut.split_fields(w)
    except Exception: return rng.choice(MUTS)(w, rng) if w else w
    subs=[k for k,(num,wt,raw) in enumerate(fields) if wt==2 and len(raw)>3]
    if subs and depth<3 and rng.random()<0.7:
        k=rng.choice(subs); num,wt,raw=fields[k]
        tag,i=wiremut.read_varint(raw,0); ln,j=wiremut.read_varint(raw,i); payload=raw[j:j+ln]
        newp=deep_mutate(payload, rng, depth+1)
        fields[k]=(num,wt,raw[:i]+wiremut.put_varint(len(newp))+newp)
        out=b"".join(r for _,_,r in fields)
        if rng.random()<0.4: out=rng.choice(MUTS[:3])(out, rng)
        return out
    return rng.choice(MUTS[:3] if rng.random()<0.8 else MUTS)(w, rng)
t0=time.time(); bad=0; n=0
budget=float(sys.argv[1]) if len(sys.argv)>1 else 300
names=[cases.A, cases.P+"CreateDocumentRequest", cases.P+"StructuredMetadata", cases.P+"Node", cases.P+"GetUserProfileResponse", cases.P+"ProcessNodeResponse", cases.WK]
seed=77000000
while time.time()-t0<budget and bad<5:
    for name in names:
        seed+=1; rng=random.Random(seed)
        w=pbgen.wire(pbgen.random_message(name, seed))
        for rep in range(4):
            try: b=deep_mutate(w, rng)
            except Exception: continue
            try:
                T._check_decode(O,H,name,b,seed%16,rep&1); T._check_coop_decode(H,name,b,seed%16,rep&1); n+=1
            except AssertionError as e:
                bad+=1; print('MISMATCH', seed, str(e)[:700])
print('cases', n, 'mismatches', bad, 'secs', round(time.time()-t0))
