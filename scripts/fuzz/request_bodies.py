import sys, time, random
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cases, hostsim, orc
import test_hostsim as T
fds = open(os.path.join(ROOT, 'tests', 'golden', 'schemas.binpb'), 'rb').read()
O = orc.Schema(fds); H = hostsim.Schema(fds)
by_input = T._tool_by_input(O)
t0=time.time(); bad=0; n=0; handled=0
budget=float(sys.argv[1]) if len(sys.argv)>1 else 300
s0=31000000
ids=[b"1", b"42", b'"abc"', b"-7", b"123456789012345", b"1.0", b"1e3", b'"a<b"', b'"\\u00e9"', b"9007199254740993", b"null", b"0", b'"x y"', b"1234567890123456", b'""', b"0.5", b"-0", b"1E2", b'"\\n"', b"true", b"[1]", b"{}"]
while time.time()-t0 < budget and bad < 4:
    rng = random.Random(s0); s0 += 1000
    for i,(name,js) in enumerate(cases.random_encode_cases(40, seed0=s0)):
        k = O.msg(name)
        if k not in by_input: continue
        tool = by_input[k][1].encode()
        for rep in range(3):
            a = js if rep == 0 else cases.mutate_json(js, rng)
            idt = rng.choice(ids)
            parts = {"jsonrpc": b'"jsonrpc":"2.0"', "id": b'"id":' + idt, "method": b'"method":"tools/call"',
                     "params": b'"params":{"name":"' + tool + b'","arguments":' + a + b"}"}
            keys = list(parts); rng.shuffle(keys)
            body = b"{" + b",".join(parts[k2] for k2 in keys) + b"}"
            if rep == 2: body = cases.mutate_json(body, rng)
            try:
                handled += T._check_request(O, H, body, i); n += 1
            except AssertionError as e:
                bad += 1; print('MISMATCH', s0, str(e)[:500])
print('bodies', n, 'handled', handled, 'mismatches', bad, 'secs', round(time.time()-t0))
