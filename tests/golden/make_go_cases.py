#!/usr/bin/env python3
"""Writes tests/golden/go/cases.jsonl: the inputs oracle/go_ref replays through the real reference path.

Run `python tests/golden/make_go_cases.py`, then (on a box with Go 1.23) `cd oracle/go_ref && go run .`;
tests/test_go_golden.py compares the oracle with the resulting golden.jsonl whenever that file exists."""
import base64
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def main():
    import cases
    import benchgen
    from test_oracle import K_BODIES
    out = []

    def args(tag, name, js):
        try:
            out.append({"id": "%s-%d" % (tag, len(out)), "kind": "args", "message": name, "args": js.decode("utf-8")})
        except UnicodeDecodeError:
            out.append({"id": "%s-%d" % (tag, len(out)), "kind": "args", "message": name, "args": base64.b64encode(js).decode(), "args_b64": True})

    def reply(tag, name, w, rid):
        out.append({"id": "%s-%d" % (tag, len(out)), "kind": "reply", "message": name, "wire_hex": w.hex(), "req_id": rid})

    rng = random.Random(2024)
    for name, js, _ in cases.K_REQUESTS:
        args("kreq", name, js)
    for name, js, _ in cases.ENCODE_EDGE:
        args("edge", name, js)
    for name, js in cases.random_encode_cases(120, seed0=61000):
        args("rand", name, js)
        args("mut", name, cases.mutate_json(js, rng))
    ids = ["1", '"abc"', "9007199254740993", '"\\u00e9<"', "-7", "1.5"]
    for name, w, _ in cases.K_REPLIES:
        reply("krep", name, bytes.fromhex(w), "1")
    for i, (name, h) in enumerate(cases.DECODE_EDGE_HEX):
        reply("dedge", name, bytes.fromhex(h), ids[i % len(ids)])
    for i, (name, w) in enumerate(cases.random_decode_cases(100, seed0=62000)):
        reply("drand", name, w, ids[i % len(ids)])
    for i, (name, w) in enumerate(cases.merge_cases()):
        reply("merge", name, w, ids[i % len(ids)])
    for name, js in cases.WKT_ENCODE:
        args("wkt", name, js)
    for i, (name, w) in enumerate(cases.wkt_decode_cases()):
        reply("wktrep", name, w, ids[i % len(ids)])
    names = {}

    def mi(n):
        names.setdefault(n, len(names))
        return names[n]

    inv = lambda: {v: k for k, v in names.items()}
    for gen, n in ((benchgen.nested, 40), (benchgen.flat, 40), (benchgen.mixed, 300)):
        wl = gen(n, mi)
        jb, wb = wl.req_json.tobytes(), wl.rep_wire.tobytes()
        nm = inv()
        for i in range(n):
            js = jb[int(wl.req_off[i]):int(wl.req_off[i + 1])]
            w = wb[int(wl.rep_off[i]):int(wl.rep_off[i + 1])]
            if len(js) < 6000:
                args(wl.name, nm[int(wl.req_msg[i])], js)
            if len(w) < 6000:
                reply(wl.name, nm[int(wl.rep_msg[i])], w, str(i))
    for b in K_BODIES:
        body = b[0]
        out.append({"id": "body-%d" % len(out), "kind": "body", "body": base64.b64encode(body).decode()})
        for v in (body.replace(b'"jsonrpc"', b'"JSONRPC"'), body.replace(b'"id":', b'"id":1,"id":'), body[:-1] + b',"extra":[1,2]}'):
            out.append({"id": "body-%d" % len(out), "kind": "body", "body": base64.b64encode(v).decode()})
    os.makedirs(os.path.join(HERE, "go"), exist_ok=True)
    path = os.path.join(HERE, "go", "cases.jsonl")
    with open(path, "w") as fh:
        for c in out:
            fh.write(json.dumps(c, ensure_ascii=True) + "\n")
    print("wrote %s: %d cases, %d bytes" % (path, len(out), os.path.getsize(path)))


if __name__ == "__main__":
    main()
