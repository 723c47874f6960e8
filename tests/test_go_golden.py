"""The oracle against the REAL reference path: consumes tests/golden/go/golden.jsonl, the output of
oracle/go_ref (protojson / dynamicpb / encoding/json of the pinned versions) on tests/golden/go/cases.jsonl.
The build image has no Go toolchain, so the golden file does not exist there and the byte-level comparison is
skipped with that reason; the structure of the case file and of the Go program is still checked."""
import base64
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GO_DIR = os.path.join(HERE, "golden", "go")
GOLDEN = os.path.join(GO_DIR, "golden.jsonl")


def _cases():
    with open(os.path.join(GO_DIR, "cases.jsonl")) as fh:
        return [json.loads(l) for l in fh]


def test_case_file_and_program_are_in_place():
    cs = _cases()
    kinds = {c["kind"] for c in cs}
    assert kinds == {"args", "reply", "body"} and len(cs) > 800
    src = open(os.path.join(os.path.dirname(HERE), "oracle", "go_ref", "main.go")).read()
    for call in ("protojson.Unmarshal", "protojson.Marshal", "proto.Unmarshal", "dynamicpb.NewMessage", "json.Marshal(args)",
                 "ValidateToolCallParams", "json.NewEncoder"):
        assert call in src, call
    mod = open(os.path.join(os.path.dirname(HERE), "oracle", "go_ref", "go.mod")).read()
    assert "google.golang.org/protobuf v1.36.6" in mod and "google.golang.org/grpc v1.74.2" in mod


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="tests/golden/go/golden.jsonl absent: no Go toolchain has run oracle/go_ref yet "
                                                          "(oracle-vs-Go byte parity stays unpinned until one does)")
def test_oracle_equals_go(oracle):
    cases = {c["id"]: c for c in _cases()}
    lines = [json.loads(l) for l in open(GOLDEN)]
    header = lines[0]
    assert header.get("header")
    flags_json = 1 if header["comma_space"] else 0
    checked = 0
    for g in lines[1:]:
        c = cases[g["id"]]
        if c["kind"] == "args":
            js = base64.b64decode(c["args"]) if c.get("args_b64") else c["args"].encode()
            if g.get("stage") == "decode":
                continue  # not JSON at all: encoding/json refuses before anything on this path runs
            canon = base64.b64decode(g.get("canon_args", ""))
            # the boundary of this repo takes the canonical text (handler.go:224-231's output)
            rc, wire, _ = oracle.encode(c["message"], canon, 2)  # ORC_F_GO_LEGACY_ORDER
            assert (rc == 0) == (g["stage"] == "ok"), (g["id"], rc, g.get("error"))
            if rc == 0:
                assert wire.hex() == g["wire_hex"], g["id"]
        elif c["kind"] == "reply":
            w = bytes.fromhex(c["wire_hex"])
            rc, text, _ = oracle.decode(c["message"], w, flags_json)
            assert (rc == 0) == (g["stage"] == "ok"), (g["id"], rc, g.get("error"))
            if rc == 0:
                assert text == base64.b64decode(g["json"]), g["id"]
                st, body = oracle.response(c["message"], w, c["req_id"].encode(), flags_json)
                assert st == 0 and body == base64.b64decode(g["http_body"]), g["id"]
        else:
            r = oracle.request(base64.b64decode(c["body"]))
            assert (r["kind"] == 0) == (g["stage"] == "ok"), (g["id"], r["kind"], g.get("error"))
            if r["kind"] == 0:
                rc, wire, _ = oracle.encode(oracle.methods()[r["method"]]["input_name"], base64.b64decode(g["canon_args"]), 2) \
                    if "input_name" in oracle.methods()[r["method"]] else (0, None, None)
                if wire is not None:
                    assert wire.hex() == g["wire_hex"], g["id"]
        checked += 1
    assert checked > 500
