"""world_size-2 CPU tests (gloo) of the multi-GPU plan: shard by batch index, no collective on the
data path, max-over-ranks timing.  The per-rank work is done by the CPU oracle here; on the GPUs the
same shards go through the engine (bench.py)."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import benchgen
from ggrmcp_b200 import shard

PER_RANK = 96


def _digest(S, wl):
    out, off, st = S.encode_batch(wl.req_msg, wl.req_json, wl.req_off, threads=2)
    assert int((np.asarray(st) != 0).sum()) == 0
    out2, off2, st2 = S.decode_batch(wl.rep_msg, wl.rep_wire, wl.rep_off, threads=2, cap=int(len(wl.rep_wire) * 3 + 64 * wl.n + 4096))
    assert int((np.asarray(st2) != 0).sum()) == 0
    return hashlib.sha256(bytes(out) + bytes(out2)).hexdigest()


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import orc
    with open(os.path.join(os.path.dirname(__file__), "golden", "schemas.binpb"), "rb") as fh:
        S = orc.Schema(fh.read())
    first, n = shard.shard_range(PER_RANK, rank, world)
    wl = benchgen.nested(n, S.msg, first=first)
    digest = _digest(S, wl)
    # the only communication: barrier + max over ranks of the (fake) step time
    dist.barrier()
    t = shard.max_over_ranks(dist, 1.0 + rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, (rank, first, n, digest))
    if rank == 0:
        ret["t"] = t
        ret["shards"] = gathered
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_split_batch():
    assert shard.split_batch(10, 4) == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert shard.split_batch(2, 4) == [(0, 1), (1, 1), (2, 0), (2, 0)]
    with pytest.raises(ValueError):
        shard.shard_range(5, 2, 2)


def test_two_ranks_shard_by_index(oracle):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret["t"] == 2.0  # max over ranks
    shards = sorted(ret["shards"])
    assert [(r, f, n) for r, f, n, _ in shards] == [(0, 0, PER_RANK), (1, PER_RANK, PER_RANK)]
    # every shard is exactly the corresponding block of the unsharded job
    for r, first, n, digest in shards:
        wl = benchgen.nested(n, oracle.msg, first=first)
        assert _digest(oracle, wl) == digest
    assert shards[0][3] != shards[1][3]
