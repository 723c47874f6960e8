"""Multi-GPU plan of the tools/call path: items are independent, so a job shards by
batch index - rank r of W owns the contiguous block [r * per_rank, (r + 1) * per_rank) - and there is
no collective on the data path.  torch.distributed is used for the launch barrier and for the
max-over-ranks reduction of the timings only (NCCL on the GPUs, gloo in the CPU tests)."""


def shard_range(per_rank, rank, world):
    """(first_item, n_items) of `rank` in a weak-scaling job with `per_rank` items per rank."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return rank * per_rank, per_rank


def split_batch(n_total, world):
    """Contiguous blocks of a fixed batch (strong scaling): [(first, count)] per rank, sizes differ by <= 1."""
    base, extra = divmod(n_total, world)
    out, first = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        out.append((first, cnt))
        first += cnt
    return out


def max_over_ranks(dist, value, device="cpu"):
    """max of a python float over all ranks (identity when dist is None / world is 1)"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class DeviceSet:
    """One process, several GPUs ("one pinned arena + stream set per GPU, results gathered by index").

    The Go gateway is a single process; this is the host-side shape it takes over the C ABI: one engine per
    device, one batching thread per engine bound to the GPU's NUMA node (ggr_bind_thread_to_device), the batch
    cut into contiguous blocks by index, every block's buffers in that GPU's own page-locked arena
    (ggr_host_alloc), the results concatenated by index.  No collective: the blocks never meet on a device."""

    def __init__(self, fds_bytes, devices, wire_order=0):
        from .engine import Engine
        self.engines = [Engine(d, wire_order) for d in devices]
        self.schemas = [e.register(fds_bytes) for e in self.engines]

    def message(self, full_name):
        return [s.message(full_name) for s in self.schemas]

    def _run(self, fn_name, msg_names, data, off, flags=0):
        """msg_names: per item full message name (ids differ per engine only in principle; they are looked up per engine)"""
        import threading
        import numpy as np
        n = len(off) - 1
        blocks = split_batch(n, len(self.engines))
        results = [None] * len(self.engines)
        errors = []

        def work(k):
            try:
                eng, sch = self.engines[k], self.schemas[k]
                eng.bind_thread()
                first, cnt = blocks[k]
                if cnt == 0:
                    results[k] = (np.zeros(0, np.uint8), np.zeros(1, np.uint64), np.zeros(0, np.int32))
                    return
                lo, hi = int(off[first]), int(off[first + cnt])
                ids = np.array([sch.message(m) for m in msg_names[first:first + cnt]], np.int32)
                h_data = eng.host_copy(np.ascontiguousarray(data[lo:hi]))
                h_off = eng.host_copy((off[first:first + cnt + 1] - off[first]).astype(np.uint64))
                results[k] = getattr(eng, fn_name)(sch, ids, h_data, h_off, flags)
            except Exception as ex:  # surfaced by the caller
                errors.append(ex)

        threads = [threading.Thread(target=work, args=(k,)) for k in range(len(self.engines))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        outs, offs, sts, base = [], [np.zeros(1, np.uint64)], [], 0
        for out, o, st in results:
            outs.append(out)
            offs.append(o[1:] + np.uint64(base))
            sts.append(st)
            base += int(o[-1])
        return np.concatenate(outs), np.concatenate(offs), np.concatenate(sts)

    def encode_batch(self, msg_names, data, off, flags=0):
        return self._run("encode_batch", msg_names, data, off, flags)

    def decode_batch(self, msg_names, data, off, flags=0):
        return self._run("decode_batch", msg_names, data, off, flags)

    def close(self):
        for s in self.schemas:
            s.release()
        for e in self.engines:
            e.close()
