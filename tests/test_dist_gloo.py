"""world_size-2 gloo run of the multi-GPU plumbing (sharding + length / stream gathers) on CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spiht_b200 import dist as sdist


def test_shard_range_covers_batch():
    for n in (1, 7, 8, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [sdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_by_cost_balances_mixed_sizes():
    rng = np.random.default_rng(0)
    sizes = rng.choice([512, 1024, 2048, 4096], 64)
    costs = [float(s * s) for s in sizes]
    parts = sdist.shard_by_cost(costs, 8)
    assert sorted(i for p in parts for i in p) == list(range(64))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) <= 1.25 * (sum(costs) / 8)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B = 3
        rng = np.random.default_rng(100 + rank)
        nbits = torch.tensor(rng.integers(1, 600, B), dtype=torch.int64)
        max_n = torch.tensor(rng.integers(0, 14, B), dtype=torch.int32)
        streams = torch.from_numpy(rng.integers(0, 256, (B, 80), dtype=np.uint8))
        all_bits, all_n = sdist.gather_lengths(nbits, max_n)
        got = sdist.gather_streams(streams, nbits, dst=0)
        q.put((rank, nbits.tolist(), max_n.tolist(), streams.numpy().tolist(), all_bits.tolist(), all_n.tolist(),
               None if got is None else [list(b) for b in got]))
    finally:
        dist.destroy_process_group()


def test_gather_lengths_and_streams_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_bits = res[0][1] + res[1][1]
    want_n = res[0][2] + res[1][2]
    for r in res:
        assert r[4] == want_bits and r[5] == want_n
    assert res[1][6] is None
    rows = res[0][3] + res[1][3]
    for i, b in enumerate(res[0][6]):
        nb = (want_bits[i] + 7) // 8
        assert b == rows[i][:nb]
