"""CPU test of the N>1 path: world_size-2 gloo processes shard a batch, 'solve' their shard (the
per-instance solve is replaced by a deterministic stand-in -- no GPU here), all-gather the 32-byte
records and take the argmin; every rank must hold the full record set in instance order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from car_racing_b200 import _capi, sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_records(lo, hi):
    rec = np.zeros(hi - lo, dtype=_capi.RECORD_DTYPE)
    idx = np.arange(lo, hi)
    rec["cost"] = 100.0 + ((idx * 7919) % 101)        # minimum value 100 at idx % 101 == 0 ... ties across shards
    rec["u0"][:, 0] = idx
    rec["u0"][:, 1] = -idx
    rec["status"] = (idx % 5 == 0).astype(np.int32)
    rec["iters"] = idx % 37
    return rec


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(B, rank, world)
    local = sharding.records_to_tensor(_fake_records(lo, hi))
    allrec = sharding.all_gather_records(local, B)
    rec = sharding.tensor_to_records(allrec)
    q.put((rank, lo, hi, rec.tobytes(), sharding.argmin_first(allrec, 0)))
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [64, 101, 7])
def test_two_rank_gather_and_argmin(B):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    [p.start() for p in ps]
    got = [q.get(timeout=120) for _ in range(world)]
    [p.join(30) for p in ps]
    full = _fake_records(0, B)
    ok = full["status"] <= 0
    expect = int(np.argmin(np.where(ok, full["cost"], np.inf)))
    ranges = sorted((g[1], g[2]) for g in got)
    assert ranges[0][0] == 0 and ranges[-1][1] == B and ranges[0][1] == ranges[1][0]
    for rank, lo, hi, raw, am in got:
        assert raw == full.tobytes()
        assert am == expect


def test_shard_range_partition():
    for B in (1, 7, 1024, 4097):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(B, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
