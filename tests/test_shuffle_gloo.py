"""CPU, world_size 2 over gloo: the partition exchange (all-gather of the spill index + variable-size all-to-all)
delivers every producer's segment of every owned partition, and merging them reproduces a single global sort.
Compute in this test is the oracle (no GPU here); the exchange code is the product's tez_b200/shuffle.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, P, n, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import tez_oracle as O
    from tez_b200 import shuffle
    kv = O.gen_c2(rank * n, n, seed=4)
    r = O.pipelined_sort_fixed(O.sorter_conf(P), kv, 16, 64)
    file_out = torch.from_numpy(np.frombuffer(r["file_out"], dtype=np.uint8).copy())
    recv, segs = shuffle.exchange_partitions(file_out, r["index"], P)
    raw = recv.numpy().tobytes()
    p0, p1 = shuffle.owner_ranges(P, world)[rank]
    merged = {}
    for p in range(p0, p1):
        mine = [raw[o:o + ln] for o, ln, lp, g in segs if lp == p - p0]
        merged[p] = O.merge(mine, O.CMP_BYTES, factor=100)["ifile"] if mine else b""
    ret[rank] = merged
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("P", [8, 5])
def test_partition_exchange_world2(P):
    from oracle import tez_oracle as O
    from tez_b200 import shuffle
    world, n = 2, 3000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), P, n, ret), nprocs=world, join=True)
    # single global sort of all records with the same partitioner: every partition must match the merged result
    kv = np.concatenate([O.gen_c2(r * n, n, seed=4) for r in range(world)])
    g = O.pipelined_sort_fixed(O.sorter_conf(P), kv, 16, 64)
    got = {}
    for r in range(world):
        got.update(ret[r])
    assert sorted(got) == list(range(P))
    for p in range(P):
        start, raw, part = (int(x) for x in g["index"][p])
        assert got[p] == g["file_out"][start:start + part]
    assert shuffle.owner_ranges(8, 2) == [(0, 4), (4, 8)]
    assert shuffle.owner_ranges(5, 2) == [(0, 3), (3, 5)]
    assert shuffle.owner_ranges(1024, 8)[3] == (384, 512)
