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


def test_pull_plan_and_segment_table_resolve_to_the_right_bytes():
    """The NVLink pull shuffle's planning (pure numpy): emulate device memory with fake address spaces and check that
    every segment address resolves to the producer's bytes, in (source rank, partition) order, for every rank."""
    import numpy as np
    from tez_b200 import shuffle
    rng = np.random.default_rng(7)
    world, P = 4, 13
    files, indexes = [], []
    for g in range(world):
        lens = rng.integers(0, 200, size=P)
        lens[rng.integers(0, P)] = 0                     # an empty partition
        idx = np.zeros((P, 3), dtype=np.int64)
        idx[:, 0] = np.concatenate([[0], np.cumsum(lens)[:-1]]) + 3   # file.out ranges (arbitrary start)
        idx[:, 1] = np.maximum(lens - 4, 0)
        idx[:, 2] = lens
        files.append(rng.integers(0, 256, size=int(lens.sum()) + 3, dtype=np.uint8))
        indexes.append(idx)
    all_idx = np.stack(indexes)
    peer_ptrs = [(g + 1) << 40 | (5 * g + 1) for g in range(world)]     # deliberately unaligned file.out addresses
    recv_base = 1 << 50
    for rank in range(world):
        ranges, seg_src, need = shuffle.pull_plan(all_idx, rank, P, peer_ptrs)
        assert [g for g, _, _, _ in ranges] == [(rank + k) % world for k in range(1, world) if any(r[0] == (rank + k) % world for r in ranges)]
        recv = np.zeros(need + 16, dtype=np.uint8)
        for g, src, off, ln in ranges:
            assert (src - (recv_base + off)) % 16 == 0           # 128-bit moves possible
            a = src - peer_ptrs[g]
            recv[off:off + ln] = files[g][a:a + ln]              # what the fetch kernel does
        segs = shuffle.pull_segments(all_idx, rank, P, peer_ptrs, seg_src, recv_base)
        p0, p1 = shuffle.owner_ranges(P, world)[rank]
        want = [(g, p) for g in range(world) for p in range(p0, p1) if indexes[g][p, 2]]
        assert [(g, p0 + lp) for _, _, lp, g in segs] == want
        for ptr, ln, lp, g in segs:
            s0, l0 = int(indexes[g][p0 + lp, 0]), int(indexes[g][p0 + lp, 2])
            assert ln == l0
            if g == rank:
                got = files[g][ptr - peer_ptrs[g]: ptr - peer_ptrs[g] + ln]
            else:
                got = recv[ptr - recv_base: ptr - recv_base + ln]
            assert np.array_equal(got, files[g][s0:s0 + l0])


def _peer_unavailable_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tez_b200 import shuffle
    try:
        shuffle.PeerExchange(1 << 20, 0)
        ret[rank] = "constructed"
    except RuntimeError as e:
        ret[rank] = "RuntimeError: %s" % e
    # the group is still usable: nobody is stuck in a collective the others skipped
    t = torch.tensor([rank + 1])
    dist.all_reduce(t)
    ret["sum%d" % rank] = int(t.item())
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.is_available(), reason="exercises the no-device failure path")
def test_peer_exchange_fails_on_every_rank_or_none():
    """Without a device the exportable buffer cannot be allocated: every rank must raise (consensus) so that a caller
    can fall back to exchange_partitions() with the process group intact."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_peer_unavailable_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        assert ret[r].startswith("RuntimeError: peer pull unavailable"), ret[r]
        assert ret["sum%d" % r] == 3


def test_verified_pull_visits_producers_in_ring_order():
    """shuffle.ring_order: the rows of the segment table (ordered by source rank for the merge's tie-breaking) are pulled
    producer rank+1 first, so that at any moment a producer serves one consumer."""
    from tez_b200 import shuffle
    world, P = 8, 16
    for rank in range(world):
        segs = [(1000 * g + p, 10 + p, p, g) for g in range(world) for p in range(P // world * 0, 2)]
        order = shuffle.ring_order(segs, rank, world)
        assert [t[3] for t in order] == [g % world for g in range(rank + 1, rank + world) for _ in range(2)]
        assert all(t[3] != rank for t in order) and len(order) == 2 * (world - 1)
        for g in range(world):
            assert [t[2] for t in order if t[3] == g] == sorted(t[2] for t in order if t[3] == g)
