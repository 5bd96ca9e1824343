"""Worker of tests/test_peer_fetch_gpu.py: one of G processes sharing cuda:0 (CUDA IPC works between processes on
the same device), rendezvous over gloo.  sort -> PeerExchange (pull from the other processes' exported buffers)
-> batched device merge, checked partition by partition against the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tez_oracle as O  # noqa: E402
import tez_b200 as T  # noqa: E402
from tez_b200 import shuffle  # noqa: E402


def main():
    n, P, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    cap = n * 82 + 10 * P + 4096
    px = shuffle.PeerExchange(cap, 0)
    sorter = T.GpuSorter(P, fixed=(16, 64))
    p0, p1 = shuffle.owner_ranges(P, world)[rank]
    merger = None
    for k in range(steps):
        kvs = [O.gen_c2(g * n + 1000 * k, n, seed=11) for g in range(world)]
        d_kv = torch.from_numpy(kvs[rank]).to(dev)
        out_len, index, _ = sorter.sort_device_fixed(d_kv.data_ptr(), n, px.out_ptr(k), cap)
        segs = px.exchange(k, index, P)
        seg_list = [(ptr, ln) for ptr, ln, _, _ in segs]
        parts = [p for _, _, p, _ in segs]
        if merger is None:
            merger = T.GpuMerger(seg_list, comparator=T.CMP_BYTES, device_ptrs=True, fixed=(16, 64), partitions=parts,
                                 num_partitions=max(1, p1 - p0), verified=px.last_verified)
        else:
            merger.reopen(seg_list, parts, verified=px.last_verified)
        assert px.last_verified is not None and sum(px.last_verified) == sum(1 for _, _, _, g in segs if g != rank)
        d_merged = torch.empty(merger.output_bound() + 64, dtype=torch.uint8, device=dev)
        mlen, mindex, _ = merger.write_partitions_device(d_merged.data_ptr(), d_merged.numel())
        got = d_merged[:mlen].cpu().numpy().tobytes()
        # oracle: every producer's file.out, then TezMerger over the G runs of each owned partition
        outs = [O.pipelined_sort_fixed(O.sorter_conf(P), kv, 16, 64) for kv in kvs]
        for p in range(p0, p1):
            runs = []
            for g in range(world):
                a, _, ln = (int(x) for x in outs[g]["index"][p])
                if ln:
                    runs.append(outs[g]["file_out"][a:a + ln])
            a, raw, ln = (int(x) for x in mindex[p - p0])
            if not runs:
                assert ln == 0 or got[a:a + ln] == O.write_ifile([])[0], "empty partition %d" % p
                continue
            exp = O.merge(runs, O.CMP_BYTES, factor=100)["ifile"]
            assert got[a:a + ln] == exp, "rank %d step %d partition %d differs from the oracle merge" % (rank, k, p)
    dist.barrier()
    merger.close()
    px.close()
    dist.destroy_process_group()
    print("peer worker %d ok" % rank)


if __name__ == "__main__":
    main()
