"""Single-GPU stand-in for one rank of the N-GPU step: sort G producers' data, keep the partitions one owner would
receive, merge them (batched, in place) and print the library's phase timings.  Used under ncu for the launch list.
usage: python tools/merge_profile.py [records_per_producer] [producers] [partitions]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import tez_b200 as T
from tez_b200 import shuffle, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 2
P = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
dev = torch.device("cuda", 0)
p0, p1 = shuffle.owner_ranges(P, G)[0]
sorter = T.GpuSorter(P, fixed=(16, 64))
cap = n * 82 + 10 * P + 4096
d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
pieces, seglist, parts = [], [], []
for g in range(G):
    d_kv = synth.gen_c2(g * n, n, seed=4, device=dev)
    torch.cuda.synchronize()   # the library runs on its own stream
    out_len, index, st = sorter.sort_device_fixed(d_kv.data_ptr(), n, d_out.data_ptr(), cap)
    a, b = int(index[p0, 0]), int(index[p1 - 1, 0] + index[p1 - 1, 2])
    piece = d_out[a:b].clone()
    torch.cuda.synchronize()
    pieces.append(piece)
    for p in range(p0, p1):
        if index[p, 2]:
            seglist.append((piece.data_ptr() + int(index[p, 0]) - a, int(index[p, 2])))
            parts.append(p - p0)
    del d_kv
sorter.close()
del d_out
total = sum(x.numel() for x in pieces)
d_merged = torch.empty(int(total * 1.05) + (1 << 20), dtype=torch.uint8, device=dev)
m = T.GpuMerger(seglist, comparator=T.CMP_BYTES, device_ptrs=True, fixed=(16, 64), partitions=parts, num_partitions=p1 - p0)
for it in range(6):
    # iterations 3..5: the runs count as verified by the transport (tezgpu_fetch_segments_verified): no checksum pass
    verified = [True] * len(seglist) if it >= 3 else None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.reopen(seglist, parts, verified=verified)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    mlen, mindex, mst = m.write_partitions_device(d_merged.data_ptr(), d_merged.numel())
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("iter %d (%s): open %.2f ms  write %.2f ms  records %d  stats %s" % (
        it, "verified in fetch" if verified else "checksums in merge", (t1 - t0) * 1e3, (t2 - t1) * 1e3, m.counts()[0],
        {k: round(v, 3) if isinstance(v, float) else v for k, v in mst.items() if k.startswith("ms_") or k == "kernel_launches"}))
