"""Scatter-gather shuffle between the GPUs of one box -- what ScatterGatherEdgeManager + ShuffleHandler +
FetcherOrderedGrouped do over HTTP in the reference (DAG/dag/app/dag/impl/ScatterGatherEdgeManager.java:108-131,
OG/FetcherOrderedGrouped.java:437-632): consumer p pulls partition p from every producer.

Here every rank is one producer and owns a contiguous block of partitions (owner(p) = p * G // P).  Because a
producer's file.out already stores its partitions in partition order, the bytes destined to one owner are ONE
contiguous range of file.out: the exchange is a single variable-size all-to-all (NCCL send/recv pairs over NVLink)
with no repacking, preceded by an all-gather of the TezSpillRecord index (P x 3 int64 per rank).

Pure torch.distributed plumbing: works on CUDA tensors over NCCL and on CPU tensors over gloo (tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def owner_ranges(num_partitions, world):
    """[p0, p1) of every rank for owner(p) = p * world // num_partitions (contiguous blocks)."""
    bounds = [(num_partitions * g + world - 1) // world for g in range(world + 1)]
    return [(bounds[g], bounds[g + 1]) for g in range(world)]


def exchange_partitions(file_out, index, num_partitions, group=None):
    """file_out: uint8 tensor with this rank's file.out bytes (device or CPU); index: [P,3] int64 numpy
    (start, rawLength, partLength).  Returns (recv uint8 tensor, segments) where segments is a list of
    (offset_in_recv, length, local_partition, source_rank) for every non-empty segment this rank now owns."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = file_out.device
    idx = torch.from_numpy(np.ascontiguousarray(index, dtype=np.int64)).to(dev)
    all_idx = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(all_idx, idx, group=group)
    all_idx = torch.stack(all_idx).cpu().numpy()  # [world, P, 3]
    ranges = owner_ranges(num_partitions, world)
    part_len = all_idx[:, :, 2]
    # what I send to g: my segments of g's partitions (contiguous in my file.out)
    send_sizes = [int(part_len[rank, a:b].sum()) for a, b in ranges]
    p0, p1 = ranges[rank]
    recv_sizes = [int(part_len[g, p0:p1].sum()) for g in range(world)]
    send_off0 = int(all_idx[rank, ranges[0][0], 0]) if num_partitions else 0
    send = file_out[send_off0:send_off0 + sum(send_sizes)]
    recv = torch.empty(max(1, sum(recv_sizes)), dtype=torch.uint8, device=dev)
    dist.all_to_all_single(recv[:sum(recv_sizes)], send, output_split_sizes=recv_sizes, input_split_sizes=send_sizes,
                           group=group)
    segments = []
    off = 0
    for g in range(world):
        for p in range(p0, p1):
            ln = int(part_len[g, p])
            if ln:
                segments.append((off, ln, p - p0, g))
            off += ln
    return recv, segments
