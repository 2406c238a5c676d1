"""One sample over N GPUs, one process per GPU: thin glue between a process group the caller already has (torchrun /
torch.distributed, any backend) and the library's own NCCL data plane (pm_comm, include/panmap_b200.h).

torch.distributed is used for exactly one thing here: carrying the 128-byte communicator id from rank 0 to the other ranks.
Everything per sample -- seeding of the rank's read slice, the hash-partitioned seed table (all-to-all), the finalized
(count, seed id) pairs, records and tie heads (all-gathers) -- is enqueued by libpanmap_b200.so on the workspace's CUDA stream
through NCCL; there is no Python on the data path."""
import numpy as np


def read_slice(offsets, rank, world):
    """contiguous slice of the sample's reads for `rank`: (first read, one past the last read)"""
    n = int(len(offsets) - 1)
    return (n * rank) // world, (n * (rank + 1)) // world


def slice_reads(reads, offsets, rank, world):
    """(bytes, offsets starting at 0) of this rank's slice"""
    lo, hi = read_slice(offsets, rank, world)
    off = np.ascontiguousarray(offsets[lo:hi + 1] - offsets[lo], dtype=np.uint64)
    return reads[int(offsets[lo]):int(offsets[hi])], off


def make_comm(workspace, group=None):
    """pm_comm over the ranks of `group` (default: the world); the workspace must sit on shard `rank` of `world` shards"""
    import torch.distributed as dist
    from .api import Comm, comm_unique_id
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return Comm.nccl(workspace, box[0], rank, world)
