"""Multi-GPU placement of one sample: one process per GPU, node range sharded (pm_index_create_shard), reads sharded for
seeding.  The exchanges are tiny relative to the streams each rank reads: (hash,count) tables, prefix-maximum records and
tie lists, moved with torch.distributed all_gather (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
import numpy as np


def all_gather_var(arr, device=None, group=None):
    """variable-length all_gather of a 1-D numpy array -> list of per-rank arrays (same dtype)"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = device if device is not None else torch.device("cpu")
    arr = np.ascontiguousarray(arr)
    n = torch.tensor([arr.size], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    raw = np.zeros(mx * arr.itemsize, np.uint8)
    raw[:arr.size * arr.itemsize] = arr.view(np.uint8).reshape(-1)
    t = torch.from_numpy(raw).to(dev)
    outs = [torch.empty(mx * arr.itemsize, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return [o.cpu().numpy()[:s * arr.itemsize].view(arr.dtype).copy() for o, s in zip(outs, sizes)]


def merge_records(per_rank_records):
    """per_rank_records: list over ranks of 5 x (rank[], node[], score[]) -> 5 x concatenated tuple"""
    out = []
    for m in range(5):
        out.append(tuple(np.concatenate([r[m][j] for r in per_rank_records]) for j in range(3)))
    return out


def place_sharded(ws, reads, offsets, total_reads, params, device=None, group=None):
    """reads/offsets: this rank's slice of the sample; ws: workspace over this rank's shard of the index.
    Returns the same Placement on every rank (== the single-GPU result)."""
    from .api import Placement, METRICS
    ws.stage_seed(reads, offsets, params)                                # A
    h, c = ws.stage_table_export()                                       # B
    hs = all_gather_var(h, device, group)
    cs = all_gather_var(c, device, group)
    ws.stage_table_import(np.concatenate(hs), np.concatenate(cs))
    ws.stage_score(params)                                               # C
    recs = ws.stage_records()                                            # D
    flat = []
    for m in range(5):
        for j in range(3):
            flat.append(all_gather_var(recs[m][j], device, group))
    world = len(flat[0])
    per_rank = [[tuple(flat[m * 3 + j][r] for j in range(3)) for m in range(5)] for r in range(world)]
    res = ws.stage_select(merge_records(per_rank), total_reads)          # E
    tied = {}
    for m, name in enumerate(METRICS):
        parts = all_gather_var(np.ascontiguousarray(res.tied[name], dtype=np.uint32), device, group)
        t = np.unique(np.concatenate(parts)) if parts else np.zeros(0, np.uint32)
        tied[name] = t.astype(np.uint32)
        res.tied[name] = tied[name]
        if t.size:
            res.best_index[name] = int(t[0])
    return res
