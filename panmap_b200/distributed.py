"""Multi-GPU placement of one sample: one process per GPU, node range sharded (pm_index_create_shard), reads sharded for
seeding.  The exchanges are tiny relative to the streams each rank reads: (hash,count) tables, prefix-maximum records and
tie lists, moved with torch.distributed all_gather (NCCL over NVLink on the GPU box, gloo in the CPU tests).

`ws` only needs the stage_* methods of panmap_b200.api.Workspace, so the protocol is testable on CPU with a stand-in."""
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
    return [tuple(np.concatenate([r[m][j] for r in per_rank_records]) for j in range(3)) for m in range(5)]


def exchange_tables_device(ws, device, group=None):
    """all-gather of the per-rank (hash,count) tables entirely on the device (NCCL): no host staging of the table"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n_local = ws.stage_table_export_dev(None, None, 0)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n_local], dtype=torch.int64, device=device), group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    buf = torch.empty((2, mx), dtype=torch.int64, device=device)           # row 0: hashes (bit pattern), row 1: counts
    buf[1].zero_()                                                         # padding entries have count 0 and are skipped on import
    torch.cuda.current_stream(device).synchronize()                         # the library runs on its own stream
    ws.stage_table_export_dev(buf[0].data_ptr(), buf[1].data_ptr(), mx)
    allbuf = torch.empty((world, 2, mx), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allbuf, buf, group=group)
    h = allbuf[:, 0, :].contiguous().view(-1)
    c = allbuf[:, 1, :].contiguous().view(-1)
    torch.cuda.current_stream(device).synchronize()
    ws.stage_table_import_dev(h.data_ptr(), c.data_ptr(), h.numel())
    return sizes


def place_sharded(ws, reads, offsets, total_reads, params, device=None, group=None, resident=False):
    """reads/offsets: this rank's slice of the sample; ws: workspace over this rank's shard of the index.
    Returns the same Placement on every rank (== the single-GPU result)."""
    from .api import METRICS
    on_gpu = device is not None and getattr(device, "type", "cpu") == "cuda" and hasattr(ws, "stage_table_export_dev")
    if resident:
        ws.stage_seed_resident(params)                                   # A (reads uploaded earlier with ws.upload)
    else:
        ws.stage_seed(reads, offsets, params)                            # A
    if on_gpu:
        exchange_tables_device(ws, device, group)                        # B (NCCL, device buffers)
    else:
        h, c = ws.stage_table_export()                                   # B (host buffers: gloo / tests)
        hs = all_gather_var(h, device, group)
        cs = all_gather_var(c, device, group)
        ws.stage_table_import(np.concatenate(hs), np.concatenate(cs))
    ws.stage_score(params)                                               # C
    recs = ws.stage_records()                                            # D
    # one gather for all record arrays: [counts(5) | ranks | nodes | scores-as-u64] packed into a single uint64 vector
    counts = np.array([len(r[0]) for r in recs], np.uint64)
    blob = np.concatenate([counts] + [np.concatenate([r[0].astype(np.uint64), r[1].astype(np.uint64), r[2].view(np.uint64)]) for r in recs])
    blobs = all_gather_var(blob, device, group)
    per_rank = []
    for b in blobs:
        cnt = b[:5].astype(np.int64)
        p = 5
        rr = []
        for m in range(5):
            n = int(cnt[m])
            rr.append((b[p:p + n].astype(np.uint32), b[p + n:p + 2 * n].astype(np.uint32), b[p + 2 * n:p + 3 * n].copy().view(np.float64)))
            p += 3 * n
        per_rank.append(rr)
    res = ws.stage_select(merge_records(per_rank), total_reads)          # E
    tcounts = np.array([len(res.tied[name]) for name in METRICS], np.uint64)
    tblob = np.concatenate([tcounts] + [np.ascontiguousarray(res.tied[name], dtype=np.uint64) for name in METRICS])
    tblobs = all_gather_var(tblob, device, group)
    for m, name in enumerate(METRICS):
        parts = []
        for b in tblobs:
            cnt = b[:5].astype(np.int64)
            p = 5 + int(cnt[:m].sum())
            parts.append(b[p:p + int(cnt[m])].astype(np.uint32))
        t = np.unique(np.concatenate(parts)) if parts else np.zeros(0, np.uint32)
        res.tied[name] = t.astype(np.uint32)
        if t.size:
            res.best_index[name] = int(t[0])
    return res
