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


_EXCH_CAP = {}


def exchange_tables_device(ws, device, group=None):
    """all-gather of the per-rank (hash,count) tables entirely on the device (NCCL), one collective per sample.
    Every rank exports into a fixed-capacity [2, cap] int64 buffer (row 0 hashes, row 1 counts; unused entries have count 0
    and are skipped on import; entry [1, cap-1] carries the rank's true entry count so an undersized cap is detected)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    key = id(ws)
    cap = _EXCH_CAP.get(key)
    if cap is None:   # first sample: size the buffers from an exact count (all ranks take the same maximum)
        n_local = ws.stage_table_export_dev(None, None, 0)
        t = torch.tensor([n_local], dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        cap = int(t.item()) * 5 // 4 + 1024
    while True:
        bufs = getattr(ws, "_exch_bufs", None)
        if bufs is None or bufs[0] != cap:
            bufs = (cap, torch.empty((2, cap + 1), dtype=torch.int64, device=device), torch.empty((world, 2, cap + 1), dtype=torch.int64, device=device))
            ws._exch_bufs = bufs
            torch.cuda.current_stream(device).synchronize()
        _, buf, allbuf = bufs
        n_local = ws.stage_table_export_dev(buf[0].data_ptr(), buf[1].data_ptr(), cap)   # returns synchronised; count tail zeroed
        buf[0, cap] = n_local
        dist.all_gather_into_tensor(allbuf, buf, group=group)
        counts = allbuf[:, 0, cap].cpu()                                     # also orders the torch stream before the import
        need = int(counts.max())
        if need <= cap:
            break
        cap = need * 5 // 4 + 1024
    _EXCH_CAP[key] = cap
    total = int(counts.sum())
    for r in range(world):   # rank slices are contiguous in the gathered tensor: import them in place
        ws.stage_table_import_dev_async(allbuf[r, 0].data_ptr(), allbuf[r, 1].data_ptr(), cap, clear_first=(r == 0), expected_total=total)


def all_gather_fixed(blob, cap, device, group=None):
    """all_gather of a uint64 vector padded to `cap` entries (first entry = length); falls back to the ragged path"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if device is None or getattr(device, "type", "cpu") != "cuda":
        return all_gather_var(blob, device, group)          # CPU / gloo: ragged path (same decision on every rank)
    flag = blob.size + 1 <= cap
    t = torch.zeros(cap, dtype=torch.int64)
    if flag:
        t[0] = blob.size
        t[1:1 + blob.size] = torch.from_numpy(blob.view(np.int64))
    else:
        t[0] = -1
    t = t.to(device) if device is not None else t
    out = torch.empty((world, cap), dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    o = out.cpu().numpy()
    if (o[:, 0] < 0).any():    # some rank did not fit: everyone takes the ragged path
        return all_gather_var(blob, device, group)
    return [o[r, 1:1 + int(o[r, 0])].view(np.uint64).copy() for r in range(world)]


def place_sharded(ws, reads, offsets, total_reads, params, device=None, group=None, resident=False):
    """reads/offsets: this rank's slice of the sample; ws: workspace over this rank's shard of the index.
    Returns the same Placement on every rank (== the single-GPU result)."""
    from .api import METRICS
    import os, time
    if getattr(params, "dedup_reads", 0):
        raise ValueError("dedup_reads needs the whole sample on one GPU: duplicates across the per-rank read shards would go unseen")
    trace = os.environ.get("PM_TRACE")
    tt = [time.perf_counter()]
    def mark():
        if trace:
            tt.append(time.perf_counter())
    on_gpu = device is not None and getattr(device, "type", "cpu") == "cuda" and hasattr(ws, "stage_table_export_dev")
    if resident:
        ws.stage_seed_resident(params)                                   # A (reads uploaded earlier with ws.upload)
    else:
        ws.stage_seed(reads, offsets, params)                            # A
    mark()
    if on_gpu:
        exchange_tables_device(ws, device, group)                        # B (NCCL, device buffers)
    else:
        h, c = ws.stage_table_export()                                   # B (host buffers: gloo / tests)
        hs = all_gather_var(h, device, group)
        cs = all_gather_var(c, device, group)
        ws.stage_table_import(np.concatenate(hs), np.concatenate(cs))
    mark()
    ws.stage_score(params)                                               # C
    mark()
    recs = ws.stage_records_all() if hasattr(ws, "stage_records_all") else ws.stage_records()   # D
    # one gather for all record arrays: [counts(5) | ranks | nodes | scores-as-u64] packed into a single uint64 vector
    counts = np.array([len(r[0]) for r in recs], np.uint64)
    blob = np.concatenate([counts] + [np.concatenate([r[0].astype(np.uint64), r[1].astype(np.uint64), r[2].view(np.uint64)]) for r in recs])
    mark()
    blobs = all_gather_fixed(blob, 2048, device, group)
    mark()
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
    mark()
    tcounts = np.array([len(res.tied[name]) for name in METRICS], np.uint64)
    tblob = np.concatenate([tcounts] + [np.ascontiguousarray(res.tied[name], dtype=np.uint64) for name in METRICS])
    tblobs = all_gather_fixed(tblob, 2048, device, group)
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
    if trace:
        mark()
        import torch.distributed as dist
        if dist.get_rank() == 0:
            names = ["seed", "exchange", "score", "records", "gather_rec", "select", "ties"]
            print("PM_TRACE " + " ".join(f"{n}={1e3*(b-a):.2f}ms" for n, a, b in zip(names, tt[:-1], tt[1:])), flush=True)
    return res
