"""N > 1 path on CPU: a world_size-2 gloo run of the sharded-sample protocol of panmap_b200/csrc/pm_multi.cu, restated with numpy, the
oracle and the host emulation of the kernels (tests/hostcheck): read slices -> local seed tables -> hash partition by seedOwner
(all-to-all) -> per-partition statistics + (count, seed id) pairs, count-1 seeds outside the index as a bare count -> all-gather ->
min-support rule from the summed statistics, magnitudes, per-shard scoring -> records all-gather -> identical tolerance chain ->
tie heads all-gather.  Buffers have fixed capacities with in-band counts like the device ones.  The result must equal the
single-process oracle placement; panmap_b200.distributed's slicing / id hand-off helpers are exercised on the way."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from oracle import cpu
from tests import helpers as H

METRICS = ("log_raw", "log_cosine", "containment", "weighted_containment", "log_containment")
NONE = 0xFFFFFFFF


def _gather(t, world):
    import torch
    import torch.distributed as dist
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return outs


def emulate_sharded(S, reads, off, rank, world, min_read_support=-1, cap_pair=4096, cap_g=8192, rec_x=64, tie_head=4096):
    """one rank of the protocol; returns (best_score[5], best_index[5], tied[5], scalars)"""
    import torch
    import panmap_b200 as pm
    hc = C.CDLL(os.path.join(H.ROOT, "tests", "hostcheck", "libhostcheck.so"))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    host = pm.HostIndex(S.hash, S.parent, S.child, S.offsets, S.parent_index, S.k, S.s, S.t, S.l)
    # the index dictionary: dense ids in first-appearance order (pm_flatten.cpp); any fixed numbering works for the protocol
    dict_hash = np.unique(S.hash)
    # P0: local table of this rank's read slice, exported by owner
    h, c = cpu.seed_table(reads, off, S.k, S.s, S.t, S.l)
    owner = np.zeros(h.size, np.uint32)
    hc.hc_seed_owner(p(h), C.c_int64(h.size), C.c_uint32(world), p(owner))
    send = torch.zeros((world, 1 + cap_pair, 2), dtype=torch.int64)
    for q in range(world):
        sel = owner == q
        assert int(sel.sum()) <= cap_pair
        send[q, 0, 0] = int(sel.sum())
        send[q, 1:1 + int(sel.sum()), 0] = torch.from_numpy(h[sel].view(np.int64))
        send[q, 1:1 + int(sel.sum()), 1] = torch.from_numpy(c[sel].astype(np.int64))
    # X0: all-to-all (gloo has no all_to_all: gather everything, keep the segments addressed to this rank)
    got = [g[rank] for g in _gather(send, world)]
    ph = np.concatenate([g[1:1 + int(g[0, 0]), 0].numpy().view(np.uint64) for g in got])
    pc = np.concatenate([g[1:1 + int(g[0, 0]), 1].numpy() for g in got])
    # P1: the partition: counts of the same seed add up; statistics; dictionary; what travels
    u, inv = np.unique(ph, return_inverse=True)
    cnt = np.zeros(u.size, np.int64)
    np.add.at(cnt, inv, pc)
    assert np.all(np.diff(np.concatenate([[0], np.cumsum(cnt > 0)])) >= 0)
    multi = cnt[cnt >= 2]
    pos = np.searchsorted(dict_hash, u)
    ids = np.where((pos < dict_hash.size) & (dict_hash[np.minimum(pos, dict_hash.size - 1)] == u), pos, NONE).astype(np.int64)
    cfg_min = max(min_read_support, 1)
    possible = cnt >= cfg_min
    ones = possible & (cnt == 1) & (ids == NONE)
    emit = possible & ~ones
    g = torch.zeros((8 + cap_g, 2), dtype=torch.int64)
    n_emit = int(emit.sum())
    assert n_emit <= cap_g
    g[0, 0] = n_emit; g[0, 1] = int(ones.sum()); g[1, 0] = int(multi.sum()); g[1, 1] = int(multi.size); g[2, 0] = int(u.size); g[2, 1] = int(cnt.sum())
    g[3, 0] = int(off.size - 1)
    g[8:8 + n_emit, 0] = torch.from_numpy(cnt[emit]); g[8:8 + n_emit, 1] = torch.from_numpy(ids[emit])
    # X1 + P2: every rank finalizes the same lists
    allg = _gather(g, world)
    m_sum = sum(int(x[1, 0]) for x in allg); m_cnt = sum(int(x[1, 1]) for x in allg)
    unique = sum(int(x[2, 0]) for x in allg); total = sum(int(x[2, 1]) for x in allg); n_reads = sum(int(x[3, 0]) for x in allg)
    min_sup = min_read_support if min_read_support >= 0 else (2 if (m_cnt > 0 and m_sum / m_cnt > 3.0) else 1)
    gc = np.concatenate([x[8:8 + int(x[0, 0]), 0].numpy() for x in allg]); gi = np.concatenate([x[8:8 + int(x[0, 0]), 1].numpy() for x in allg])
    keep = (gc >= min_sup) & (gc != 0)
    n1 = sum(int(x[0, 1]) for x in allg) if min_sup <= 1 else 0
    sums = np.zeros(2)
    kc = np.ascontiguousarray(gc[keep], np.uint32)
    hc.hc_gathered_sums(p(kc), C.c_int64(kc.size), C.c_uint64(n1), p(sums))
    kept = int(keep.sum()) + n1
    in_idx = keep & (gi != NONE)
    t_hash = np.ascontiguousarray(dict_hash[gi[in_idx]], np.uint64); t_log = np.log1p(gc[in_idx].astype(np.float64))
    o = np.argsort(t_hash); t_hash, t_log = t_hash[o], t_log[o]
    N = host.n_nodes
    d = host.desc()
    scores = np.zeros((N, 5)); metrics = np.zeros((N, 5)); wc = C.c_double(); nb = C.c_uint32(); ne = C.c_uint32()
    rc = hc.hc_emulate_scoring(C.byref(d), rank, world, p(t_hash), p(t_log), C.c_int64(t_hash.size), C.c_double(kept), C.c_double(np.sqrt(sums[0])),
                               C.c_double(sums[1]), p(metrics), p(scores), C.byref(wc), C.byref(nb), C.byref(ne))
    assert rc == 0
    b, e = nb.value, ne.value
    order = cpu.bfs_order(S.parent_index)
    bfs = np.zeros(order.size, np.uint32); bfs[order] = np.arange(order.size, dtype=np.uint32)
    nodes = np.arange(b, e); nodes = nodes[np.argsort(bfs[nodes], kind="stable")]
    # records of the shard (strict prefix maxima in BFS order), X2, the chain on all of them
    rec = torch.zeros((5, 1 + rec_x, 3), dtype=torch.float64)
    for m in range(5):
        run, k = 0.0, 0
        for v in nodes:
            x = scores[v, m]
            if x > run:
                assert k < rec_x
                rec[m, 1 + k, 0] = x; rec[m, 1 + k, 1] = float(bfs[v]); rec[m, 1 + k, 2] = float(v); k += 1
            run = max(run, x)
        rec[m, 0, 0] = k
    allr = _gather(rec, world)
    best_s, best_n, tied = [], [], []
    ties = torch.full((5, 1 + tie_head), -1, dtype=torch.int64)
    lasts = []
    for m in range(5):
        rows = np.concatenate([x[m, 1:1 + int(x[m, 0, 0])].numpy() for x in allr]) if any(int(x[m, 0, 0]) for x in allr) else np.zeros((0, 3))
        rows = rows[np.argsort(rows[:, 1], kind="stable")] if rows.size else rows
        best, bn, last = 0.0, NONE, -1
        for sc_, rk_, nd_ in rows:
            if sc_ > best + max(best * 0.0001, 1e-9):
                best, bn, last = float(sc_), int(nd_), int(rk_)
        lo = best - max(best * 0.0001, 1e-9)
        loc = [int(v) for v in nodes if bfs[v] > last and scores[v, m] >= lo and scores[v, m] > 0]
        assert len(loc) <= tie_head
        ties[m, 0] = len(loc)
        if loc:
            ties[m, 1:1 + len(loc)] = torch.tensor(loc, dtype=torch.int64)
        best_s.append(best); best_n.append(bn); lasts.append(last)
    allt = _gather(ties, world)
    for m in range(5):
        t = [int(v) for x in allt for v in x[m, 1:1 + int(x[m, 0])].tolist()]
        if best_n[m] != NONE or t:
            t.append(best_n[m])
        t = np.unique(np.array(t, np.uint32))
        tied.append(t)
        if t.size:
            best_n[m] = int(t[0])
    return best_s, best_n, tied, dict(unique=unique, total=total, kept=kept, min_support=min_sup, magnitude=float(np.sqrt(sums[0])), log_sum=float(sums[1]),
                                      n_reads=n_reads, wc=wc.value)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from panmap_b200 import distributed as pmd
    from tools.synth import synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    S = synth.generate(900, 5000, 1.5, 1200, seed=33)
    reads, off = pmd.slice_reads(S.reads, S.read_offsets, rank, world)
    lo, hi = pmd.read_slice(S.read_offsets, rank, world)
    ok = (hi - lo) == off.size - 1 and int(off[0]) == 0 and int(off[-1]) == reads.size
    # the hand-off of the communicator id (bytes made on rank 0 reach every rank unchanged)
    box = [os.urandom(128) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    chk = [None] * world
    dist.all_gather_object(chk, box[0])
    ok = ok and all(c == chk[0] and len(c) == 128 for c in chk)
    for mrs in (-1, 1, 3):
        bs, bn, tied, sc = emulate_sharded(S, reads, off, rank, world, min_read_support=mrs)
        exp = cpu.place(S.reads, S.read_offsets, S, min_read_support=mrs)
        ok = ok and sc["unique"] == exp["unique_seeds"] and sc["kept"] == exp["kept"] and sc["total"] == exp["total_frequency"] and sc["min_support"] == exp["min_support"]
        ok = ok and sc["n_reads"] == 1200 and abs(sc["magnitude"] - exp["magnitude"]) <= 1e-11 * exp["magnitude"] and abs(sc["log_sum"] - exp["log_sum"]) <= 1e-11 * exp["log_sum"]
        ok = ok and all(bn[m] == int(exp["best_index"][m]) and np.array_equal(tied[m], exp["tied"][m]) and
                        abs(bs[m] - exp["best_score"][m]) <= 1e-10 * max(abs(exp["best_score"][m]), 1e-9) for m in range(5))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_protocol_two_ranks_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(out) == [(0, True), (1, True)]
