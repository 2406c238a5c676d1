"""N > 1 path on CPU: world_size-2 gloo run of panmap_b200.distributed.place_sharded with a stand-in workspace that executes
the stage protocol through the oracle + the host emulation of the kernels (tests/hostcheck).  Checks that the sharded,
exchanged result equals the single-process oracle placement."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from oracle import cpu
from tests import helpers as H

METRICS = ("log_raw", "log_cosine", "containment", "weighted_containment", "log_containment")


class Res:
    pass


class FakeWorkspace:
    """CPU stand-in with the stage_* interface of panmap_b200.api.Workspace"""

    def __init__(self, idx, shard, n_shards):
        import panmap_b200 as pm
        self.idx, self.shard, self.n_shards = idx, shard, n_shards
        self.host = pm.HostIndex(idx.hash, idx.parent, idx.child, idx.offsets, idx.parent_index, idx.k, idx.s, idx.t, idx.l)
        self.hc = C.CDLL(os.path.join(H.ROOT, "tests", "hostcheck", "libhostcheck.so"))
        order = cpu.bfs_order(idx.parent_index)
        self.rank = np.zeros(order.size, np.uint32)
        self.rank[order] = np.arange(order.size, dtype=np.uint32)

    def stage_seed(self, reads, offsets, params):
        self.h, self.c = cpu.seed_table(reads, offsets, self.idx.k, self.idx.s, self.idx.t, self.idx.l)

    def stage_table_export(self):
        return self.h, self.c

    def stage_table_import(self, h, c):
        u, inv = np.unique(h, return_inverse=True)
        cc = np.zeros(u.size, np.int64)
        np.add.at(cc, inv, c)
        self.h, self.c = u, cc

    def stage_score(self, params):
        ms = cpu.resolve_min_read_support(self.c, params.min_read_support)
        logv, sc = cpu.read_magnitudes(self.c, ms)
        N = self.host.n_nodes
        d = self.host.desc()
        self.scores = np.zeros((N, 5)); metrics = np.zeros((N, 5)); wc = C.c_double(); nb = C.c_uint32(); ne = C.c_uint32()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = self.hc.hc_emulate_scoring(C.byref(d), self.shard, self.n_shards, p(self.h), p(logv), C.c_int64(self.h.size), C.c_double(sc["kept"]),
                                        C.c_double(sc["magnitude"]), C.c_double(sc["log_sum"]), p(metrics), p(self.scores), C.byref(wc), C.byref(nb), C.byref(ne))
        assert rc == 0
        self.b, self.e = nb.value, ne.value

    def stage_records(self):
        nodes = np.arange(self.b, self.e)
        nodes = nodes[np.argsort(self.rank[nodes], kind="stable")]
        out = []
        for m in range(5):
            run = 0.0
            rr, nn, ss = [], [], []
            for v in nodes:
                x = self.scores[v, m]
                if x > run:
                    rr.append(self.rank[v]); nn.append(v); ss.append(x)
                run = max(run, x)
            out.append((np.array(rr, np.uint32), np.array(nn, np.uint32), np.array(ss, np.float64)))
        return out

    def stage_select(self, records, total_reads):
        res = Res(); res.tied, res.best_index, res.best_score = {}, {}, {}
        for m, name in enumerate(METRICS):
            rk, nd, sc = records[m]
            o = np.argsort(rk, kind="stable")
            best, bn, last = 0.0, 0xFFFFFFFF, -1
            for i in o:
                if sc[i] > best + max(best * 0.0001, 1e-9):
                    best, bn, last = float(sc[i]), int(nd[i]), int(rk[i])
            lo = best - max(best * 0.0001, 1e-9)
            loc = np.arange(self.b, self.e)
            t = [int(v) for v in loc if self.rank[v] > last and self.scores[v, m] >= lo and self.scores[v, m] > 0]
            if bn != 0xFFFFFFFF or t:
                t.append(bn)
            res.tied[name] = np.unique(np.array(t, np.uint32))
            res.best_index[name] = int(res.tied[name][0]) if len(res.tied[name]) else bn
            res.best_score[name] = best
        return res


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import panmap_b200 as pm
    from panmap_b200 import distributed as pmd
    from tools.synth import synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    S = synth.generate(900, 5000, 1.5, 1200, seed=33)
    n = 1200
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    off = S.read_offsets[lo:hi + 1] - S.read_offsets[lo]
    reads = S.reads[int(S.read_offsets[lo]):int(S.read_offsets[hi])]
    ws = FakeWorkspace(S, rank, world)
    res = pmd.place_sharded(ws, reads, off, n, pm.PlaceParams())
    exp = cpu.place(S.reads, S.read_offsets, S)
    ok = all(res.best_index[name] == int(exp["best_index"][m]) and np.array_equal(res.tied[name], exp["tied"][m]) and
             abs(res.best_score[name] - exp["best_score"][m]) <= 1e-12 * max(abs(exp["best_score"][m]), 1e-9) for m, name in enumerate(METRICS))
    # the helper gathers must round-trip ragged arrays
    parts = pmd.all_gather_var(np.arange(rank * 3 + 1, dtype=np.uint64))
    ok = ok and [len(x) for x in parts] == [r * 3 + 1 for r in range(world)]
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_place_sharded_two_ranks_gloo():
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
