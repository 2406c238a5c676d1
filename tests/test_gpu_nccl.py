"""GPU parity test (-m gpu, needs >= 2 devices) of the NCCL transport: one process per GPU, the communicator id handed over through
torch.distributed (gloo), then pm_place_sharded on every rank -- seeding of the rank's read slice, all-to-all of the seed partitions,
all-gathers of the finalized pairs / records / tie heads, all enqueued by the library on its own stream.  Every rank must return the
oracle's placement (integers, best nodes and tie lists bit-exact, scores to 1e-12), from host buffers and from resident reads,
also right after a much smaller sample (capacities regrown in lock-step on all ranks).  Both transports of the exchanges are run:
NCCL collectives, and the peer-memory one (every rank's receive buffers mapped into the others through CUDA IPC, payloads stored over
NVLink, NCCL only as the bootstrap), which must actually be the one in use when it is asked for."""
import os
import socket

import numpy as np
import pytest

import panmap_b200 as pm

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q, peer):
    try:
        os.environ["PM_PEER_EXCHANGE"] = peer
        import torch.distributed as dist
        from oracle import cpu
        from panmap_b200 import distributed as pmd
        from tests import helpers as H
        from tools.synth import synth
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        S = synth.generate(6000, 8000, 1.5, 30000, seed=21)
        host = pm.HostIndex(S.hash, S.parent, S.child, S.offsets, S.parent_index, S.k, S.s, S.t, S.l)
        ws = pm.Workspace(pm.Index(host, device=rank, shard=rank, n_shards=world))
        comm = pmd.make_comm(ws)
        ok = True
        exp = cpu.place(S.reads, S.read_offsets, S, want_scores=True)
        small = synth.generate(6000, 8000, 1.5, 300, seed=21)      # same index, tiny sample first: capacities must regrow afterwards
        r0, o0 = pmd.slice_reads(small.reads, small.read_offsets, rank, world)
        e0 = cpu.place(small.reads, small.read_offsets, S)
        res = comm.place_sharded(r0, o0)
        ok = ok and all(res.best_index[n] == int(e0["best_index"][m]) and np.array_equal(res.tied[n], e0["tied"][m]) for m, n in enumerate(pm.METRICS))
        reads, off = pmd.slice_reads(S.reads, S.read_offsets, rank, world)
        for mode in ("host", "resident", "resident"):
            if mode == "host":
                res = comm.place_sharded(reads, off)
            else:
                ws.upload(reads, off)
                res = comm.place_sharded_resident()
            ok = ok and res.raw.unique_seeds == exp["unique_seeds"] and res.raw.read_unique_seed_count == exp["kept"] and res.raw.total_reads == 30000
            ok = ok and res.raw.total_read_seed_frequency == exp["total_frequency"] and res.raw.min_read_support == exp["min_support"]
            ok = ok and H.relerr(res.raw.read_magnitude, exp["magnitude"]).max() < 1e-12 and H.relerr(res.raw.log_containment_denominator, exp["log_sum"]).max() < 1e-12
            b, e = ws.index.shard_range()
            ok = ok and H.relerr(ws.node_scores()[b:e], exp["scores"][b:e]).max() < 1e-12
            for m, name in enumerate(pm.METRICS):
                ok = ok and res.best_index[name] == int(exp["best_index"][m]) and np.array_equal(res.tied[name], exp["tied"][m])
                ok = ok and H.relerr(res.best_score[name], exp["best_score"][m]).max() < 1e-12
        sent, recv = comm.last_traffic()
        ok = ok and sent > 0 and recv > 0
        tr = comm.transport()
        ok = ok and (("peer-memory" in tr) if peer == "1" else tr == "nccl")
        comm.close()
        q.put((rank, bool(ok), "" if ok else f"transport={tr}"))
        dist.destroy_process_group()
    except Exception as e:  # surface the failure in the parent
        import traceback
        q.put((rank, False, traceback.format_exc()))


@pytest.mark.parametrize("peer", ["1", "0"])
def test_nccl_sharded_sample_matches_oracle_on_every_rank(peer):
    world = pm.device_count()
    if world < 2:
        pytest.skip("needs at least two CUDA devices")
    world = min(world, 8)
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q, peer)) for r in range(world)]
    for p in ps:
        p.start()
    out = [q.get(timeout=600) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(o[1] for o in out), [o for o in out if not o[1]]
