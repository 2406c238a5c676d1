#!/usr/bin/env python3
"""One C3-shaped sample placed by N in-process ranks that all sit on device 0 (local transport): every kernel of the sharded data plane
at its per-rank size, for an ncu launch list (times are per rank; the ranks run one after the other on the one GPU).
usage: tools/multi_probe.py <n_ranks> [steps] [workload]   (workload: a bench.py name, default c3; the result is compared with the
one-GPU placement of the same sample: best nodes, scores and tie lists)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench  # noqa: E402
import panmap_b200 as pm  # noqa: E402
from panmap_b200 import distributed as pmd  # noqa: E402


def main():
    n = int(sys.argv[1]); steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    S, w = bench.make_workload(sys.argv[3] if len(sys.argv) > 3 else "c3")
    host = pm.HostIndex(S.hash, S.parent, S.child, S.offsets, S.parent_index, S.k, S.s, S.t, S.l)
    wss = [pm.Workspace(pm.Index(host, device=0, shard=r, n_shards=n)) for r in range(n)]
    comms = pm.Comm.local(wss)
    for w_ in wss:
        w_.stage_timers(True)
    for r in range(n):
        reads, off = pmd.slice_reads(S.reads, S.read_offsets, r, n)
        wss[r].upload(reads, off)
    params = pm.PlaceParams()
    for _ in range(steps):
        res = pm.place_multi_resident(comms, params, full=False)
    print("stage_ms", [round(x, 4) for x in res.stage_ms])
    full = pm.place_multi_resident(comms, params)
    one = pm.Workspace(pm.Index(host, device=0))
    ref = one.place(S.reads, S.read_offsets, params)
    same = all(full.best_index[m] == ref.best_index[m] and full.best_score[m] == ref.best_score[m] and np.array_equal(full.tied[m], ref.tied[m]) for m in pm.METRICS)
    print("same_as_one_gpu", same, {m: int(full.best_index[m]) for m in pm.METRICS}, "truth", int(S.truth), "one-gpu ms", round(ref.stage_ms[7], 3))


if __name__ == "__main__":
    main()
