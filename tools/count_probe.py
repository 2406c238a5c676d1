#!/usr/bin/env python3
"""Counting experiments on one resident sample: direct counting against partitioned counting at several bucket sizes.
usage: tools/count_probe.py [workload=c4] [steps=4]   (prints step / seeding / per-kernel times per setting)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench  # noqa: E402
import panmap_b200 as pm  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c4"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    S, w = bench.make_workload(name)
    host = pm.HostIndex(S.hash, S.parent, S.child, S.offsets, S.parent_index, S.k, S.s, S.t, S.l)
    index = pm.Index(host)
    params = pm.PlaceParams()
    settings = [("direct", {"PM_BUCKET_MIN_SLOTS": str(1 << 40)})] + [(f"buckets of {mb} MB", {"PM_BUCKET_MIN_SLOTS": str(1 << 22), "PM_BUCKET_BYTES": str(mb << 20)}) for mb in (8, 16, 32, 64)]
    base = None
    for label, env in settings:
        for k in ("PM_BUCKET_MIN_SLOTS", "PM_BUCKET_BYTES"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ws = pm.Workspace(index); ws.stage_timers(True)
        ws.upload(S.reads, S.read_offsets)
        for _ in range(3):
            r = ws.place_resident(params, full=False)
        st = np.zeros(8); km = np.zeros(3)
        for _ in range(steps):
            r = ws.place_resident(params, full=False)
            st += np.array(list(r.stage_ms)); km += np.array(ws.last_kernel_ms())
        st /= steps; km /= steps
        sig = (tuple(int(r.best_index[i]) for i in range(5)), int(r.unique_seeds), int(r.read_unique_seed_count), tuple(float(r.best_score[i]) for i in range(5)))
        base = base or sig
        print(f"{label:>18s}: step {st[7]:8.3f} ms  seeding {st[1]:8.3f}  syncmers {km[1]:7.3f}  counting {km[2]:8.3f}  finalize {st[2]:6.3f}  same result: {sig == base}", flush=True)
        del ws


if __name__ == "__main__":
    main()
