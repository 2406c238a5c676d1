#!/usr/bin/env python3
"""Tuning probe (GPU box): seeding kernel times on resident reads and host-buffer e2e times for the current PM_AGG_MIN_READS /
PM_SLICES_ASCII / PM_SLICES_PACKED environment.  usage: tune_probe.py <n_reads> [resident|e2e]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panmap_b200 as pm
from tools.synth import synth

n = int(sys.argv[1]); mode = sys.argv[2] if len(sys.argv) > 2 else "resident"
S = synth.generate(100_000, 30_000, 1.0, n, read_len=150, seed=0)
host = pm.HostIndex(S.hash, S.parent, S.child, S.offsets, S.parent_index, S.k, S.s, S.t, S.l)
ws = pm.Workspace(pm.Index(host))
ws.stage_timers(os.environ.get("PM_STAGE_EVENTS", "1") != "0")   # per-kernel events on unless the A/B is about them
p = pm.PlaceParams()
tag = " ".join(f"{k}={os.environ[k]}" for k in ("PM_AGG_MIN_READS", "PM_COUNT_WARP_BELOW", "PM_SLICES_ASCII", "PM_SLICES_PACKED", "PM_RANK_STAGE", "PM_RESIDENT_SLICES", "PM_TABLE_FIT_LO", "PM_LIST_PREFETCH", "PM_SIDE_SCALARS", "PM_SIDE_CLEAR", "PM_STAGE_EVENTS") if k in os.environ)
if mode == "resident":
    ws.upload(S.reads, S.read_offsets)
    for _ in range(4):
        ws.place_resident(p, full=False)
    k = np.zeros(3); tot = 0.0
    for _ in range(10):
        r = ws.place_resident(p, full=False); k += np.array(ws.last_kernel_ms()); tot += r.stage_ms[7]
    print(f"n={n} resident  syncmers {k[1] / 10:.4f}  count {k[2] / 10:.4f}  step {tot / 10:.4f} ms  [{tag}]")
else:
    L = pm.lib()
    def pin(a):
        q = L.pm_host_alloc(a.nbytes + 64); C.memmove(q, a.ctypes.data, a.nbytes); return q
    hr, ho = pin(S.reads), pin(S.read_offsets)
    hp = pin(pm.host_pack_reads(S.reads, S.read_offsets))
    for name, fn in (("ascii", lambda: ws.place_raw(hr, ho, n, p)), ("packed", lambda: ws.place_packed_raw(hp, ho, n, p))):
        for _ in range(4):
            fn()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        print(f"n={n} e2e {name} {(time.perf_counter() - t0) * 100:.4f} ms  [{tag}]")
