#!/usr/bin/env python3
"""Index-builder timing (GPU box): pm_index_build (genomes seeded on the GPU) against the reference's IndexBuilder (oracle/_ref, host cores)
on the bundled PanMANs, --flank-mask 0.  usage: build_probe.py [threads for the reference builder]"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panmap_b200 as pm
from oracle import ref
from tests import helpers as H

threads = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for name, pan, sp in (("extended_mammoth", H.MAMMOTH_PANMAN, dict(k=15, s=8, t=0, l=1)), ("rsv_4K", H.RSV_PANMAN, dict(k=19, s=8, t=0, l=3)),
                      ("sars_20000", H.SARS_PANMAN, dict(k=19, s=8, t=0, l=3))):
    pm.HostIndex.build_from_panman(H.MAMMOTH_PANMAN, k=15, s=8, t=0, l=1)   # warm the context
    t0 = time.perf_counter(); B = pm.HostIndex.build_from_panman(pan, **sp); t_gpu = time.perf_counter() - t0
    line = f"{name}: {B.n_nodes} nodes, {B.n_deltas} deltas; pm_index_build {t_gpu:.2f} s"
    if ref.available():
        with tempfile.TemporaryDirectory() as d:
            for th in ((1, threads) if threads > 0 else ()):
                t0 = time.perf_counter(); ref.build_index(pan, os.path.join(d, "r.idx"), flank_mask=0, threads=th, **sp); line += f"; reference IndexBuilder {th} thread(s) {time.perf_counter() - t0:.2f} s"
    print(line, flush=True)
