#!/usr/bin/env python3
"""Phase times of FASTQ file -> result through the C++ shim (PM_INGEST_TIMING=1: scan / landing buffers / fill / pm_place on stderr).
usage: PM_INGEST_TIMING=1 python tools/ingest_probe.py [c3|c3-small|c1]"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import panmap_b200 as pm  # noqa: E402

S, w = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
host = getattr(S, "host", None) or pm.HostIndex(S.hash, S.parent, S.child, S.offsets, S.parent_index, S.k, S.s, S.t, S.l)
ws = pm.Workspace(pm.Index(host))
with tempfile.TemporaryDirectory() as td:
    fq, fq2 = S.fastq if hasattr(S, "fastq") else (bench.write_fastq(S, w["n_reads"], os.path.join(td, "r.fastq")), "")
    for i in range(4):
        t = time.perf_counter()
        pm.place_files(ws, fq, fq2, os.path.join(td, "o.tsv"))
        print(f"call {i}: {1e3 * (time.perf_counter() - t):.2f} ms", file=sys.stderr)
