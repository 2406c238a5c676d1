#!/usr/bin/env python3
"""Generates tests/golden/* from the reference's own code (oracle/_ref/libpanmap_ref.so).  Run in the build container
(needs /root/reference); the outputs are committed so the oracle stays pinned where the reference is absent."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from tests import helpers as H  # noqa: E402


def main():
    os.makedirs(H.GOLDEN, exist_ok=True)
    rng = np.random.default_rng(20260101)
    # 1. rollingSyncmers / hashSeq vectors
    cases = []
    for i in range(60):
        k = int(rng.integers(5, 33)); s = int(rng.integers(1, k + 1)); t = int(rng.integers(0, k - s + 1)); op = bool(rng.integers(0, 2))
        if i < 10:
            k, s, t, op = [(19, 8, 0, False), (15, 8, 0, False)][i % 2]
        seq = H.random_reads(rng, 1, lo=k, hi=260, p_n=0.02, p_lower=0.03)[0]
        h, rev, syn, pos = ref.rolling_syncmers(seq, k, s, op, t, True)
        cases.append(dict(seq=seq.decode(), k=k, s=s, t=t, open=op, hash=h, rev=rev, syn=syn))
    np.savez_compressed(os.path.join(H.GOLDEN, "rolling_syncmers.npz"), n=len(cases),
                        **{f"{key}_{i}": np.array(c[key]) for i, c in enumerate(cases) for key in c})
    hs = []
    for k in (5, 15, 19, 31, 32):
        for _ in range(5):
            seq = "".join(rng.choice(list("ACGT"), size=k))
            f, r = ref.hash_seq(seq)
            hs.append((seq, f, r))
    np.savez_compressed(os.path.join(H.GOLDEN, "hash_seq.npz"), seq=np.array([x[0] for x in hs]), f=np.array([x[1] for x in hs], np.uint64),
                        r=np.array([x[2] for x in hs], np.uint64))
    # 2. tolerance chain vectors
    ch = []
    for i in range(40):
        n = int(rng.integers(1, 200))
        base = rng.random() * 10
        sc = base * (1 + rng.normal(0, 2e-4, size=n)) * (rng.random(n) > 0.2)
        if i % 5 == 0:
            sc = sc * 1e-10
        order = rng.permutation(n).astype(np.uint32)
        bs, bi, tied = ref.select_chain(order, sc)
        ch.append((order, sc, bs, bi, tied))
    np.savez_compressed(os.path.join(H.GOLDEN, "select_chain.npz"), n=len(ch), **{f"order_{i}": c[0] for i, c in enumerate(ch)},
                        **{f"score_{i}": c[1] for i, c in enumerate(ch)}, **{f"best_{i}": np.array([c[2]]) for i, c in enumerate(ch)},
                        **{f"idx_{i}": np.array([c[3]], np.uint32) for i, c in enumerate(ch)}, **{f"tied_{i}": c[4] for i, c in enumerate(ch)})
    # 3. config 1: sars_20000 + isolate reads through the reference placeLite
    if not os.path.exists(H.SARS_IDX):
        ref.build_index("/root/reference/examples/data/panmans/sars_20000_twilight_dipper.panman", H.SARS_IDX)
    R = ref.RefIndex(H.SARS_IDX)
    rp = R.place("/root/reference/examples/data/reads/isolate_R1.fastq.gz", "/root/reference/examples/data/reads/isolate_R2.fastq.gz",
                 out_tsv=os.path.join(H.GOLDEN, "isolate.placement.tsv"))
    metrics, scores, scal = R.node_metrics(rp["table_hash"], rp["table_count"], -1)
    nodes = np.unique(np.concatenate([np.arange(0, R.n_nodes, 97), rp["best_index"], np.concatenate(rp["tied"])])).astype(np.int64)
    np.savez_compressed(os.path.join(H.GOLDEN, "sars_isolate_node_metrics_sample.npz"), nodes=nodes, metrics=metrics[nodes], scores=scores[nodes])
    top = np.argsort(-rp["table_count"], kind="stable")[:2000]
    np.savez_compressed(os.path.join(H.GOLDEN, "sars_isolate_summary.npz"), best_score=rp["best_score"], best_index=rp["best_index"],
                        **{f"tied_{m}": rp["tied"][m] for m in range(5)}, kept=rp["kept"], unique_seeds=rp["unique_seeds"],
                        total_frequency=rp["total_frequency"], magnitude=rp["magnitude"], min_support=scal["min_support"],
                        log_sum=scal["log_sum"], wc_denominator=scal["wc_denominator"],
                        table_xor=np.bitwise_xor.reduce(rp["table_hash"]), table_count_sum=rp["table_count"].sum(),
                        top_hash=rp["table_hash"][top], top_count=rp["table_count"][top])
    # first 400 isolate reads: full seed table (small enough to commit)
    print("golden written to", H.GOLDEN)


if __name__ == "__main__":
    main()
