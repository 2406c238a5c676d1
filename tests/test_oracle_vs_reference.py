"""Pins oracle/panmap_oracle.c against the reference's own translation units (oracle/_ref/libpanmap_ref.so, built by
oracle/ref_build/Makefile from /root/reference).  Skipped where the library has not been built."""
import os
import tempfile

import numpy as np
import pytest

from oracle import cpu, ref
from tests import helpers as H

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libpanmap_ref.so not built")


def test_rolling_syncmers_bit_exact_random():
    rng = np.random.default_rng(42)
    for it in range(400):
        k = int(rng.integers(2, 33)); s = int(rng.integers(1, k + 1)); t = int(rng.integers(0, k - s + 1)); op = bool(rng.integers(0, 2))
        seq = H.random_reads(rng, 1, lo=1, hi=300, p_n=0.03, p_lower=0.03)[0]
        for ra in (True, False):
            a = cpu.rolling_syncmers(seq, k, s, op, t, ra); b = ref.rolling_syncmers(seq, k, s, op, t, ra)
            assert all(np.array_equal(x, y) for x, y in zip(a, b)), (k, s, t, op, ra, seq)


def test_select_chain_matches_reference_including_tolerance_edges():
    rng = np.random.default_rng(7)
    for it in range(300):
        n = int(rng.integers(1, 120))
        base = float(rng.random() * 100)
        sc = base * (1 + rng.choice([0, 1e-4, -1e-4, 0.99e-4, 1.01e-4, 5e-5], size=n) * rng.integers(0, 3, size=n))
        sc = sc * (rng.random(n) > 0.15)
        if it % 7 == 0:
            sc = sc * 1e-11
        order = rng.permutation(n).astype(np.uint32)
        assert_same(cpu.select_chain(order, sc), ref.select_chain(order, sc))


def assert_same(a, b):
    assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2])


def test_node_metrics_bit_identical_on_synthetic_index():
    rng = np.random.default_rng(3)
    idx, hashes, genomes = H.synthetic_index(1500, rng, big_node=(20, 700))
    th = np.sort(rng.choice(hashes, size=900, replace=False)).astype(np.uint64)
    tc = rng.integers(1, 40, size=th.size).astype(np.int64)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "s.idx")
        ref.write_index(p, idx)
        R = ref.RefIndex(p)
        rm, rs, rscal = R.node_metrics(th, tc, -1)
        R.close()
    ms = cpu.resolve_min_read_support(tc, -1)
    logv, sc = cpu.read_magnitudes(tc, ms)
    assert ms == rscal["min_support"] and sc["kept"] == rscal["kept"] and sc["total"] == rscal["total_frequency"]
    denW = cpu.weighted_denominator(idx, th, logv)
    om, osc = cpu.node_metrics(idx, th, logv, sc["kept"], sc["magnitude"], sc["log_sum"], denW)
    assert np.array_equal(om, rm)                       # all 7 accumulators, every node, bit for bit
    assert H.relerr(denW, rscal["wc_denominator"]) < 1e-13
    assert H.relerr(osc, rs).max() < 1e-12              # scores differ only through the hash-map-order denominators


def test_place_matches_reference_on_synthetic_genome_index():
    from tools.synth import synth
    import bench
    for (k, s, l, lam, seed) in [(19, 8, 3, 1.5, 11), (15, 8, 1, 3.0, 12)]:
        S = synth.generate(1500, 6000, lam, 4000, k=k, s=s, l=l, seed=seed)
        with tempfile.TemporaryDirectory() as td:
            dt, r = bench.reference_step(S, 4000, 1, td, cache={})
        o = cpu.place(S.reads, S.read_offsets, S)
        assert np.array_equal(o["best_index"], r["best_index"])
        for m in range(5):
            assert np.array_equal(o["tied"][m], r["tied"][m])
        assert o["kept"] == r["kept"] and o["unique_seeds"] == r["unique_seeds"] and o["total_frequency"] == r["total_frequency"]
        assert H.relerr(o["best_score"], r["best_score"]).max() < 1e-12
        eh, ec = cpu.seed_table(S.reads, S.read_offsets, k, s, 0, l)
        assert np.array_equal(eh, r["table_hash"]) and np.array_equal(ec, r["table_count"])


def _fastq(path, reads):
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)))


def test_dedup_trim_force_leaf_and_hpc_match_reference_placeLite():
    """non-default options of the path against the reference's own placeLite: --dedup (placement.cpp:1550-1620), trims, --force-leaf,
    and an index flagged hpc (reads are homopolymer-compressed first, placement.cpp:1145-1165)"""
    from tools.synth import synth
    rng = np.random.default_rng(21)
    S = synth.generate(1200, 6000, 1.5, 3000, k=19, s=8, l=3, seed=5)
    off = S.read_offsets.astype(np.int64); buf = S.reads.tobytes()
    reads = [buf[off[i]:off[i + 1]] for i in range(3000)]
    reads = reads + reads[:700] + [r.lower() for r in reads[:50]]            # duplicates, and case variants that are NOT duplicates
    reads = [reads[i] for i in rng.permutation(len(reads))]
    stretched = [b"".join(bytes([c]) * int(rng.choice([1, 1, 2, 3])) for c in r) for r in reads[:1500]]
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "s.idx"); ref.write_index(p, S)
        fq = os.path.join(td, "r.fastq"); _fastq(fq, reads)
        R = ref.RefIndex(p)
        import panmap_b200 as pm
        rb, ro = pm.pack_reads(reads)
        for kw, okw in [(dict(dedup=True), dict(dedup=True)), (dict(trim_start=6, trim_end=11), dict(trim_start=6, trim_end=11)),
                        (dict(force_leaf=True, dedup=True), dict(force_leaf=True, dedup=True))]:
            r = R.place(fq, **kw)
            o = cpu.place(rb, ro, S, **okw)
            assert np.array_equal(o["best_index"], r["best_index"]), kw
            assert all(np.array_equal(o["tied"][m], r["tied"][m]) for m in range(5)), kw
            assert o["kept"] == r["kept"] and o["unique_seeds"] == r["unique_seeds"] and o["total_frequency"] == r["total_frequency"], kw
            assert H.relerr(o["best_score"], r["best_score"]).max() < 1e-12
        R.close()
        S.hpc = 1
        ph = os.path.join(td, "h.idx"); ref.write_index(ph, S)
        S.hpc = 0
        fqh = os.path.join(td, "h.fastq"); _fastq(fqh, stretched)
        R = ref.RefIndex(ph)
        r = R.place(fqh)
        R.close()
        cb, co = pm.pack_reads([cpu.hpc_compress(x) for x in stretched])
        o = cpu.place(cb, co, S)
        assert np.array_equal(o["best_index"], r["best_index"])
        assert all(np.array_equal(o["tied"][m], r["tied"][m]) for m in range(5))
        assert o["kept"] == r["kept"] and o["unique_seeds"] == r["unique_seeds"] and o["total_frequency"] == r["total_frequency"]
        eh, ec = cpu.seed_table(cb, co, 19, 8, 0, 3)
        assert np.array_equal(eh, r["table_hash"]) and np.array_equal(ec, r["table_count"])


def test_seed_mask_fraction_matches_reference_placeLite_at_a_tie_free_cut():
    """--seed-mask-fraction (placement.cpp:1748-1799): the reference sorts by count only, so the masked set is only defined when
    the cut does not fall inside a group of equal counts; fractions are chosen so that it does not, and then everything must match"""
    from tools.synth import synth
    import panmap_b200 as pm
    S = synth.generate(900, 5000, 1.5, 4000, k=19, s=8, l=3, seed=9)
    off = S.read_offsets.astype(np.int64); buf = S.reads.tobytes()
    reads = [buf[off[i]:off[i + 1]] for i in range(4000)]
    rb, ro = pm.pack_reads(reads)
    eh, ec = cpu.seed_table(rb, ro, 19, 8, 0, 3)
    U = eh.size
    desc = np.sort(ec)[::-1]
    cuts = [int(m) for m in np.nonzero(desc[:-1] > desc[1:])[0] + 1]          # M seeds masked <=> desc[M-1] > desc[M]
    assert len(cuts) >= 3
    picks = [cuts[0], cuts[len(cuts) // 2], cuts[-1]]
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "s.idx"); ref.write_index(p, S)
        fq = os.path.join(td, "r.fastq"); _fastq(fq, reads)
        R = ref.RefIndex(p)
        for M in picks:
            frac = (M + 0.5) / U
            assert int(frac * U) == M
            r = R.place(fq, seed_mask_fraction=frac)
            o = cpu.place(rb, ro, S, seed_mask_fraction=frac)
            assert o["unique_seeds"] == r["unique_seeds"] == U - M
            assert o["kept"] == r["kept"] and o["total_frequency"] == r["total_frequency"], M
            assert np.array_equal(o["best_index"], r["best_index"]), M
            assert all(np.array_equal(o["tied"][m], r["tied"][m]) for m in range(5)), M
            assert H.relerr(o["best_score"], r["best_score"]).max() < 1e-12
            mh, mc = cpu.mask_top_seeds(eh, ec, frac)
            assert np.array_equal(mh, r["table_hash"]) and np.array_equal(mc, r["table_count"]), M
        R.close()


def _fastq_q(path, reads, quals):
    with open(path, "wb") as f:
        for i, (r, q) in enumerate(zip(reads, quals)):
            f.write(b"@r%d\n%s\n+\n%s\n" % (i, r, q))


@pytest.mark.parametrize("l", [3, 1, 0])
def test_min_seed_quality_path_matches_reference_placeLite(l):
    """--min-seed-quality > 0 (placement.cpp:1179-1240 for l == 0, :1388-1533 for l >= 1): syncmers whose k bases average below the
    threshold (or start in a trimmed flank) drop out, k-min-mers need l consecutive passing syncmers, reads are not deduplicated"""
    from tools.synth import synth
    import panmap_b200 as pm
    rng = np.random.default_rng(40 + l)
    S = synth.generate(700, 5000, 1.5, 2500, k=19, s=8, l=max(l, 1), seed=11)
    S.l = l
    off = S.read_offsets.astype(np.int64); buf = S.reads.tobytes()
    reads = [buf[off[i]:off[i + 1]] for i in range(2500)]
    reads = reads + reads[:300] + [b"", b"ACGT", b"ACGTNACGTTGCATGCATGCATGCAACGGTCA" * 3]
    quals = []
    for r in reads:                       # blocks of good and bad quality so that windows straddle the threshold
        q = np.repeat(rng.choice([2, 12, 19, 20, 21, 30, 40], size=len(r) // 7 + 1), 7)[:len(r)] + rng.integers(0, 3, len(r))
        quals.append(bytes((q + 33).astype(np.uint8)))
    rb, ro = pm.pack_reads(reads)
    qb, _ = pm.pack_reads(quals)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "s.idx"); ref.write_index(p, S)
        fq = os.path.join(td, "r.fastq"); _fastq_q(fq, reads, quals)
        R = ref.RefIndex(p)
        for kw in [dict(min_seed_quality=20), dict(min_seed_quality=13, trim_start=5, trim_end=9), dict(min_seed_quality=31, dedup=True)]:
            r = R.place(fq, **kw)
            o = cpu.place(rb, ro, S, quals=qb, **kw)
            assert o["unique_seeds"] == r["unique_seeds"] and o["unique_seeds"] > 0, kw
            assert o["kept"] == r["kept"] and o["total_frequency"] == r["total_frequency"], kw
            assert np.array_equal(o["best_index"], r["best_index"]), kw
            assert all(np.array_equal(o["tied"][m], r["tied"][m]) for m in range(5)), kw
            assert H.relerr(o["best_score"], r["best_score"]).max() < 1e-12
            eh, ec = cpu.seed_table(rb, ro, 19, 8, 0, l, quals=qb, min_seed_quality=kw["min_seed_quality"],
                                    trim_start=kw.get("trim_start", 0), trim_end=kw.get("trim_end", 0))
            assert np.array_equal(eh, r["table_hash"]) and np.array_equal(ec, r["table_count"]), kw
        # the filter must have removed something, or the test proves nothing
        fh, fc = cpu.seed_table(rb, ro, 19, 8, 0, l)
        assert fc.sum() > ec.sum()
        R.close()
