"""GPU parity tests (-m gpu): the CUDA path through the C ABI vs the CPU oracle on the same seeded inputs.
Integer / byte / index results are bit-exact; f64 scores agree to 1e-12 relative (absolute floor 1e-9 for values that
are mathematically zero), the tolerance BASELINE.json's north_star states."""
import os

import numpy as np
import pytest

import panmap_b200 as pm
from oracle import cpu
from tests import helpers as H

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def test_rolling_syncmers_matches_oracle():
    rng = np.random.default_rng(11)
    for (k, s, t, op) in [(19, 8, 0, False), (15, 8, 0, False), (31, 8, 3, False), (19, 8, 2, True), (32, 1, 0, False), (8, 8, 0, False), (5, 2, 3, True)]:
        reads = H.random_reads(rng, 300, lo=1, hi=200) + [b"", b"A", b"N" * 50, b"ACGT" * 100, b"a" * 60]
        got = pm.rolling_syncmers(reads, k, s, op, t)
        for r, (h, rev, pos) in zip(reads, got):
            eh, erev, esyn, epos = cpu.rolling_syncmers(r, k, s, op, t, return_all=False)
            assert np.array_equal(h, eh) and np.array_equal(rev, erev) and np.array_equal(pos, epos), (k, s, t, op, r)


def test_read_seeds_match_oracle_including_trim_and_long_reads():
    rng = np.random.default_rng(12)
    reads = H.random_reads(rng, 400, lo=10, hi=180) + H.random_reads(rng, 3, lo=5000, hi=16000, p_n=0.0005)
    for (k, s, t, l, op, ts, te) in [(19, 8, 0, 3, False, 0, 0), (15, 8, 0, 1, False, 0, 0), (19, 8, 0, 3, False, 7, 12),
                                     (21, 10, 1, 2, True, 0, 0), (19, 8, 0, 0, False, 0, 5), (31, 8, 0, 5, False, 0, 0)]:
        got = pm.read_seeds(reads, k, s, t, l, op, ts, te)
        for r, g in zip(reads, got):
            e = cpu.read_seeds(r, k, s, t, l, op, ts, te)
            assert np.array_equal(g, e), (k, s, t, l, op, ts, te, len(r))


def _place_and_check(idx, reads, params=None, **okw):
    buf, off = pm.pack_reads(reads)
    host = pm.HostIndex(idx.hash, idx.parent, idx.child, idx.offsets, idx.parent_index, idx.k, idx.s, idx.t, idx.l, idx.open)
    index = pm.Index(host)
    ws = pm.Workspace(index)
    res = ws.place(buf, off, params)
    exp = cpu.place(buf, off, idx, want_scores=True, **okw)
    # integers: bit-exact
    assert res.raw.unique_seeds == exp["unique_seeds"]
    assert res.raw.read_unique_seed_count == exp["kept"]
    assert res.raw.total_read_seed_frequency == exp["total_frequency"]
    assert res.raw.min_read_support == exp["min_support"]
    th, tc = ws.seed_table()
    eh, ec = cpu.seed_table(buf, off, idx.k, idx.s, idx.t, idx.l, idx.open, okw.get("trim_start", 0), okw.get("trim_end", 0), okw.get("dedup", False))
    if okw.get("seed_mask_fraction", 0) > 0:
        eh, ec = cpu.mask_top_seeds(eh, ec, okw["seed_mask_fraction"])
    keep = tc > 0
    assert np.array_equal(th[keep], eh) and np.array_equal(tc[keep], ec)
    # f64
    assert H.relerr(res.raw.read_magnitude, exp["magnitude"]).max() < RTOL
    assert H.relerr(res.raw.log_containment_denominator, exp["log_sum"]).max() < RTOL
    assert H.relerr(res.raw.weighted_containment_denominator, exp["wc_denominator"]).max() < RTOL
    sc = ws.node_scores()
    assert H.relerr(sc, exp["scores"]).max() < RTOL
    for m, name in enumerate(pm.METRICS):
        assert H.relerr(res.best_score[name], exp["best_score"][m]).max() < RTOL, name
        assert res.best_index[name] == exp["best_index"][m], name
        assert np.array_equal(res.tied[name], exp["tied"][m]), name
    return res, exp, ws, index


def test_place_synthetic_index_with_table_only_reads():
    """reads drawn so that seeds rarely hit the synthetic index: exercises table/filters/empty intersections"""
    rng = np.random.default_rng(5)
    idx, _, _ = H.synthetic_index(700, rng)
    reads = H.random_reads(rng, 500)
    _place_and_check(idx, reads)


@pytest.mark.parametrize("k,s,t,l,op", [(21, 10, 1, 2, True), (15, 8, 0, 1, False), (31, 8, 0, 5, False), (15, 8, 0, 3, False), (12, 5, 2, 0, False)])
def test_place_other_seeding_parameters(k, s, t, l, op):
    """parameter sets besides panmap's default: the generic syncmer kernel + pack_reads + the flat counting kernel (any k <= 32, s, t,
    open/closed, l), and the specialised k=15 kernels"""
    rng = np.random.default_rng(100 + k + l)
    idx, _, _ = H.synthetic_index(300, rng, k=k, s=s, t=t, l=l)
    idx.open = int(op)
    _place_and_check(idx, H.random_reads(rng, 300, lo=10, hi=220))


def test_place_dedup_counts_every_distinct_read_string_once():
    """--dedup (placement.cpp:1550-1620): byte-identical reads count once; reads that differ only in case or in the ambiguity
    letter are different strings for the reference and must both count.  Small and sliced (>= 65536 reads) paths."""
    rng = np.random.default_rng(8)
    idx, _, _ = H.synthetic_index(400, rng)
    base = H.random_reads(rng, 300)
    reads = base + base[:120] + [r.lower() for r in base[:40]] + [base[7]] * 9 + [b"", b"", b"ACGTN" * 9, b"ACGTR" * 9, b"ACGTN" * 9]
    order = rng.permutation(len(reads))
    reads = [reads[i] for i in order]
    _place_and_check(idx, reads, pm.PlaceParams(dedup_reads=1), dedup=True)
    big = [base[i % 300] for i in range(70000)] + H.random_reads(rng, 500)
    os.environ["PM_SLICE_MIN_BYTES"] = "1000000"          # the sliced host-buffer pipeline (normally samples >= 48 MB): the read set spans slices
    try:
        _place_and_check(idx, big, pm.PlaceParams(dedup_reads=1), dedup=True)
    finally:
        del os.environ["PM_SLICE_MIN_BYTES"]


@pytest.mark.parametrize("frac", [0.0004, 0.013, 0.2, 0.77, 1.0])
def test_place_seed_mask_fraction_drops_the_most_frequent_seeds(frac):
    """--seed-mask-fraction (placement.cpp:1748-1799): the floor(frac * U) most frequent seeds leave the table before the min-support
    rule and the magnitudes; ties at the cut go by ascending hash like the oracle (the reference leaves them unspecified).  Duplicated
    reads give a wide spread of counts, so the cut lands inside groups of equal counts as well as between them."""
    rng = np.random.default_rng(31)
    idx, _, _ = H.synthetic_index(500, rng)
    base = H.random_reads(rng, 600)
    reads = base + base[:300] * 2 + base[:80] * 5 + base[:9] * 40
    res, exp, ws, _ = _place_and_check(idx, reads, pm.PlaceParams(seed_mask_fraction=frac), seed_mask_fraction=frac)
    full = cpu.place(*pm.pack_reads(reads), idx)
    assert exp["unique_seeds"] == full["unique_seeds"] - int(frac * full["unique_seeds"])
    # the debug re-run of the scoring stages must not mask a second time
    assert H.relerr(ws.node_scores(), exp["scores"]).max() < RTOL
    ws.node_metrics()
    assert ws.place(*pm.pack_reads(reads), pm.PlaceParams(seed_mask_fraction=frac)).raw.unique_seeds == exp["unique_seeds"]


def _quality_strings(rng, reads):
    """blocks of good and bad base qualities so that k-mer windows straddle the threshold"""
    out = []
    for r in reads:
        q = np.repeat(rng.choice([2, 12, 19, 20, 21, 30, 40], size=len(r) // 7 + 1), 7)[:len(r)] + rng.integers(0, 3, len(r))
        out.append(bytes((q + 33).astype(np.uint8)))
    return out


@pytest.mark.parametrize("k,s,t,l,op", [(19, 8, 0, 3, False), (19, 8, 0, 1, False), (15, 8, 0, 0, False), (21, 10, 1, 2, True)])
def test_place_min_seed_quality_filters_syncmers_by_average_phred(k, s, t, l, op, tmp_path):
    """--min-seed-quality (placement.cpp:1179-1240, 1388-1533): only syncmers whose k bases average >= Q (and start inside the trimmed
    range) count, k-min-mers need l consecutive passing syncmers, dedup is ignored; also through the C++ shim from a FASTQ file"""
    rng = np.random.default_rng(60 + k + l)
    idx, _, _ = H.synthetic_index(300, rng, k=k, s=s, t=t, l=l)
    idx.open = int(op)
    reads = H.random_reads(rng, 400, lo=10, hi=220)
    reads = reads + reads[:50] + [b"", b"ACGT", bytes([200, 65, 67]) * 30]
    quals = _quality_strings(rng, reads)
    quals[-1] = bytes([250, 33, 126]) * 30          # bytes >= 128 are negative for the reference's (signed) char
    buf, off = pm.pack_reads(reads)
    qbuf, _ = pm.pack_reads(quals)
    host = pm.HostIndex(idx.hash, idx.parent, idx.child, idx.offsets, idx.parent_index, idx.k, idx.s, idx.t, idx.l, idx.open)
    ws = pm.Workspace(pm.Index(host))
    plain = cpu.seed_table(buf, off, k, s, t, l, op)[1].sum()
    for kw in [dict(min_seed_quality=20), dict(min_seed_quality=13, trim_start=5, trim_end=9), dict(min_seed_quality=31, dedup_reads=1)]:
        res = ws.place_quality(buf, qbuf, off, pm.PlaceParams(**kw))
        okw = dict(trim_start=kw.get("trim_start", 0), trim_end=kw.get("trim_end", 0), min_seed_quality=kw["min_seed_quality"])
        exp = cpu.place(buf, off, idx, want_scores=True, quals=qbuf, **okw)
        assert res.raw.unique_seeds == exp["unique_seeds"] and res.raw.read_unique_seed_count == exp["kept"]
        assert res.raw.total_read_seed_frequency == exp["total_frequency"] and res.raw.min_read_support == exp["min_support"]
        th, tc = ws.seed_table()
        eh, ec = cpu.seed_table(buf, off, k, s, t, l, op, quals=qbuf, **okw)
        assert np.array_equal(th[tc > 0], eh) and np.array_equal(tc[tc > 0], ec)
        assert 0 < ec.sum() < plain
        assert H.relerr(ws.node_scores(), exp["scores"]).max() < RTOL
        for m, name in enumerate(pm.METRICS):
            assert res.best_index[name] == exp["best_index"][m] and np.array_equal(res.tied[name], exp["tied"][m]), name
    # min_seed_quality without qualities must fail loudly, and 0 must ignore them
    with pytest.raises(pm.PanmapError):
        ws.place(buf, off, pm.PlaceParams(min_seed_quality=20))
    assert ws.place_quality(buf, qbuf, off, pm.PlaceParams()).raw.total_read_seed_frequency == plain
    # files through the shim: FASTQ with these qualities, and a FASTA (the reference substitutes 'I' = Q40)
    keep = [i for i, r in enumerate(reads) if r and max(r) < 128]
    fq = tmp_path / "r.fastq"
    with open(fq, "wb") as f:
        for i in keep:
            f.write(b"@r%d\n%s\n+\n%s\n" % (i, reads[i], quals[i].replace(bytes([250]), b"!")))
    fbuf, foff = pm.pack_reads([reads[i] for i in keep]); fq_q, _ = pm.pack_reads([quals[i].replace(bytes([250]), b"!") for i in keep])
    r1 = pm.place_files(ws, str(fq), "", str(tmp_path / "o.tsv"), pm.PlaceParams(min_seed_quality=20))
    e1 = cpu.place(fbuf, foff, idx, quals=fq_q, min_seed_quality=20)
    assert r1.read_unique_seed_count == e1["kept"] and r1.total_read_seed_frequency == e1["total_frequency"]
    assert list(r1.best_index) == list(e1["best_index"])
    fa = tmp_path / "r.fa"
    with open(fa, "wb") as f:
        for i in keep:
            f.write(b">r%d\n%s\n" % (i, reads[i]))
    r2 = pm.place_files(ws, str(fa), "", str(tmp_path / "o2.tsv"), pm.PlaceParams(min_seed_quality=40))
    e2 = cpu.place(fbuf, foff, idx, quals=np.full(fbuf.size, ord("I"), np.uint8), min_seed_quality=40)
    assert r2.total_read_seed_frequency == e2["total_frequency"] > 0 and list(r2.best_index) == list(e2["best_index"])


def test_place_min_seed_quality_on_an_hpc_index_compresses_qualities_in_lockstep():
    """hpc index + --min-seed-quality (placement.cpp:1147-1159): the quality of the first base of every run stays with it"""
    rng = np.random.default_rng(77)
    idx, _, _ = H.synthetic_index(300, rng)
    raw = []
    for r in H.random_reads(rng, 300, lo=20, hi=200):
        a = bytearray()
        for c in r:
            a += bytes([c]) * int(rng.choice([1, 1, 1, 2, 3, 5]))
        raw.append(bytes(a))
    raw += [b"", b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA", b"aAcCgGtTnN" * 12]
    quals = _quality_strings(rng, raw)
    comp, cq = [], []
    for r, q in zip(raw, quals):
        keep = [i for i in range(len(r)) if i == 0 or bytes([r[i]]).upper() != bytes([r[i - 1]]).upper()]
        comp.append(bytes(r[i] for i in keep)); cq.append(bytes(q[i] for i in keep))
    assert comp == [cpu.hpc_compress(r) for r in raw]
    buf, off = pm.pack_reads(raw); qbuf, _ = pm.pack_reads(quals)
    host = pm.HostIndex(idx.hash, idx.parent, idx.child, idx.offsets, idx.parent_index, idx.k, idx.s, idx.t, idx.l, idx.open, hpc=1)
    ws = pm.Workspace(pm.Index(host))
    res = ws.place_quality(buf, qbuf, off, pm.PlaceParams(min_seed_quality=18))
    cbuf, coff = pm.pack_reads(comp); cqb, _ = pm.pack_reads(cq)
    exp = cpu.place(cbuf, coff, idx, want_scores=True, quals=cqb, min_seed_quality=18)
    assert res.raw.unique_seeds == exp["unique_seeds"] and res.raw.read_unique_seed_count == exp["kept"]
    assert res.raw.total_read_seed_frequency == exp["total_frequency"] > 0
    th, tc = ws.seed_table()
    eh, ec = cpu.seed_table(cbuf, coff, idx.k, idx.s, idx.t, idx.l, quals=cqb, min_seed_quality=18)
    assert np.array_equal(th[tc > 0], eh) and np.array_equal(tc[tc > 0], ec)
    assert H.relerr(ws.node_scores(), exp["scores"]).max() < RTOL


def test_place_hpc_index_compresses_reads_on_the_device():
    """index built with --hpc (placement.cpp:1145-1165): the reads are homopolymer-compressed before seeding; also with --dedup,
    which then compares the compressed strings"""
    rng = np.random.default_rng(12)
    idx, _, _ = H.synthetic_index(300, rng)
    raw = []
    for r in H.random_reads(rng, 400, lo=20, hi=200):
        a = bytearray()
        for c in r:                      # stretch runs so that compression changes most reads
            a += bytes([c]) * int(rng.choice([1, 1, 1, 2, 3, 5]))
        raw.append(bytes(a))
    raw += [b"", b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA", b"aAcCgGtTnN" * 12, raw[3], raw[3].lower()]
    comp = [cpu.hpc_compress(r) for r in raw]
    for dedup in (0, 1):
        buf, off = pm.pack_reads(raw)
        host = pm.HostIndex(idx.hash, idx.parent, idx.child, idx.offsets, idx.parent_index, idx.k, idx.s, idx.t, idx.l, idx.open, hpc=1)
        ws = pm.Workspace(pm.Index(host))
        res = ws.place(buf, off, pm.PlaceParams(dedup_reads=dedup))
        cbuf, coff = pm.pack_reads(comp)
        exp = cpu.place(cbuf, coff, idx, want_scores=True, dedup=bool(dedup))
        assert res.raw.unique_seeds == exp["unique_seeds"] and res.raw.read_unique_seed_count == exp["kept"]
        assert res.raw.total_read_seed_frequency == exp["total_frequency"]
        th, tc = ws.seed_table()
        eh, ec = cpu.seed_table(cbuf, coff, idx.k, idx.s, idx.t, idx.l, idx.open, 0, 0, bool(dedup))
        assert np.array_equal(th[tc > 0], eh) and np.array_equal(tc[tc > 0], ec)
        assert H.relerr(ws.node_scores(), exp["scores"]).max() < RTOL
        for m, name in enumerate(pm.METRICS):
            assert res.best_index[name] == exp["best_index"][m] and np.array_equal(res.tied[name], exp["tied"][m])
        res2 = ws.place(buf, off, pm.PlaceParams(dedup_reads=dedup))      # in-place compression is idempotent across calls
        assert res2.raw.unique_seeds == res.raw.unique_seeds and res2.best_index == res.best_index


def test_hpc_index_resident_reads_are_compressed_once_per_upload():
    """hpc_compress works in place and is NOT idempotent: repeated pm_place_resident calls (and table-growth retries) on the same
    upload must reuse the compressed bytes, with and without qualities"""
    rng = np.random.default_rng(21)
    idx, _, _ = H.synthetic_index(200, rng)
    raw = []
    for r in H.random_reads(rng, 500, lo=20, hi=200):
        a = bytearray()
        for c in r:
            a += bytes([c]) * int(rng.choice([1, 1, 2, 2, 3]))
        raw.append(bytes(a))
    comp = [cpu.hpc_compress(r) for r in raw]
    buf, off = pm.pack_reads(raw)
    cbuf, coff = pm.pack_reads(comp)
    host = pm.HostIndex(idx.hash, idx.parent, idx.child, idx.offsets, idx.parent_index, idx.k, idx.s, idx.t, idx.l, idx.open, hpc=1)
    ws = pm.Workspace(pm.Index(host))
    exp = cpu.place(cbuf, coff, idx, want_scores=True)
    eh, ec = cpu.seed_table(cbuf, coff, idx.k, idx.s, idx.t, idx.l, idx.open, 0, 0, False)
    ws.upload(buf, off)
    for _ in range(3):
        res = ws.place_resident()
        assert res.raw.unique_seeds == exp["unique_seeds"] and res.raw.total_read_seed_frequency == exp["total_frequency"]
        th, tc = ws.seed_table()
        assert np.array_equal(th[tc > 0], eh) and np.array_equal(tc[tc > 0], ec)
        assert H.relerr(ws.node_scores(), exp["scores"]).max() < RTOL


def test_workspace_grows_its_table_for_a_much_larger_sample_and_shrinks_back():
    """the count table is fitted to the previous sample: a sample with >10x the unique seeds must trigger the grow-and-redo path and
    succeed (not shrink back on the retry), in either order, also on an hpc index where the retry must not compress twice"""
    rng = np.random.default_rng(22)
    idx, _, _ = H.synthetic_index(300, rng)
    small = H.random_reads(rng, 40)
    big = H.random_reads(rng, 9000, lo=100, hi=200)        # ~9000 * 40 unique random seeds >> 2^16 slots
    for hpc in (0, 1):
        host = pm.HostIndex(idx.hash, idx.parent, idx.child, idx.offsets, idx.parent_index, idx.k, idx.s, idx.t, idx.l, idx.open, hpc=hpc)
        ws = pm.Workspace(pm.Index(host))
        for reads in (small, big, small, big, big, small):
            rr = [cpu.hpc_compress(r) for r in reads] if hpc else reads
            buf, off = pm.pack_reads(reads)
            res = ws.place(buf, off)
            eh, ec = cpu.seed_table(*pm.pack_reads(rr), idx.k, idx.s, idx.t, idx.l, idx.open, 0, 0, False)
            assert res.raw.unique_seeds == eh.size, (hpc, len(reads))
            th, tc = ws.seed_table()
            assert np.array_equal(th[tc > 0], eh) and np.array_equal(tc[tc > 0], ec)
        # the same through the resident path (retries re-run the seeding stage on the uploaded reads)
        ws2 = pm.Workspace(ws.index)
        for reads in (small, big):
            rr = [cpu.hpc_compress(r) for r in reads] if hpc else reads
            ws2.upload(*pm.pack_reads(reads))
            res = ws2.place_resident()
            eh, ec = cpu.seed_table(*pm.pack_reads(rr), idx.k, idx.s, idx.t, idx.l, idx.open, 0, 0, False)
            th, tc = ws2.seed_table()
            assert res.raw.unique_seeds == eh.size and np.array_equal(th[tc > 0], eh) and np.array_equal(tc[tc > 0], ec)


@pytest.mark.parametrize("l", [1, 0])
def test_homopolymer_kmers_are_erased_from_the_read_table(l):
    """placement.cpp:41-76, 1708-1722: the canonical hashes of the four homopolymer k-mers leave the table before anything is counted;
    at l <= 1 seeds ARE syncmer hashes, so poly-A/C/G/T reads put exactly those keys into the table"""
    rng = np.random.default_rng(23 + l)
    idx, _, _ = H.synthetic_index(200, rng, k=19, s=8, t=0, l=l)
    reads = H.random_reads(rng, 200) + [b"A" * 70, b"C" * 45, b"G" * 19, b"T" * 150, b"a" * 33, b"ACGT" * 10 + b"A" * 40 + b"CCGT" * 9] * 3
    res, exp, ws, _ = _place_and_check(idx, reads)
    homo = cpu.seed_table(*pm.pack_reads([b"A" * 19, b"C" * 19]), 19, 8, 0, l, 0, 0, 0, False)
    assert homo[0].size == 0                               # the oracle's table has them erased ...
    th, tc = ws.seed_table()
    raw = set(int(x) for x in cpu.read_seeds(b"A" * 19, 19, 8, 0, l, False, 0, 0)) | set(int(x) for x in cpu.read_seeds(b"C" * 19, 19, 8, 0, l, False, 0, 0))
    assert len(raw) == 2 and not (raw & set(int(x) for x in th[tc > 0]))     # ... and so has the device table (the keys were inserted, then zeroed)


@pytest.mark.parametrize("k,s,t,l,op,ts,te", [(19, 8, 0, 3, False, 0, 0), (19, 8, 0, 3, False, 6, 9), (15, 8, 0, 1, False, 0, 0), (21, 10, 1, 2, True, 0, 0)])
def test_place_packed_reads_equal_ascii_reads(k, s, t, l, op, ts, te):
    """pm_place_packed: the reads arrive as 4-bit codes (half the PCIe bytes, no ASCII on the device); same table, scores and placement as
    pm_place on the ASCII bytes, for the specialised and the generic syncmer kernels, small (one slice) and sliced (>= 65536 reads) samples"""
    rng = np.random.default_rng(300 + k + l + ts)
    idx, _, _ = H.synthetic_index(300, rng, k=k, s=s, t=t, l=l)
    idx.open = int(op)
    host = pm.HostIndex(idx.hash, idx.parent, idx.child, idx.offsets, idx.parent_index, k, s, t, l, int(op))
    ws = pm.Workspace(pm.Index(host))
    prm = pm.PlaceParams(trim_start=ts, trim_end=te)
    base = H.random_reads(rng, 500, lo=0, hi=260, p_n=0.02, p_lower=0.05)
    for reads in (base, [base[i % 500] for i in range(70000)]):
        buf, off = pm.pack_reads(reads)
        if len(reads) > 500:
            os.environ["PM_SLICE_MIN_BYTES"] = "1000000"  # the sliced pipeline (normally samples >= 48 MB)
        try:
            a = ws.place(buf, off, prm)
            ta = ws.seed_table()
            b = ws.place_packed(pm.host_pack_reads(buf, off), off, prm)
            tb = ws.seed_table()
        finally:
            os.environ.pop("PM_SLICE_MIN_BYTES", None)
        assert np.array_equal(ta[0], tb[0]) and np.array_equal(ta[1], tb[1])
        assert a.raw.unique_seeds == b.raw.unique_seeds and a.raw.read_magnitude == b.raw.read_magnitude
        assert all(a.best_index[m] == b.best_index[m] and a.best_score[m] == b.best_score[m] and np.array_equal(a.tied[m], b.tied[m]) for m in pm.METRICS)
    exp = cpu.place(*pm.pack_reads(base), idx, trim_start=ts, trim_end=te)
    got = ws.place_packed(pm.host_pack_reads(*pm.pack_reads(base)), pm.pack_reads(base)[1], prm)
    assert got.raw.unique_seeds == exp["unique_seeds"] and all(got.best_index[n] == exp["best_index"][m] for m, n in enumerate(pm.METRICS))
    with pytest.raises(pm.PanmapError) as e:
        ws.place_packed(pm.host_pack_reads(*pm.pack_reads(base)), pm.pack_reads(base)[1], pm.PlaceParams(dedup_reads=1))
    assert e.value.code == -5


def test_place_empty_and_short_reads():
    rng = np.random.default_rng(6)
    idx, _, _ = H.synthetic_index(50, rng)
    _place_and_check(idx, [b"ACGT", b"", b"NNNNNNNNNNNNNNNNNNNNNNNNN"])
    _place_and_check(idx, [])


@pytest.mark.skipif(not os.path.exists(H.SARS_IDX), reason="reference-built sars_20000 index not staged")
def test_place_sars20000_isolate_matches_reference_golden_tsv():
    """BASELINE config 1: the reference's only numeric golden for the path (examples/expected/single_sample/
    isolate.placement.tsv), reproduced byte for byte through the C ABI."""
    host = pm.HostIndex.read(H.SARS_IDX)
    reads = H.isolate_reads()
    buf, off = pm.pack_reads(reads)
    index = pm.Index(host)
    ws = pm.Workspace(index)
    res = ws.place(buf, off)
    with open(H.ISOLATE_TSV) as f:
        assert res.tsv() == f.read()
    exp = cpu.place(buf, off, host, want_scores=True)
    assert res.raw.read_unique_seed_count == exp["kept"] == 117645
    assert res.raw.unique_seeds == exp["unique_seeds"] == 317148
    sc = ws.node_scores()
    assert H.relerr(sc, exp["scores"]).max() < RTOL
    for m, name in enumerate(pm.METRICS):
        assert res.best_index[name] == exp["best_index"][m]
        assert np.array_equal(res.tied[name], exp["tied"][m])
    met = ws.node_metrics()
    gold = np.load(os.path.join(H.GOLDEN, "sars_isolate_node_metrics_sample.npz"))
    assert np.array_equal(met[gold["nodes"], 2], gold["metrics"][:, 2])          # presence: exact
    assert H.relerr(met[gold["nodes"]], gold["metrics"][:, :5]).max() < RTOL


@pytest.mark.skipif(not os.path.exists(H.SARS_IDX), reason="reference-built sars_20000 index not staged")
def test_place_options_force_leaf_skip_node_min_support():
    host = pm.HostIndex.read(H.SARS_IDX)
    reads = H.isolate_reads()[:6000]
    buf, off = pm.pack_reads(reads)
    index = pm.Index(host)
    ws = pm.Workspace(index)
    for kw, okw in [(dict(force_leaf=1), dict(force_leaf=True)), (dict(skip_node_index=15189), dict(skip_node=15189)),
                    (dict(min_read_support=1), dict(min_read_support=1)), (dict(min_read_support=3, trim_start=5, trim_end=9), dict(min_read_support=3, trim_start=5, trim_end=9))]:
        res = ws.place(buf, off, pm.PlaceParams(**kw))
        exp = cpu.place(buf, off, host, **okw)
        for m, name in enumerate(pm.METRICS):
            assert res.best_index[name] == exp["best_index"][m], (kw, name)
            assert np.array_equal(res.tied[name], exp["tied"][m]), (kw, name)
            assert H.relerr(res.best_score[name], exp["best_score"][m]).max() < RTOL


def test_sharded_scoring_is_identical_to_single_shard():
    """node-sharded runs (multi-GPU layout) on one device: scores of every shard equal the 1-shard scores bit for bit."""
    rng = np.random.default_rng(9)
    idx, hashes, genomes = H.synthetic_index(3000, rng, big_node=(40, 900))
    host = pm.HostIndex(idx.hash, idx.parent, idx.child, idx.offsets, idx.parent_index, 19, 8, 0, 3)
    reads = H.random_reads(rng, 200)
    buf, off = pm.pack_reads(reads)
    # a read table that hits the index: import (hash,count) pairs directly through the staged API
    leaf = 2999
    th = np.array([hashes[j] for j in genomes[leaf].keys()], np.uint64)
    tc = rng.integers(1, 50, size=th.size).astype(np.int64)
    full = None
    for n_sh in (1, 2, 5):
        scores = np.zeros((3000, 5))
        for sh in range(n_sh):
            index = pm.Index(host, shard=sh, n_shards=n_sh)
            ws = pm.Workspace(index)
            p = pm.PlaceParams()
            ws.stage_seed(buf, off, p)
            ws.stage_table_import(th, tc)
            ws.stage_score(p)
            b, e = index.shard_range()
            recs = ws.stage_records()
            r = ws.stage_select(recs, len(reads))
            scores[b:e] = ws.node_scores()[b:e]
        if full is None:
            full = scores
        else:
            assert np.array_equal(scores, full)


def test_reference_unit_test_cases_on_the_gpu():
    """the hand-derived cases of the reference's own unit tests, through the staged ABI (one-node index holding the seed changes, read
    table imported as (hash, count) pairs): test_placement.cpp:100-180 (computeChildMetrics) and :243-295 (min support, magnitudes)"""
    def run(changes, th, tc, min_support=-1):
        idx = H.FlatIdx([c[0] for c in changes], [c[1] for c in changes], [c[2] for c in changes], [0, len(changes)], [0], 15, 8, 0, 1)
        host = pm.HostIndex(idx.hash, idx.parent, idx.child, idx.offsets, idx.parent_index, 15, 8, 0, 1, 0)
        ws = pm.Workspace(pm.Index(host))
        p = pm.PlaceParams(min_read_support=min_support)
        buf, off = pm.pack_reads([b"ACGT"])                   # shorter than k: an empty table to import into
        ws.stage_seed(buf, off, p)
        ws.stage_table_import(np.array(th, np.uint64), np.array(tc, np.int64))
        ws.stage_score(p)
        r = ws.stage_select(ws.stage_records(), 1)
        return ws.node_scores()[0], r
    Hh, X = 0xAAAA, 0xBBBB
    # r=3, g=2: logRaw = (l/2)/l, logCosine 1, containment 1, logContainment 1; weighted = (1/2) / (root's own 1/2)
    s, r = run([(Hh, 0, 2)], [Hh], [3])
    assert np.allclose(s, [0.5, 1.0, 1.0, 1.0, 1.0], rtol=1e-12) and r.best_index["log_raw"] == 0
    assert r.raw.min_read_support == 1 and r.raw.read_unique_seed_count == 1 and r.raw.total_read_seed_frequency == 3
    assert abs(r.raw.read_magnitude - np.log1p(3.0)) < 1e-15
    s, r = run([(X, 0, 5)], [Hh], [3])                      # a genome seed that no read has
    assert np.all(s == 0.0)
    s, r = run([(Hh, 2, 2)], [Hh], [3])                     # unchanged count: no contribution
    assert np.all(s == 0.0)
    H1, H2 = 0x1111, 0x2222                                 # test_placement.cpp:144-180
    s, r = run([(H1, 0, 2), (H2, 0, 2)], [H1, H2], [3, 3])
    assert np.allclose(s, [1 / np.sqrt(2), 1.0, 1.0, 1.0, 1.0], rtol=1e-12)
    s, r = run([(H1, 0, 2)], [H1, H2], [3, 3])
    assert np.allclose(s, [0.5 / np.sqrt(2), 1 / np.sqrt(2), 0.5, 1.0, 0.5], rtol=1e-12)
    # test_placement.cpp:243-295
    for counts, cfg, want in [([5, 4, 3], -1, 2), ([2, 1, 1], -1, 1), ([1], -1, 1), ([5], 7, 7)]:
        _, r = run([(Hh, 0, 1)], [0x10 + i for i in range(len(counts))], counts, cfg)
        assert r.raw.min_read_support == want, counts
    _, r = run([(Hh, 0, 1)], [0x10, 0x11, 0x12], [5, 3, 1], 2)
    la, lb = np.log1p(5.0), np.log1p(3.0)
    assert r.raw.read_unique_seed_count == 2 and r.raw.total_read_seed_frequency == 9
    assert np.isclose(r.raw.log_containment_denominator, la + lb, rtol=1e-14) and np.isclose(r.raw.read_magnitude, np.sqrt(la * la + lb * lb), rtol=1e-14)
    _, r = run([(Hh, 0, 1)], [0x10, 0x11, 0x12], [5, 3, 1], 1)
    assert r.raw.read_unique_seed_count == 3


def test_many_empty_reads_between_real_ones():
    """runs of zero-length reads (more than a pack block can index locally) must not disturb the chunk -> read mapping"""
    rng = np.random.default_rng(21)
    reads = []
    for r in H.random_reads(rng, 40, lo=40, hi=150):
        reads.append(r)
        reads.extend([b""] * int(rng.integers(0, 700)))
    got = pm.read_seeds(reads, 19, 8, 0, 3)
    for r, g in zip(reads, got):
        assert np.array_equal(g, cpu.read_seeds(r, 19, 8, 0, 3))


@pytest.mark.skipif(not os.path.exists(H.SARS_IDX), reason="reference-built sars_20000 index not staged")
def test_place_files_through_cpp_shim_writes_the_golden_tsv(tmp_path):
    """placement::placeLite mirror (panmap_b200/host/placement.cpp): .fastq.gz files in, <prefix>.placement.tsv out"""
    host = pm.HostIndex.read(H.SARS_IDX)
    ws = pm.Workspace(pm.Index(host))
    out = str(tmp_path / "isolate.placement.tsv")
    res = pm.place_files(ws, H.ISOLATE_R1, H.ISOLATE_R2, out)
    assert open(out).read() == open(H.ISOLATE_TSV).read()
    assert res.best_index[4] == 15189 and res.total_reads == 102338
    # whole genome as ONE long read (the reference's e2e test feeds a FASTA, run_e2e.sh:50-56)
    res = pm.place_files(ws, os.path.join(H.REF_DATA, "MZ515733.1.fa"), "", str(tmp_path / "g.tsv"))
    buf, off = pm.read_fastx(os.path.join(H.REF_DATA, "MZ515733.1.fa"))
    exp = cpu.place(buf, off, host)
    assert list(res.best_index) == list(exp["best_index"])


@pytest.mark.skipif(not os.path.exists(H.RSV_IDX), reason="reference-built rsv_4K index not staged")
def test_rsv4k_fastq_places_on_its_genome_like_the_reference_e2e(tmp_path):
    """src/test/e2e/run_e2e.sh:93-99: MZ515733.1.fastq -> node MZ515733.1.  The root of rsv_4K has no seeds, so the
    weighted-containment denominator is 0 and that row has score 0 and an empty node field (golden TSV from the reference)."""
    host = pm.HostIndex.read(H.RSV_IDX)
    ws = pm.Workspace(pm.Index(host))
    out = str(tmp_path / "rsv.tsv")
    res = pm.place_files(ws, os.path.join(H.REF_DATA, "MZ515733.1.fastq"), "", out)
    assert open(out).read() == open(os.path.join(H.GOLDEN, "rsv4k_MZ515733.placement.tsv")).read()
    assert host.node_ids[res.best_index[0]] == "MZ515733.1" and res.best_index[3] == 0xFFFFFFFF
    buf, off = pm.read_fastx(os.path.join(H.REF_DATA, "MZ515733.1.fastq"))
    exp = cpu.place(buf, off, host, want_scores=True)
    r = ws.place(buf, off)
    assert H.relerr(ws.node_scores(), exp["scores"]).max() < RTOL
    for m, name in enumerate(pm.METRICS):
        assert r.best_index[name] == exp["best_index"][m] and np.array_equal(r.tied[name], exp["tied"][m])


@pytest.mark.skipif(not os.path.exists(H.MAMMOTH_IDX), reason="reference-built extended_mammoth index not staged")
def test_mammoth_mtdna_k15_s8_l1_matches_oracle():
    """BASELINE config 2 stand-in (v_mtdna inputs are absent): a real mtDNA PanMAN indexed by the reference at k=15 s=8 l=1 and
    reads simulated from the reference's own reconstruction of node_5 (tests/golden/mammoth_node5_reads.fa.gz)."""
    host = pm.HostIndex.read(H.MAMMOTH_IDX)
    assert (host.k, host.s, host.l) == (15, 8, 1)
    buf, off = pm.read_fastx(os.path.join(H.GOLDEN, "mammoth_node5_reads.fa.gz"))
    ws = pm.Workspace(pm.Index(host))
    res = ws.place(buf, off)
    exp = cpu.place(buf, off, host, want_scores=True)
    assert res.raw.unique_seeds == exp["unique_seeds"] and res.raw.read_unique_seed_count == exp["kept"]
    th, tc = ws.seed_table()
    eh, ec = cpu.seed_table(buf, off, 15, 8, 0, 1)
    assert np.array_equal(th[tc > 0], eh) and np.array_equal(tc[tc > 0], ec)
    assert H.relerr(ws.node_scores(), exp["scores"]).max() < RTOL
    for m, name in enumerate(pm.METRICS):
        assert res.best_index[name] == exp["best_index"][m] and np.array_equal(res.tied[name], exp["tied"][m])
    assert "node_5" in [host.node_ids[int(i)] for i in res.tied["log_containment"]] or host.node_ids[res.best_index["log_containment"]] == "node_5"


def test_batch_mode_concurrent_workspaces_share_one_index():
    """the reference's batch mode places samples from TBB worker threads on one shared tree (main.cpp:1581-1653): here one
    pm_index, one pm_workspace (stream) per thread, results identical to serial placement"""
    import threading
    from tools.synth import synth
    S = synth.generate(3000, 8000, 1.5, 6000, seed=8)
    host = pm.HostIndex(S.hash, S.parent, S.child, S.offsets, S.parent_index, S.k, S.s, S.t, S.l)
    index = pm.Index(host)
    n = 6000
    samples = []
    for j in range(4):
        lo, hi = j * 1500, (j + 1) * 1500
        samples.append((S.reads[int(S.read_offsets[lo]):int(S.read_offsets[hi])].copy(), (S.read_offsets[lo:hi + 1] - S.read_offsets[lo]).copy()))
    serial = []
    ws0 = pm.Workspace(index)
    for b, o in samples:
        serial.append(ws0.place(b, o))
    out = [None] * 4

    def work(j):
        ws = pm.Workspace(index)
        for _ in range(5):
            out[j] = ws.place(*samples[j])
    ts = [threading.Thread(target=work, args=(j,)) for j in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for j in range(4):
        exp = cpu.place(samples[j][0], samples[j][1], S)
        for m, name in enumerate(pm.METRICS):
            assert out[j].best_index[name] == serial[j].best_index[name] == exp["best_index"][m]
            assert out[j].best_score[name] == serial[j].best_score[name]
            assert np.array_equal(out[j].tied[name], exp["tied"][m])


def test_full_size_sample_matches_oracle_and_is_shard_and_call_invariant():
    """BASELINE configs[2] at full size (1M-node tree, 15.6M deltas, 1M x 150 bp reads; the workload bench.py times): the whole
    result against the CPU oracle -- integers, table, best nodes and tie lists bit-exact, all 5M scores to 1e-12 -- plus the
    size-independent properties: a node without deltas scores exactly like its parent, the truth leaf is the placement, a second
    call and a 3-shard index give bit-identical scores."""
    from tools.synth import synth
    S = synth.generate(1_000_000, 30_000, 1.0, 1_000_000, read_len=150, seed=0)
    host = pm.HostIndex(S.hash, S.parent, S.child, S.offsets, S.parent_index, S.k, S.s, S.t, S.l)
    index = pm.Index(host)
    ws = pm.Workspace(index)
    res = ws.place(S.reads, S.read_offsets)
    sc = ws.node_scores()
    exp = cpu.place(S.reads, S.read_offsets, S, want_scores=True)
    assert res.raw.unique_seeds == exp["unique_seeds"] and res.raw.read_unique_seed_count == exp["kept"]
    assert res.raw.total_read_seed_frequency == exp["total_frequency"] and res.raw.min_read_support == exp["min_support"]
    assert H.relerr(sc, exp["scores"]).max() < RTOL
    for m, name in enumerate(pm.METRICS):
        assert res.best_index[name] == exp["best_index"][m], name
        assert np.array_equal(res.tied[name], exp["tied"][m]), name
        assert H.relerr(res.best_score[name], exp["best_score"][m]).max() < RTOL
    assert S.truth in res.tied["log_raw"] and S.truth in res.tied["log_containment"]
    # exact prefix: nodes without deltas repeat their parent's scores bit for bit
    nd = np.diff(S.offsets.astype(np.int64))
    z = np.nonzero(nd[1:] == 0)[0] + 1
    assert z.size > 1000 and np.array_equal(sc[z], sc[S.parent_index[z]])
    # determinism across calls (atomics order must not matter) and across shardings of the index
    ws.place(S.reads, S.read_offsets)
    assert np.array_equal(ws.node_scores(), sc)
    del ws, index
    got = np.zeros_like(sc)
    for sh in range(3):
        ix = pm.Index(host, shard=sh, n_shards=3)
        w2 = pm.Workspace(ix)
        w2.place(S.reads, S.read_offsets)
        b, e = ix.shard_range()
        got[b:e] = w2.node_scores()[b:e]
        del w2, ix
    assert np.array_equal(got, sc)


def test_bacterial_shape_reduced_matches_oracle():
    """BASELINE configs[3] shape at reduced size (the oracle needs minutes at 100k nodes x 10M reads): a 600 kb genome, lambda = 50 mutations
    per edge, so a root of ~190 k deltas spanning hundreds of delta chunks, millions of distinct read seeds (mostly sequencing-error seeds
    seen once), the hot-id shared-memory copy covering a sliver of the seed ids.  Whole result against the oracle, as for configs[2]."""
    from tools.synth import synth
    S = synth.generate(4000, 600_000, 50.0, 300_000, read_len=150, seed=3)
    host = pm.HostIndex(S.hash, S.parent, S.child, S.offsets, S.parent_index, S.k, S.s, S.t, S.l)
    ws = pm.Workspace(pm.Index(host))
    res = ws.place(S.reads, S.read_offsets)
    sc = ws.node_scores()
    exp = cpu.place(S.reads, S.read_offsets, S, want_scores=True)
    assert int(S.offsets[1]) > 100_000 and exp["unique_seeds"] > 1_000_000
    assert res.raw.unique_seeds == exp["unique_seeds"] and res.raw.read_unique_seed_count == exp["kept"]
    assert res.raw.total_read_seed_frequency == exp["total_frequency"] and res.raw.min_read_support == exp["min_support"]
    assert H.relerr(sc, exp["scores"]).max() < RTOL
    for m, name in enumerate(pm.METRICS):
        assert res.best_index[name] == exp["best_index"][m], name
        assert np.array_equal(res.tied[name], exp["tied"][m]), name
        assert H.relerr(res.best_score[name], exp["best_score"][m]).max() < RTOL
    assert S.truth in res.tied["log_raw"]
    # resident path and a 2-shard index: bit-identical scores
    ws.upload(S.reads, S.read_offsets)
    ws.place_resident()
    assert np.array_equal(ws.node_scores(), sc)


def test_table_estimate_is_clamped_to_the_device_limit_and_a_real_overflow_fails(monkeypatch):
    """the first table of a workspace is sized from the number of k-mer windows (an upper bound); at bacterial scale that estimate exceeds
    the linear-texture width the table is bound to.  With the limit lowered: a sample whose distinct seeds fit is placed correctly in a
    table clamped to the limit, one whose distinct seeds cannot fit is refused (no silent truncation)."""
    rng = np.random.default_rng(5)
    idx, _, _ = H.synthetic_index(200, rng)
    host = pm.HostIndex(idx.hash, idx.parent, idx.child, idx.offsets, idx.parent_index, idx.k, idx.s, idx.t, idx.l, idx.open)
    monkeypatch.setenv("PM_TABLE_SLOT_LIMIT", str(1 << 16))
    ws = pm.Workspace(pm.Index(host))
    one = H.random_reads(rng, 1, lo=180, hi=181)
    many_copies = one * 9000                               # 9000 * 160 windows / 4 = 360 k > 65,536 slots, but only ~50 distinct seeds
    buf, off = pm.pack_reads(many_copies)
    ws.upload(buf, off)
    res = ws.place_resident()
    eh, ec = cpu.seed_table(buf, off, idx.k, idx.s, idx.t, idx.l, idx.open, 0, 0, False)
    th, tc = ws.seed_table()
    assert res.raw.unique_seeds == eh.size and np.array_equal(th[tc > 0], eh) and np.array_equal(tc[tc > 0], ec)
    big = H.random_reads(rng, 9000, lo=100, hi=200)        # ~300 k distinct random seeds: cannot fit 65,536 slots
    with pytest.raises(pm.PanmapError):
        ws.place(*pm.pack_reads(big))
    res2 = ws.place(buf, off)                              # the workspace stays usable
    assert res2.raw.unique_seeds == eh.size


@pytest.mark.parametrize("k,s,t,l", [(19, 8, 0, 3), (15, 8, 0, 3), (15, 8, 0, 1)])
def test_partitioned_counting_gives_the_same_table(monkeypatch, k, s, t, l):
    """tables larger than L2 are filled by scatter_seeds_lane + count_buckets (seed instances partitioned by table region, counted region by
    region).  With the threshold lowered, a small sample goes that way through every entry point (resident, host buffers in one piece and
    sliced, 4-bit codes) and must leave exactly the oracle's table and placement; with regions of 4 entries nearly every seed takes the
    full-region fallback."""
    rng = np.random.default_rng(31)
    idx, _, _ = H.synthetic_index(200, rng, k=k, s=s, t=t, l=l)
    host = pm.HostIndex(idx.hash, idx.parent, idx.child, idx.offsets, idx.parent_index, idx.k, idx.s, idx.t, idx.l, idx.open)
    reads = H.random_reads(rng, 6000, lo=60, hi=260) + [b"", b"ACGT", b"N" * 70]
    buf, off = pm.pack_reads(reads)
    eh, ec = cpu.seed_table(buf, off, idx.k, idx.s, idx.t, idx.l, idx.open, 0, 0, False)
    ws0 = pm.Workspace(pm.Index(host))
    want = ws0.place(buf, off)
    monkeypatch.setenv("PM_BUCKET_MIN_SLOTS", "4096")
    for region_cap, slice_bytes in ((None, None), (None, "65536"), ("4", None)):
        if region_cap:
            monkeypatch.setenv("PM_BUCKET_REGION_CAP", region_cap)
        if slice_bytes:
            monkeypatch.setenv("PM_SLICE_MIN_BYTES", slice_bytes)
        ws = pm.Workspace(pm.Index(host))
        launches0 = pm.launch_count()
        for how in ("host", "resident", "packed", "host"):
            if how == "host":
                res = ws.place(buf, off)
            elif how == "resident":
                ws.upload(buf, off); res = ws.place_resident()
            else:
                res = ws.place_packed(pm.host_pack_reads(buf, off), off)
            th, tc = ws.seed_table()
            assert res.raw.unique_seeds == eh.size and np.array_equal(th[tc > 0], eh) and np.array_equal(tc[tc > 0], ec), (how, region_cap, slice_bytes)
            for m in pm.METRICS:
                assert res.best_index[m] == want.best_index[m] and res.best_score[m] == want.best_score[m] and np.array_equal(res.tied[m], want.tied[m])
        monkeypatch.delenv("PM_BUCKET_REGION_CAP", raising=False); monkeypatch.delenv("PM_SLICE_MIN_BYTES", raising=False)


def test_hash_seq_matches_oracle_and_rejects_non_acgt():
    """seeding::hashSeq (seeding.cpp:20-30) on the GPU for a batch of k-mers of every length 1..40 (rotations wrap at 64 like the reference's)"""
    rng = np.random.default_rng(77)
    seqs = [bytes(rng.choice(np.frombuffer(b"ACGTacgt", np.uint8), size=n)) for n in list(range(1, 41)) * 3] + [b"A" * 70, b"ACGT" * 33]
    f, r = pm.hash_seq(seqs)
    for s, a, b in zip(seqs, f, r):
        assert (int(a), int(b)) == cpu.hash_seq(s), s
    with pytest.raises(pm.PanmapError, match="non canonical base"):
        pm.hash_seq([b"ACGT", b"ACNT"])
    f, r = pm.hash_seq([])
    assert f.size == 0
