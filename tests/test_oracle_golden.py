"""The CPU oracle (oracle/panmap_oracle.c) against the committed golden vectors in tests/golden/, which were produced
by the reference's own code (tools/make_golden.py over oracle/_ref/libpanmap_ref.so).  Runs without /root/reference."""
import os

import numpy as np
import pytest

from oracle import cpu
from tests import helpers as H


def test_hash_seq_golden():
    g = np.load(os.path.join(H.GOLDEN, "hash_seq.npz"))
    for seq, f, r in zip(g["seq"], g["f"], g["r"]):
        assert cpu.hash_seq(str(seq)) == (int(f), int(r))
    with pytest.raises(ValueError):
        cpu.hash_seq("ACGNACG")   # reference throws std::invalid_argument (seeding.cpp:25)


def test_hash_seq_strand_invariance():
    # test_seeding.cpp:20-35: forward hash of a k-mer == reverse hash of its reverse complement
    rng = np.random.default_rng(0)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    for k in (15, 19, 31):
        s = "".join(rng.choice(list("ACGT"), size=k))
        rc = "".join(comp[c] for c in reversed(s))
        f, r = cpu.hash_seq(s)
        f2, r2 = cpu.hash_seq(rc)
        assert (f, r) == (r2, f2)


def test_rolling_syncmers_golden():
    g = np.load(os.path.join(H.GOLDEN, "rolling_syncmers.npz"))
    for i in range(int(g["n"])):
        seq = str(g[f"seq_{i}"]); k, s, t, op = int(g[f"k_{i}"]), int(g[f"s_{i}"]), int(g[f"t_{i}"]), bool(g[f"open_{i}"])
        h, rev, syn, pos = cpu.rolling_syncmers(seq, k, s, op, t, True)
        assert np.array_equal(h, g[f"hash_{i}"]) and np.array_equal(rev, g[f"rev_{i}"]) and np.array_equal(syn, g[f"syn_{i}"])
        assert np.array_equal(pos, np.arange(len(seq) - k + 1))
        # test_seeding.cpp:37-82: syncmer-only output == flagged subset; every syncmer hash == min(hashSeq(kmer))
        h2, rev2, syn2, pos2 = cpu.rolling_syncmers(seq, k, s, op, t, False)
        assert np.array_equal(h2, h[syn == 1]) and np.array_equal(pos2, pos[syn == 1])
        for hh, p in zip(h2[:5], pos2[:5]):
            assert int(hh) == min(cpu.hash_seq(seq[int(p):int(p) + k].upper()))


def test_select_chain_golden():
    g = np.load(os.path.join(H.GOLDEN, "select_chain.npz"))
    for i in range(int(g["n"])):
        bs, bi, tied = cpu.select_chain(g[f"order_{i}"], g[f"score_{i}"])
        assert bs == float(g[f"best_{i}"][0]) and bi == int(g[f"idx_{i}"][0]) and np.array_equal(tied, g[f"tied_{i}"])


def test_child_metrics_hand_derived():
    """test_placement.cpp:100-180: r=3,g=2 -> logRaw 0.5, logCosine 1, containment 1, weighted 0.5, logContainment 1; etc."""
    Hh, X = 0xAAAA, 0xBBBB
    th = np.array([Hh], np.uint64); lv = np.array([np.log1p(3.0)])

    def run(changes, th=th, lv=lv, U1=1.0, mag=np.log1p(3.0), denL=np.log1p(3.0), denW=1.0):
        idx = H.FlatIdx([c[0] for c in changes], [c[1] for c in changes], [c[2] for c in changes], [0, len(changes)], [0], 15, 8, 0, 1)
        m, s = cpu.node_metrics(idx, th, lv, U1, mag, denL, denW)
        return m[0], s[0]
    m, s = run([(Hh, 0, 2)])
    assert np.allclose(s, [0.5, 1.0, 1.0, 0.5, 1.0], rtol=1e-12) and m[2] == 1
    m, s = run([(X, 0, 5)])
    assert abs(s[0]) < 1e-12 and m[2] == 0 and m[5] > 0
    m, s = run([(Hh, 2, 2)])
    assert m[0] == 0 and m[5] == 0 and m[2] == 0
    # two seeds at count 3 in the reads, genome counts 2 (test_placement.cpp:144-180)
    H1, H2 = 0x1111, 0x2222
    th2 = np.array([H1, H2], np.uint64); l4 = np.log1p(3.0); lv2 = np.array([l4, l4])
    kw = dict(th=th2, lv=lv2, U1=2.0, mag=np.sqrt(2 * l4 * l4), denL=2 * l4, denW=1.0)
    m, s = run([(H1, 0, 2), (H2, 0, 2)], **kw)
    assert np.allclose(s, [1 / np.sqrt(2), 1.0, 1.0, 1.0, 1.0], rtol=1e-12) and m[2] == 2
    m, s = run([(H1, 0, 2)], **kw)
    assert np.allclose(s, [0.5 / np.sqrt(2), 1 / np.sqrt(2), 0.5, 0.5, 0.5], rtol=1e-12) and m[2] == 1


def test_min_support_and_magnitudes():
    """test_placement.cpp:243-295"""
    assert cpu.resolve_min_read_support([5, 4, 3], -1) == 2
    assert cpu.resolve_min_read_support([2, 1, 1], -1) == 1
    assert cpu.resolve_min_read_support([1], -1) == 1
    assert cpu.resolve_min_read_support([5], 7) == 7
    logv, sc = cpu.read_magnitudes([5, 3, 1], 2)
    la, lb = np.log1p(5.0), np.log1p(3.0)
    assert sc["dropped"] == 1 and sc["kept"] == 2 and sc["total"] == 9
    assert np.isclose(sc["log_sum"], la + lb, rtol=1e-15) and np.isclose(sc["magnitude"], np.sqrt(la * la + lb * lb), rtol=1e-15)
    assert logv[2] == 0.0
    assert cpu.read_magnitudes([5, 3, 1], 1)[1]["kept"] == 3


@pytest.mark.skipif(not os.path.exists(H.SARS_IDX), reason="reference-built sars_20000 index not staged (oracle/_ref/data)")
def test_oracle_reproduces_config1_golden():
    """BASELINE config 1 end to end on the CPU oracle: same TSV as examples/expected/single_sample/isolate.placement.tsv"""
    import panmap_b200 as pm
    host = pm.HostIndex.read(H.SARS_IDX)
    buf, off = pm.pack_reads(H.isolate_reads())
    r = cpu.place(buf, off, host)
    g = np.load(os.path.join(H.GOLDEN, "sars_isolate_summary.npz"))
    assert np.array_equal(r["best_index"], g["best_index"])
    for m in range(5):
        assert np.array_equal(r["tied"][m], g[f"tied_{m}"])
    assert r["kept"] == int(g["kept"]) and r["unique_seeds"] == int(g["unique_seeds"]) and r["total_frequency"] == int(g["total_frequency"])
    assert H.relerr(r["best_score"], g["best_score"]).max() < 1e-12
    lines = ["metric\tscore\tnodes"]
    for m, name in enumerate(pm.METRICS):
        lines.append(f"{name}\t{r['best_score'][m]:.6f}\t" + ",".join(host.node_ids[int(i)] for i in r["tied"][m]))
    assert "\n".join(lines) + "\n" == open(os.path.join(H.GOLDEN, "isolate.placement.tsv")).read()
