"""GPU parity tests (-m gpu) of ONE sample placed by several ranks together (pm_comm, include/panmap_b200.h): node range sharded,
reads sliced, seed table hash-partitioned, everything enqueued without host round trips.  The in-process transport lets all ranks
sit on one device, so the whole data plane (every kernel, every buffer layout, the overflow / regrow protocol, the long-tie-list
path) is checked against the CPU oracle on a single GPU; tests/test_gpu_nccl.py repeats the comparison over NCCL on >= 2 GPUs.
Integers, best nodes and tie lists are bit-exact, f64 scores agree to 1e-12 -- the same bar as the one-GPU path."""
import os

import numpy as np
import pytest

import panmap_b200 as pm
from oracle import cpu
from tests import helpers as H
from tools.synth import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _group(host, n_ranks, device=0):
    idxs = [pm.Index(host, device=device, shard=r, n_shards=n_ranks) for r in range(n_ranks)]
    wss = [pm.Workspace(i) for i in idxs]
    return pm.Comm.local(wss)


def _check(comms, res, exp, n_nodes):
    assert res.raw.unique_seeds == exp["unique_seeds"]
    assert res.raw.read_unique_seed_count == exp["kept"]
    assert res.raw.total_read_seed_frequency == exp["total_frequency"]
    assert res.raw.min_read_support == exp["min_support"]
    assert H.relerr(res.raw.read_magnitude, exp["magnitude"]).max() < RTOL
    assert H.relerr(res.raw.log_containment_denominator, exp["log_sum"]).max() < RTOL
    assert H.relerr(res.raw.weighted_containment_denominator, exp["wc_denominator"]).max() < RTOL
    if exp.get("scores") is not None:
        sc = np.zeros((n_nodes, 5))
        covered = 0
        for c in comms:                                   # every rank holds the scores of its own node range
            b, e = c.ws.index.shard_range()
            sc[b:e] = c.ws.node_scores()[b:e]
            covered += e - b
        assert covered == n_nodes
        assert H.relerr(sc, exp["scores"]).max() < RTOL
    for m, name in enumerate(pm.METRICS):
        assert H.relerr(res.best_score[name], exp["best_score"][m]).max() < RTOL, name
        assert res.best_index[name] == exp["best_index"][m], name
        assert np.array_equal(res.tied[name], exp["tied"][m]), name
    # every rank ends with the same result
    for c in comms[1:]:
        for m, name in enumerate(pm.METRICS):
            t = np.zeros(len(res.tied[name]) + 1, np.uint32)
            pm.api._ck(pm.lib().pm_get_tied(c.ws._h, m, t.ctypes.data, t.size))
            assert np.array_equal(t[:-1], res.tied[name])


def _host_of(idx, **kw):
    return pm.HostIndex(idx.hash, idx.parent, idx.child, idx.offsets, idx.parent_index, idx.k, idx.s, idx.t, idx.l, getattr(idx, "open", 0), **kw)


@pytest.mark.parametrize("n_ranks", [1, 2, 3, 8])
def test_sharded_sample_equals_oracle_on_genome_derived_index(n_ranks):
    S = synth.generate(2500, 6000, 1.5, 4000, seed=7)
    comms = _group(_host_of(S), n_ranks)
    res = pm.place_multi(comms, S.reads, S.read_offsets)
    exp = cpu.place(S.reads, S.read_offsets, S, want_scores=True)
    _check(comms, res, exp, S.n_nodes)
    assert S.truth in res.tied["log_raw"] and res.raw.total_reads == 4000
    sent, recv = comms[0].last_traffic()
    assert (sent > 0) == (n_ranks > 1)


@pytest.mark.parametrize("k,s,t,l,op", [(19, 8, 0, 3, False), (15, 8, 0, 1, False), (21, 10, 1, 2, True), (12, 5, 2, 0, False)])
def test_sharded_sample_other_seeding_parameters_and_table_only_reads(k, s, t, l, op):
    """random reads against a random index: most read seeds are NOT in the index, so the count-1 seeds that travel as a count only
    (never as entries) decide the magnitudes; index seeds with genome counts >= 2 exercise the general-delta path per shard"""
    rng = np.random.default_rng(200 + k + l)
    idx, _, _ = H.synthetic_index(600, rng, k=k, s=s, t=t, l=l)
    idx.open = int(op)
    reads = H.random_reads(rng, 700, lo=10, hi=220)
    reads = reads + reads[:150] + reads[:30] * 3                     # counts 1, 2 and 5 side by side
    buf, off = pm.pack_reads(reads)
    for n_ranks in (2, 5):
        comms = _group(_host_of(idx), n_ranks)
        for mrs in (-1, 0, 1, 2, 4):
            res = pm.place_multi(comms, buf, off, pm.PlaceParams(min_read_support=mrs))
            exp = cpu.place(buf, off, idx, want_scores=True, min_read_support=mrs)
            _check(comms, res, exp, 600)


def test_sharded_sample_options_and_unsupported_ones():
    S = synth.generate(1500, 5000, 1.2, 2500, seed=9)
    comms = _group(_host_of(S), 3)
    for kw, okw in [(dict(force_leaf=1), dict(force_leaf=True)), (dict(skip_node_index=int(S.truth)), dict(skip_node=int(S.truth))),
                    (dict(trim_start=4, trim_end=11), dict(trim_start=4, trim_end=11))]:
        res = pm.place_multi(comms, S.reads, S.read_offsets, pm.PlaceParams(**kw))
        exp = cpu.place(S.reads, S.read_offsets, S, want_scores=False, **okw)
        _check(comms, res, exp, S.n_nodes)
    for kw in (dict(dedup_reads=1), dict(seed_mask_fraction=0.01), dict(min_seed_quality=20)):
        with pytest.raises(pm.PanmapError) as e:
            pm.place_multi(comms, S.reads, S.read_offsets, pm.PlaceParams(**kw))
        assert e.value.code == -5
    # the group still works after a refused call
    _check(comms, pm.place_multi(comms, S.reads, S.read_offsets), cpu.place(S.reads, S.read_offsets, S), S.n_nodes)


def test_sharded_sample_regrows_its_exchange_buffers_and_tables():
    """capacities follow the previous sample: a much larger one must trip the in-band overflow flags on some rank, reach every
    rank with the last all-gather and be redone with larger buffers -- and a small one afterwards must shrink back and still agree"""
    rng = np.random.default_rng(41)
    idx, _, _ = H.synthetic_index(500, rng)
    small = H.random_reads(rng, 60)
    big = H.random_reads(rng, 12000, lo=100, hi=200)
    comms = _group(_host_of(idx), 4)
    for reads in (small, big, small, big, big):
        buf, off = pm.pack_reads(reads)
        res = pm.place_multi(comms, buf, off)
        _check(comms, res, cpu.place(buf, off, idx), 500)


def test_sharded_sample_long_tie_lists_take_the_second_exchange():
    """hardly any mutations: thousands of nodes carry the root's genome and tie exactly; lists longer than the 64 heads that travel
    with the result need one more all-gather"""
    S = synth.generate(3000, 4000, 0.02, 1500, seed=11)
    comms = _group(_host_of(S), 4)
    res = pm.place_multi(comms, S.reads, S.read_offsets)
    exp = cpu.place(S.reads, S.read_offsets, S, want_scores=True)
    assert max(len(t) for t in exp["tied"]) > 64 * 4
    _check(comms, res, exp, S.n_nodes)


def test_sharded_sample_resident_reads_and_repeat_calls():
    S = synth.generate(2000, 6000, 1.5, 3000, seed=4)
    n_ranks = 4
    comms = _group(_host_of(S), n_ranks)
    from panmap_b200 import distributed as pmd
    for c in comms:
        r, o = pmd.slice_reads(S.reads, S.read_offsets, c.rank, n_ranks)
        c.ws.upload(r, o)
    exp = cpu.place(S.reads, S.read_offsets, S, want_scores=True)
    for _ in range(3):
        _check(comms, pm.place_multi_resident(comms), exp, S.n_nodes)
    # and the plain one-GPU call on a full-range workspace gives the same answer as the group
    ws = pm.Workspace(pm.Index(_host_of(S)))
    one = ws.place(S.reads, S.read_offsets)
    grp = pm.place_multi_resident(comms)
    assert all(one.best_index[m] == grp.best_index[m] and np.array_equal(one.tied[m], grp.tied[m]) and one.best_score[m] == grp.best_score[m] for m in pm.METRICS)
    assert one.raw.read_magnitude == grp.raw.read_magnitude and one.raw.log_containment_denominator == grp.raw.log_containment_denominator


def test_sharded_sample_empty_and_tiny_inputs():
    rng = np.random.default_rng(6)
    idx, _, _ = H.synthetic_index(50, rng)
    comms = _group(_host_of(idx), 4)
    for reads in ([], [b"ACGT", b"", b"NNNNNNNNNNNNNNNNNNNNNNNNN"], H.random_reads(rng, 3)):
        buf, off = pm.pack_reads(reads)
        _check(comms, pm.place_multi(comms, buf, off), cpu.place(buf, off, idx), 50)


@pytest.mark.skipif(not os.path.exists(H.SARS_IDX), reason="reference-built sars_20000 index not staged")
def test_sharded_sars20000_isolate_writes_the_reference_golden_tsv():
    """BASELINE config 1 over 4 ranks: the reference's only numeric golden for the path, byte for byte"""
    host = pm.HostIndex.read(H.SARS_IDX)
    buf, off = pm.pack_reads(H.isolate_reads())
    comms = _group(host, 4)
    res = pm.place_multi(comms, buf, off)
    with open(H.ISOLATE_TSV) as f:
        assert res.tsv() == f.read()
    exp = cpu.place(buf, off, host, want_scores=True)
    _check(comms, res, exp, host.n_nodes)
    assert res.raw.read_unique_seed_count == 117645 and res.raw.unique_seeds == 317148


def test_sharded_sample_across_all_visible_devices():
    """in-process transport with the ranks on different devices (peer copies); runs on whatever the box has"""
    n_dev = pm.device_count()
    if n_dev < 2:
        pytest.skip("one device")
    S = synth.generate(2500, 6000, 1.5, 4000, seed=7)
    host = _host_of(S)
    wss = [pm.Workspace(pm.Index(host, device=r, shard=r, n_shards=n_dev)) for r in range(n_dev)]
    comms = pm.Comm.local(wss)
    _check(comms, pm.place_multi(comms, S.reads, S.read_offsets), cpu.place(S.reads, S.read_offsets, S, want_scores=True), S.n_nodes)
