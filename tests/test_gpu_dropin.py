"""GPU test (-m gpu) of the COMPILED drop-in: panmap_b200/integration/panmap_adapter.cpp is placement::placeLite with the reference's own
signature (placement.hpp:237-244), built by oracle/ref_build/Makefile against the reference's own headers into oracle/_ref/libpanmap_dropin.so
together with the reference's unmodified translation units.  The test driver (oracle/ref_build/dropin_driver.cpp) runs the reference's caller
sequence (map the .idx -> FlatArrayMessageReader -> LiteTree::initialize -> placeLite, main.cpp:1668-1750) twice in the same process:
through the adapter (GPU) and through the reference's own placeLite (renamed at build time), and every PlacementResult field, the seed
table handed to the alignment stage and the TSV files are compared."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
LIB = os.path.join(H.ROOT, "oracle", "_ref", "libpanmap_dropin.so")


class Out(C.Structure):
    _fields_ = [("best_score", C.c_double * 5), ("best_index", C.c_uint32 * 5), ("tied_count", C.c_int64 * 5), ("total_reads", C.c_int64),
                ("read_unique_seed_count", C.c_uint64), ("total_read_seed_frequency", C.c_int64), ("read_magnitude", C.c_double),
                ("seed_table_size", C.c_int64), ("k", C.c_int32), ("s", C.c_int32), ("t", C.c_int32), ("open", C.c_int32),
                ("node_score_rows", C.c_int64), ("best_id", (C.c_char * 64) * 5)]


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libpanmap_dropin.so not built (needs /root/reference at build time)")
    L = C.CDLL(LIB)
    L.dropin_last_error.restype = C.c_char_p
    L.dropin_open.restype = C.c_void_p
    L.dropin_open.argtypes = [C.c_char_p]
    L.dropin_close.argtypes = [C.c_void_p]
    L.dropin_place.restype = C.c_void_p
    L.dropin_place.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_int, C.c_int, C.POINTER(Out)]
    L.dropin_free.argtypes = [C.c_void_p]
    L.dropin_tied.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.dropin_seed_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.dropin_node_scores.argtypes = [C.c_void_p, C.c_void_p]
    L.dropin_place_concurrent.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p]
    return L


def _place(L, h, which, r1, r2, tsv, **kw):
    o = Out()
    k = L.dropin_place(h, which, os.fsencode(r1), os.fsencode(r2), os.fsencode(tsv), kw.get("threads", 4), kw.get("min_read_support", -1),
                       kw.get("seed_mask_fraction", 0.0), kw.get("trim_start", 0), kw.get("trim_end", 0), int(kw.get("dedup", 0)), int(kw.get("force_leaf", 0)),
                       int(kw.get("store_diag", 0)), kw.get("min_seed_quality", 0), C.byref(o))
    if not k:
        raise RuntimeError(L.dropin_last_error().decode())
    tied = []
    for m in range(5):
        t = np.zeros(max(o.tied_count[m], 1), np.uint32)
        L.dropin_tied(k, m, t.ctypes.data)
        tied.append(t[:o.tied_count[m]].copy())
    hsh = np.zeros(max(o.seed_table_size, 1), np.uint64); cnt = np.zeros(max(o.seed_table_size, 1), np.int64)
    L.dropin_seed_table(k, hsh.ctypes.data, cnt.ctypes.data)
    order = np.argsort(hsh[:o.seed_table_size], kind="stable")
    ns = np.zeros((max(o.node_score_rows, 1), 5), np.float32)
    if o.node_score_rows:
        L.dropin_node_scores(k, ns.ctypes.data)
    L.dropin_free(k)
    return dict(o=o, tied=tied, table=(hsh[:o.seed_table_size][order], cnt[:o.seed_table_size][order]), node_scores=ns[:o.node_score_rows],
                ids=[bytes(o.best_id[m]).split(b"\0")[0].decode() for m in range(5)], tsv=open(tsv).read() if os.path.exists(tsv) else "")


def _same(g, r, scores_rtol=1e-12):
    a, b = g["o"], r["o"]
    for m in range(5):
        assert a.best_index[m] == b.best_index[m] and np.array_equal(g["tied"][m], r["tied"][m]) and g["ids"][m] == r["ids"][m], m
        assert abs(a.best_score[m] - b.best_score[m]) <= scores_rtol * max(abs(b.best_score[m]), 1e-9), m
    assert a.total_reads == b.total_reads and a.read_unique_seed_count == b.read_unique_seed_count
    assert a.total_read_seed_frequency == b.total_read_seed_frequency and a.seed_table_size == b.seed_table_size
    assert abs(a.read_magnitude - b.read_magnitude) <= 1e-12 * b.read_magnitude if b.read_magnitude else a.read_magnitude == 0
    assert (a.k, a.s, a.t, a.open) == (b.k, b.s, b.t, b.open)
    assert np.array_equal(g["table"][0], r["table"][0]) and np.array_equal(g["table"][1], r["table"][1])   # what the alignment stage receives
    assert g["tsv"] == r["tsv"] and g["tsv"].startswith("metric\tscore\tnodes\n")


@pytest.mark.skipif(not os.path.exists(H.SARS_IDX), reason="reference-built sars_20000 index not staged")
def test_dropin_placelite_equals_reference_placelite_on_config1(lib, tmp_path):
    h = lib.dropin_open(os.fsencode(H.SARS_IDX))
    assert h, lib.dropin_last_error()
    try:
        g = _place(lib, h, 1, H.ISOLATE_R1, H.ISOLATE_R2, str(tmp_path / "gpu.tsv"))
        r = _place(lib, h, 0, H.ISOLATE_R1, H.ISOLATE_R2, str(tmp_path / "ref.tsv"))
        _same(g, r)
        assert g["tsv"] == open(H.ISOLATE_TSV).read()                   # the reference's golden file, byte for byte
        assert g["o"].total_reads == 102338 and g["o"].read_unique_seed_count == 117645 and (g["o"].k, g["o"].s, g["o"].t, g["o"].open) == (19, 8, 0, 0)
        # second call on the same tree: the device index is cached (like seedChangesLoaded), the result is the same
        _same(_place(lib, h, 1, H.ISOLATE_R1, H.ISOLATE_R2, str(tmp_path / "gpu2.tsv")), dict(r, tsv=r["tsv"]))
        for kw in (dict(dedup=1), dict(trim_start=5, trim_end=9), dict(force_leaf=1), dict(min_read_support=1), dict(min_read_support=3)):
            _same(_place(lib, h, 1, H.ISOLATE_R1, H.ISOLATE_R2, str(tmp_path / "g.tsv"), **kw), _place(lib, h, 0, H.ISOLATE_R1, H.ISOLATE_R2, str(tmp_path / "r.tsv"), **kw))
        # --dump-all-scores: per-node float scores
        gd = _place(lib, h, 1, H.ISOLATE_R1, H.ISOLATE_R2, str(tmp_path / "gd.tsv"), store_diag=1)
        rd = _place(lib, h, 0, H.ISOLATE_R1, H.ISOLATE_R2, str(tmp_path / "rd.tsv"), store_diag=1)
        assert gd["node_scores"].shape == rd["node_scores"].shape == (39999, 5)
        assert np.allclose(gd["node_scores"], rd["node_scores"], rtol=2e-7, atol=1e-30)
        # the batch caller's shape: TBB workers calling concurrently on one LiteTree (main.cpp:1574-1592)
        best = np.zeros(6, np.uint32)
        assert lib.dropin_place_concurrent(h, 6, os.fsencode(H.ISOLATE_R1), os.fsencode(H.ISOLATE_R2), os.fsencode(str(tmp_path / "c")), best.ctypes.data) == 1
        assert np.all(best == r["o"].best_index[4])
        # R1 / R2 with different read counts: the reference prints and exits (placement.cpp:189-192); the drop-in raises
        with pytest.raises(RuntimeError, match="does not contain the same number of reads"):
            _place(lib, h, 1, H.ISOLATE_R1, os.path.join(H.REF_DATA, "MZ515733.1.fastq"), str(tmp_path / "x.tsv"))
    finally:
        lib.dropin_close(h)


@pytest.mark.skipif(not os.path.exists(H.RSV_IDX), reason="reference-built rsv_4K index not staged")
def test_dropin_on_the_reference_e2e_fixture_incl_quality_filter_and_empty_metric(lib, tmp_path):
    """rsv_4K + MZ515733.1.fastq (run_e2e.sh): the root carries no seeds, so weighted_containment stays unplaced (empty TSV field)"""
    h = lib.dropin_open(os.fsencode(H.RSV_IDX))
    assert h, lib.dropin_last_error()
    try:
        fq = os.path.join(H.REF_DATA, "MZ515733.1.fastq")
        g = _place(lib, h, 1, fq, "", str(tmp_path / "gpu.tsv"))
        r = _place(lib, h, 0, fq, "", str(tmp_path / "ref.tsv"))
        _same(g, r)
        assert g["ids"][0] == "MZ515733.1" and g["ids"][3] == "" and "weighted_containment\t0.000000\t\n" in g["tsv"]
        for q in (15, 30):
            _same(_place(lib, h, 1, fq, "", str(tmp_path / "gq.tsv"), min_seed_quality=q), _place(lib, h, 0, fq, "", str(tmp_path / "rq.tsv"), min_seed_quality=q))
        # FASTA input (no qualities) and no reads at all
        fa = os.path.join(H.REF_DATA, "MZ515733.1.fa")
        _same(_place(lib, h, 1, fa, "", str(tmp_path / "ga.tsv")), _place(lib, h, 0, fa, "", str(tmp_path / "ra.tsv")))
    finally:
        lib.dropin_close(h)
