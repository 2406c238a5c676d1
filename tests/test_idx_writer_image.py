"""the product's own `.idx` writer (IndexBuilder::writeIndex, index_single_mode.cpp:1593-1636) and the cached
image of the flattened index (the reference's cache rule: main.cpp:371-396).  CPU part: files round-trip through the product's reader, are
read by the REFERENCE's own IndexReader + placeLite (oracle/_ref) with an unchanged result, and the image survives a host round trip and
refuses stale / damaged files.  GPU part: an index opened through the cache places exactly like one created from the arrays."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import panmap_b200 as pm
from oracle import ref
from tests import helpers as H


def _same_index(a, b):
    assert (a.k, a.s, a.t, a.l, a.open, a.hpc) == (b.k, b.s, b.t, b.l, b.open, b.hpc)
    for f in ("hash", "parent", "child", "offsets"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert np.array_equal(a.parent_index[1:], b.parent_index[1:])
    assert a.node_ids == b.node_ids


@pytest.mark.parametrize("level", [-1, 3])
def test_writer_roundtrip_tiny(tmp_path, level):
    src = pm.HostIndex.read(os.path.join(H.GOLDEN, "tiny.idx"))
    src.identical_to_parent = (np.arange(src.n_nodes) % 3 == 1).astype(np.uint8)
    src.block_ranges = np.array([[0, 99], [100, 2999], [3000, 3000]], np.uint32)
    src.substitution_matrix = np.linspace(0.0, 1.0, 16)
    p = str(tmp_path / "w.idx")
    n = src.write(p, zstd_level=level)
    assert n == os.path.getsize(p)
    raw = open(p, "rb").read(32)
    assert raw[:4] == b"PMI1" and raw[26] == (1 if level < 0 else 0)
    assert np.frombuffer(raw[8:24], np.int32).tolist() == [src.k, src.s, src.t, src.l]
    back = pm.HostIndex.read(p)
    _same_index(src, back)
    assert np.array_equal(back.identical_to_parent, src.identical_to_parent)
    assert np.array_equal(back.block_ranges, src.block_ranges)
    assert np.array_equal(back.substitution_matrix, src.substitution_matrix)


def test_writer_defaults_and_edge_shapes(tmp_path):
    """no ids / no extras; a one-node index without deltas; an hpc + open-syncmer flagged index"""
    one = pm.HostIndex(np.zeros(0, np.uint64), np.zeros(0, np.int16), np.zeros(0, np.int16), np.zeros(2, np.uint64), np.zeros(1, np.uint32), 15, 8, 1, 1, open=1, hpc=1)
    p = str(tmp_path / "one.idx")
    one.write(p)
    b = pm.HostIndex.read(p)
    assert (b.n_nodes, b.n_deltas, b.k, b.s, b.t, b.l, b.open, b.hpc) == (1, 0, 15, 8, 1, 1, 1, 1) and b.node_ids == ["node_0"]
    assert b.block_ranges is None and b.substitution_matrix is None
    with pytest.raises(pm.PanmapError):      # offsets that do not end at n_deltas
        pm.HostIndex(np.zeros(3, np.uint64), np.zeros(3, np.int16), np.ones(3, np.int16), np.array([0, 2], np.uint64), np.zeros(1, np.uint32), 19, 8, 0, 3).write(p)
    with pytest.raises(pm.PanmapError) as e:
        one.write("/nonexistent_dir/x.idx")
    assert e.value.code == -4


@pytest.mark.skipif(not (ref.available() and os.path.exists(H.RSV_IDX)), reason="needs oracle/_ref (built from /root/reference)")
def test_reference_reads_what_the_writer_wrote(tmp_path):
    """the reference's own capnp reader + LiteTree::initialize + placeLite on a file from the product's writer: the TSV it writes is the
    one it writes for its own file (rsv_4K e2e case of run_e2e.sh).  Uncompressed container: the oracle build maps the payload directly
    (the reference's zstd inflate is one of the two link stubs); the zstd frames are checked through libzstd by the product's reader."""
    src = pm.HostIndex.read(H.RSV_IDX)
    p = str(tmp_path / "rsv_rewritten.idx")
    src.write(p, zstd_level=-1)
    _same_index(src, pm.HostIndex.read(p))
    fq = os.path.join(H.REF_DATA, "MZ515733.1.fastq")
    a = ref.RefIndex(H.RSV_IDX).place(fq, "", out_tsv=str(tmp_path / "a.tsv"))
    b = ref.RefIndex(p).place(fq, "", out_tsv=str(tmp_path / "b.tsv"))
    assert open(tmp_path / "a.tsv").read() == open(tmp_path / "b.tsv").read()
    assert list(a["best_index"]) == list(b["best_index"]) and list(a["best_score"]) == list(b["best_score"])


def _hostcheck():
    so = os.path.join(H.ROOT, "tests", "hostcheck", "libhostcheck.so")
    srcs = [os.path.join(H.ROOT, "tests", "hostcheck", "hostcheck.cpp"), os.path.join(H.ROOT, "panmap_b200", "csrc", "pm_flatten.cpp"),
            os.path.join(H.ROOT, "panmap_b200", "csrc", "pm_image.cpp")]
    deps = srcs + [os.path.join(H.ROOT, "panmap_b200", "csrc", "pm_host.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                        "-Wl,--version-script=" + os.path.join(H.ROOT, "tests", "hostcheck", "exports.map"), "-o", so] + srcs, check=True)
    return C.CDLL(so)


@pytest.mark.parametrize("shard,n_shards", [(0, 1), (1, 3)])
def test_image_host_roundtrip(tmp_path, shard, n_shards):
    """flatten -> image -> flatten'; a second write of what was read is byte-identical; stale stamp / flipped bit / truncation are misses"""
    hc = _hostcheck()
    hc.hc_last_error.restype = C.c_char_p
    src = pm.HostIndex.read(os.path.join(H.GOLDEN, "tiny.idx"))
    d = src.desc()
    ids = (C.c_char_p * src.n_nodes)(*[i.encode() for i in src.node_ids])
    rc = hc.hc_image_roundtrip(C.byref(d), C.c_uint32(shard), C.c_uint32(n_shards), ids, os.fsencode(str(tmp_path)))
    assert rc == 0, (rc, hc.hc_last_error())


def test_image_write_is_host_only(tmp_path):
    """pm_index_image_write needs no device (an index can be flattened where it is built); opening it does"""
    src = pm.HostIndex.read(os.path.join(H.GOLDEN, "tiny.idx"))
    d = src.desc()
    n = C.c_uint64()
    p = str(tmp_path / "t.pmflat")
    rc = pm.lib().pm_index_image_write(C.byref(d), None, 0, 1, os.fsencode(p), C.byref(n))
    assert rc == 0 and n.value == os.path.getsize(p) and n.value > 4 * src.n_deltas
    if pm.device_count() < 1:
        with pytest.raises(pm.PanmapError) as e:
            pm.Index.from_image(p)
        assert e.value.code == -2


@pytest.mark.gpu
def test_cached_index_places_like_a_fresh_one(tmp_path):
    """miss -> image written; hit -> same placement, ids from the image; a changed source (mtime) is a miss again; a damaged image too"""
    import shutil
    from tools.synth import synth
    S = synth.generate(3000, 8000, 1.5, 4000, seed=9)
    host = pm.HostIndex(S.hash, S.parent, S.child, S.offsets, S.parent_index, S.k, S.s, S.t, S.l, node_ids=[f"n{i}" for i in range(S.n_nodes)])
    idx = str(tmp_path / "s.idx")
    host.write(idx, zstd_level=3)
    want = pm.Workspace(pm.Index(host)).place(S.reads, S.read_offsets)
    a = pm.Index.open_cached(idx)
    assert a.cache_hit is False and os.path.exists(idx + ".pmflat")
    b = pm.Index.open_cached(idx)
    assert b.cache_hit is True and b.node_id(7) == "n7"
    for ix in (a, b, pm.Index.from_image(idx + ".pmflat")):
        got = pm.Workspace(ix).place(S.reads, S.read_offsets)
        for m in pm.METRICS:
            assert got.best_index[m] == want.best_index[m] and got.best_score[m] == want.best_score[m] and np.array_equal(got.tied[m], want.tied[m])
        assert got.tsv() == want.tsv()
    os.utime(idx, ns=(os.stat(idx).st_atime_ns, os.stat(idx).st_mtime_ns + 1_000_000_000))
    c = pm.Index.open_cached(idx)
    assert c.cache_hit is False
    assert pm.Index.open_cached(idx).cache_hit is True
    raw = bytearray(open(idx + ".pmflat", "rb").read()); raw[len(raw) // 3] ^= 1
    open(idx + ".pmflat", "wb").write(raw)
    d = pm.Index.open_cached(idx)
    assert d.cache_hit is False and pm.Index.open_cached(idx).cache_hit is True      # rewritten
    # shards get their own images
    s1 = pm.Index.open_cached(idx, shard=1, n_shards=2)
    assert s1.cache_hit is False and os.path.exists(idx + ".pmflat.1of2") and pm.Index.open_cached(idx, shard=1, n_shards=2).cache_hit is True
    assert s1.shard_range() == pm.Index(host, shard=1, n_shards=2).shard_range()
    # the real sars_20000 index (config 1): the golden TSV through the cache
    if os.path.exists(H.SARS_IDX):
        local = str(tmp_path / "sars.idx"); shutil.copyfile(H.SARS_IDX, local)
        pm.Index.open_cached(local)
        ws = pm.Workspace(pm.Index.open_cached(local))
        assert ws.index.cache_hit is True
        pm.place_files(ws, H.ISOLATE_R1, H.ISOLATE_R2, str(tmp_path / "iso.tsv"))
        assert open(tmp_path / "iso.tsv").read() == open(H.ISOLATE_TSV).read()
