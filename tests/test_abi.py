"""The C-ABI library loads without a GPU, exports every symbol include/panmap_b200.h declares, reads .idx files, validates
its inputs, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import panmap_b200 as pm
from tests import helpers as H


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(H.ROOT, "include", "panmap_b200.h")).read()
    names = sorted(set(re.findall(r"\b(pm_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 30
    L = pm.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.pm_abi_version() == 5


def test_idx_reader_roundtrip_of_a_reference_written_index():
    """tests/golden/tiny.idx was written by the reference's capnp schema (oracle/ref.py write_index)"""
    g = np.load(os.path.join(H.GOLDEN, "tiny_index.npz"))
    hi = pm.HostIndex.read(os.path.join(H.GOLDEN, "tiny.idx"))
    assert (hi.k, hi.s, hi.t, hi.l, hi.open) == (19, 8, 0, 3, 0)
    assert np.array_equal(hi.hash, g["hash"]) and np.array_equal(hi.parent, g["parent"]) and np.array_equal(hi.child, g["child"])
    assert np.array_equal(hi.offsets, g["offsets"]) and np.array_equal(hi.parent_index[1:], g["parent_index"][1:])
    assert hi.node_ids[:3] == ["node_0", "node_1", "node_2"]


def test_idx_reader_errors():
    with pytest.raises(pm.PanmapError) as e:
        pm.HostIndex.read("/nonexistent/x.idx")
    assert e.value.code == -4
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".idx") as f:
        f.write(b"PMI1" + b"\x00" * 60); f.flush()      # right magic, wrong header version / truncated message
        with pytest.raises(pm.PanmapError):
            pm.HostIndex.read(f.name)


def test_idx_reader_zstd_framed_payload(tmp_path):
    """the reference writes the payload as independent zstd frames by default (index_single_mode.cpp:1615-1633)"""
    import ctypes.util
    z = C.CDLL(ctypes.util.find_library("zstd") or "libzstd.so.1")
    z.ZSTD_compressBound.restype = C.c_size_t; z.ZSTD_compressBound.argtypes = [C.c_size_t]
    z.ZSTD_compress.restype = C.c_size_t; z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    raw = open(os.path.join(H.GOLDEN, "tiny.idx"), "rb").read()
    hdr, payload = bytearray(raw[:32]), raw[32:]
    hdr[26] = 0
    frames = b""
    cut = (len(payload) // 2) & ~7
    for part in (payload[:cut], payload[cut:]):           # two frames, like the reference's 64 MB framing
        cap = z.ZSTD_compressBound(len(part)); dst = C.create_string_buffer(cap)
        n = z.ZSTD_compress(dst, cap, part, len(part), 3)
        frames += dst.raw[:n]
    p = tmp_path / "z.idx"
    p.write_bytes(bytes(hdr) + frames)
    a = pm.HostIndex.read(str(p)); b = pm.HostIndex.read(os.path.join(H.GOLDEN, "tiny.idx"))
    assert np.array_equal(a.hash, b.hash) and np.array_equal(a.offsets, b.offsets) and a.node_ids == b.node_ids and (a.k, a.l) == (b.k, b.l)


@pytest.mark.skipif(pm.device_count() > 0, reason="a CUDA device is present")
def test_compute_fails_loudly_without_a_device():
    rng = np.random.default_rng(0)
    idx, _, _ = H.synthetic_index(20, rng)
    host = pm.HostIndex(idx.hash, idx.parent, idx.child, idx.offsets, idx.parent_index, 19, 8, 0, 3)
    with pytest.raises(pm.PanmapError) as e:
        pm.Index(host)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(pm.PanmapError) as e:
        pm.rolling_syncmers([b"ACGTACGTACGTACGTACGTACGT"], 19, 8)
    assert e.value.code == -2


@pytest.mark.skipif(not os.path.exists(H.ISOLATE_R1), reason="isolate reads not staged (oracle/_ref/data)")
def test_fastx_ingest_matches_kseq_semantics():
    """extractReadSequences of the C++ host shim (placement.cpp:164-197): gz FASTQ pairs interleaved, FASTA, plain FASTQ"""
    b, o = pm.read_fastx(H.ISOLATE_R1, H.ISOLATE_R2)
    eb, eo = pm.pack_reads(H.isolate_reads())
    assert np.array_equal(b, eb) and np.array_equal(o, eo) and o.size - 1 == 102338
    b, o = pm.read_fastx(os.path.join(H.REF_DATA, "MZ515733.1.fa"))
    assert o.size - 1 == 1 and int(o[-1]) == 14942
    with pytest.raises(pm.PanmapError):     # R1/R2 count mismatch (placement.cpp:189-192)
        pm.read_fastx(H.ISOLATE_R1, os.path.join(H.REF_DATA, "MZ515733.1.fastq"))


def test_fastx_ingest_small_cases(tmp_path):
    p = tmp_path / "a.fq"
    p.write_text("@r1 x\nACGT\nNN\n+\nIIIIII\n@r2\nGG\n+r2\n@@\n")      # multi-line sequence, '@' inside a quality line
    b, o = pm.read_fastx(str(p))
    assert bytes(b) == b"ACGTNNGG" and list(o) == [0, 6, 8]
    q = tmp_path / "b.fa"
    q.write_text(">s1\nAC\r\nGT\n>s2\n\n>s3\nTT")
    b, o = pm.read_fastx(str(q))
    assert bytes(b) == b"ACGTTT" and list(o) == [0, 4, 4, 6]


def test_fastq_parallel_ingest_equals_the_serial_parser(tmp_path):
    """plain strict four-line FASTQ goes through the multi-threaded flat parser (the reference's parallelFastqSeqs fast path,
    placement.cpp:96-162); the same records gzipped go through the serial kseq-style parser: both must give the same reads, in file
    order, also for paired files, \\r\\n line ends, '@' opening a quality line, an empty read and a missing final newline"""
    import gzip
    rng = np.random.default_rng(3)
    reads = H.random_reads(rng, 12000, lo=1, hi=250) + [b"", b"ACGT"]
    quals = []
    for r in reads:
        q = (rng.integers(0, 41, len(r)) + 33).astype(np.uint8)
        if len(r):
            q[0] = ord("@")                                # a quality line that looks like a header
        quals.append(q.tobytes())
    def records(rs, qs, eol=b"\n", last_eol=True):
        out = b"".join(b"@r%d some text" % i + eol + r + eol + b"+" + eol + q + eol for i, (r, q) in enumerate(zip(rs, qs)))
        return out if last_eol else out[:-len(eol)]
    eb, eo = pm.pack_reads(reads)
    for name, blob in [("a.fq", records(reads, quals)), ("b.fq", records(reads, quals, b"\r\n")), ("c.fq", records(reads, quals, last_eol=False))]:
        p = tmp_path / name
        p.write_bytes(blob)
        assert len(blob) > (1 << 20)                       # large enough for more than one parser thread
        b, o = pm.read_fastx(str(p))
        assert np.array_equal(o, eo) and np.array_equal(b, eb), name
        g = tmp_path / (name + ".gz")
        with gzip.open(g, "wb", compresslevel=1) as f:
            f.write(blob)
        b2, o2 = pm.read_fastx(str(g))
        assert np.array_equal(o2, eo) and np.array_equal(b2, eb), name
    # pairs: R1 plain (parallel parser), R2 gzipped (serial parser), interleaved
    r2 = [r[::-1] for r in reads]
    (tmp_path / "r2.fq").write_bytes(records(r2, quals))
    with gzip.open(tmp_path / "r2.fq.gz", "wb", compresslevel=1) as f:
        f.write(records(r2, quals))
    inter = [x for pair in zip(reads, r2) for x in pair]
    ib, io = pm.pack_reads(inter)
    for second in ("r2.fq", "r2.fq.gz"):
        b, o = pm.read_fastx(str(tmp_path / "a.fq"), str(tmp_path / second))
        assert np.array_equal(o, io) and np.array_equal(b, ib), second
    # not four lines per record: the serial parser takes over
    m = tmp_path / "m.fq"
    m.write_bytes(b"@x\nAC\nGT\n+\nIIII\n" * 3)
    b, o = pm.read_fastx(str(m))
    assert bytes(b) == b"ACGT" * 3 and list(o) == [0, 4, 8, 12]


def test_idx_reader_rejects_truncated_and_corrupt_files(tmp_path):
    """every list body of the schema-less Cap'n Proto walk is bounds-checked: a file cut short, or one whose list headers name more
    elements than its segments hold, must come back as a clean PM_ERR_INVALID (the reference's capnp reader throws), never a crash"""
    raw = open(os.path.join(H.GOLDEN, "tiny.idx"), "rb").read()
    good = pm.HostIndex.read(os.path.join(H.GOLDEN, "tiny.idx"))
    n_ok = 0
    for cut in list(range(40, len(raw), max(1, len(raw) // 97))) + [len(raw) - 8, len(raw) - 1]:
        p = tmp_path / "cut.idx"
        p.write_bytes(raw[:cut])
        try:
            hi = pm.HostIndex.read(str(p))
            # a cut inside trailing padding may still parse: then it must parse to the same index
            assert np.array_equal(hi.hash, good.hash) and np.array_equal(hi.offsets, good.offsets)
            n_ok += 1
        except pm.PanmapError as e:
            assert e.code == -1, (cut, e)
    assert n_ok <= 3
    # corrupt list headers: blow up the element count of every list pointer in turn
    import struct
    words = (len(raw) - 32) // 8
    hits = 0
    for w in range(2, words):
        v, = struct.unpack_from("<Q", raw, 32 + 8 * w)
        if (v & 3) != 1 or (v >> 35) == 0 or (v >> 35) > 1 << 20:   # list pointers with a plausible element count
            continue
        bad = bytearray(raw)
        struct.pack_into("<Q", bad, 32 + 8 * w, (v & ((1 << 35) - 1)) | ((1 << 27) << 35))
        p = tmp_path / "bad.idx"
        p.write_bytes(bytes(bad))
        try:
            pm.HostIndex.read(str(p))
        except pm.PanmapError as e:
            assert e.code == -1
            hits += 1
    assert hits >= 3


def test_fastx_serial_parser_keeps_kseq_corner_cases(tmp_path):
    """kseq keeps blanks inside a sequence line (only the line end and a trailing CR go) and `while (kseq_read(seq) >= 0)` stops at a
    record whose quality string is shorter than its sequence, whether or not the qualities were asked for (placement.cpp:164-176)"""
    p = tmp_path / "t.fq"
    p.write_text("@a\nAC GT\n+\nIIIII\n@b\nACGT\n+\nII\n")         # record b: quality cut short by the end of the file
    b, o = pm.read_fastx(str(p))
    assert bytes(b) == b"AC GT" and list(o) == [0, 5]
    p.write_text("@a\nACGT\n+\nII\n@c\nGG\n+\nII\n")               # kseq reads "II" + "@c" as a's four quality bytes, then finds no header
    b, o = pm.read_fastx(str(p))
    assert bytes(b) == b"ACGT" and list(o) == [0, 4]


def test_host_packer_writes_the_4bit_layout_of_pm_place_packed():
    """pm_pack_reads (multi-threaded above 4096 reads): A/C/G/T in either case -> 0..3, everything else 4, 32 bases per 16-byte chunk with
    base j in bits [4j, 4j+4), slots past a read's end = 4, every read on a chunk boundary"""
    rng = np.random.default_rng(17)
    reads = H.random_reads(rng, 9000, lo=0, hi=170, p_n=0.03, p_lower=0.05) + [b"", b"ACGTRYKM-*acgtn", b"T" * 32, b"G" * 33, b"C" * 31]
    buf, off = pm.pack_reads(reads)
    pk = pm.host_pack_reads(buf, off)
    assert pk.ctypes.data % 16 == 0
    lut = np.full(256, 4, np.uint8)
    for ch, v in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
        lut[ch] = v
    lens = np.diff(off).astype(np.int64)
    chunks = (lens + 31) // 32
    assert pk.size == int(chunks.sum()) * 16 == int(pm.lib().pm_packed_chunks(off.ctypes.data, len(reads))) * 16
    nib = np.empty(pk.size * 2, np.uint8)
    nib[0::2] = pk & 15; nib[1::2] = pk >> 4
    exp = np.full(nib.size, 4, np.uint8)
    starts = np.concatenate([[0], np.cumsum(chunks)[:-1]]) * 32
    for r in range(len(reads)):
        exp[starts[r]:starts[r] + lens[r]] = lut[buf[int(off[r]):int(off[r + 1])]]
    assert np.array_equal(nib, exp)
    one = pm.host_pack_reads(buf, off, threads=1)
    assert np.array_equal(one, pk)
