"""ctypes binding of oracle/panmap_oracle.c (TEST INFRASTRUCTURE: the checker, never the product)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "panmap_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O2", "-std=gnu11", "-ffp-contract=off", "-fPIC", "-shared", "-D_GNU_SOURCE", "-o", _SO, src, "-lm"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_rolling_syncmers.restype = C.c_int64
        L.orc_read_seeds.restype = C.c_int64
        L.orc_seed_table.restype = C.c_int64
        L.orc_seed_table_quality.restype = C.c_int64
        L.orc_mask_top_seeds.restype = C.c_int64
        L.orc_resolve_min_read_support.restype = C.c_int64
        L.orc_select_chain.restype = C.c_int64
        L.orc_weighted_denominator.restype = C.c_double
        L.orc_chash.restype = C.c_uint64
        L.orc_rol.restype = C.c_uint64
        L.orc_rol.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_ror.restype = C.c_uint64
        L.orc_ror.argtypes = [C.c_uint64, C.c_uint64]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def hash_seq(seq):
    b = seq.encode() if isinstance(seq, str) else bytes(seq)
    f, r = C.c_uint64(), C.c_uint64()
    rc = lib().orc_hash_seq(b, len(b), C.byref(f), C.byref(r))
    if rc != 0:
        raise ValueError("Kmer contains non canonical base")
    return f.value, r.value


def rolling_syncmers(seq, k, s, open=False, t=0, return_all=True):
    b = seq.encode() if isinstance(seq, str) else bytes(seq)
    cap = max(len(b) - k + 1, 1)
    h = np.zeros(cap, np.uint64); rv = np.zeros(cap, np.uint8); sy = np.zeros(cap, np.uint8); ps = np.zeros(cap, np.int64)
    n = lib().orc_rolling_syncmers(b, C.c_int64(len(b)), k, s, int(bool(open)), t, int(bool(return_all)), _p(h), _p(rv), _p(sy), _p(ps), C.c_int64(cap))
    return h[:n], rv[:n], sy[:n], ps[:n]


def read_seeds(seq, k, s, t, l, open=False, trim_start=0, trim_end=0):
    b = seq.encode() if isinstance(seq, str) else bytes(seq)
    cap = max(len(b) - k + 1, 1)
    out = np.zeros(cap, np.uint64)
    n = lib().orc_read_seeds(b, C.c_int64(len(b)), k, s, t, l, int(bool(open)), trim_start, trim_end, _p(out), C.c_int64(cap))
    return out[:n]


def seed_table(reads, offsets, k, s, t, l, open=False, trim_start=0, trim_end=0, dedup=False, quals=None, min_seed_quality=0):
    """quals + min_seed_quality > 0: the --min-seed-quality construction (placement.cpp:1179-1240, 1388-1533), which ignores dedup"""
    reads = np.ascontiguousarray(reads, np.uint8); offsets = np.ascontiguousarray(offsets, np.uint64)
    ph, pc = C.c_void_p(), C.c_void_p()
    if quals is not None and min_seed_quality > 0:
        quals = np.ascontiguousarray(quals, np.uint8)
        assert quals.size == reads.size
        U = lib().orc_seed_table_quality(_p(reads), _p(quals), _p(offsets), C.c_uint64(offsets.size - 1), k, s, t, l, int(bool(open)),
                                         trim_start, trim_end, int(min_seed_quality), C.byref(ph), C.byref(pc))
    else:
        U = lib().orc_seed_table(_p(reads), _p(offsets), C.c_uint64(offsets.size - 1), k, s, t, l, int(bool(open)), trim_start, trim_end,
                                 int(bool(dedup)), C.byref(ph), C.byref(pc))
    h = np.ctypeslib.as_array(C.cast(ph, C.POINTER(C.c_uint64)), shape=(max(U, 1),))[:U].copy()
    c = np.ctypeslib.as_array(C.cast(pc, C.POINTER(C.c_int64)), shape=(max(U, 1),))[:U].copy()
    lib().orc_free(ph); lib().orc_free(pc)
    return h, c


def mask_top_seeds(hash_, count, frac):
    """orc_mask_top_seeds (placement.cpp:1748-1799): the table without its floor(frac * U) most frequent seeds, sorted by hash;
    ties at the cut go by ascending hash (unspecified in the reference)."""
    h = np.array(hash_, np.uint64); c = np.array(count, np.int64)
    n = lib().orc_mask_top_seeds(_p(h), _p(c), C.c_int64(h.size), C.c_double(frac))
    return h[:n].copy(), c[:n].copy()


def resolve_min_read_support(counts, configured=-1):
    counts = np.ascontiguousarray(counts, np.int64)
    return int(lib().orc_resolve_min_read_support(_p(counts), C.c_int64(counts.size), configured))


def read_magnitudes(counts, min_support):
    counts = np.ascontiguousarray(counts, np.int64)
    logv = np.zeros(max(counts.size, 1), np.float64); scal = np.zeros(5, np.float64)
    lib().orc_read_magnitudes(_p(counts), C.c_int64(counts.size), C.c_int64(min_support), _p(logv), _p(scal))
    return logv[:counts.size], dict(kept=int(scal[0]), magnitude=scal[1], log_sum=scal[2], total=int(scal[3]), dropped=int(scal[4]))


def bfs_order(parent_index):
    parent_index = np.ascontiguousarray(parent_index, np.uint32)
    order = np.zeros(parent_index.size, np.uint32)
    lib().orc_bfs_order(_p(parent_index), C.c_uint64(parent_index.size), _p(order))
    return order


def node_metrics(idx, table_hash, logv, U1, mag, denL, denW):
    """idx: object with hash/parent/child/offsets/parent_index arrays. Returns metrics [N,7], scores [N,5]."""
    N = idx.parent_index.size
    metrics = np.zeros((N, 7), np.float64); scores = np.zeros((N, 5), np.float64)
    th = np.ascontiguousarray(table_hash, np.uint64); lv = np.ascontiguousarray(logv, np.float64)
    lib().orc_node_metrics(_p(idx.hash), _p(idx.parent), _p(idx.child), _p(idx.offsets), _p(idx.parent_index), C.c_uint64(N),
                           _p(th), _p(lv), C.c_int64(th.size), C.c_double(U1), C.c_double(mag), C.c_double(denL), C.c_double(denW),
                           _p(metrics), _p(scores))
    return metrics, scores


def weighted_denominator(idx, table_hash, logv):
    th = np.ascontiguousarray(table_hash, np.uint64); lv = np.ascontiguousarray(logv, np.float64)
    return float(lib().orc_weighted_denominator(_p(idx.hash), _p(idx.child), C.c_uint64(int(idx.offsets[0])), C.c_uint64(int(idx.offsets[1])),
                                                _p(th), _p(lv), C.c_int64(th.size)))


def select_chain(order, score):
    order = np.ascontiguousarray(order, np.uint32); score = np.ascontiguousarray(score, np.float64)
    tied = np.zeros(2 * order.size + 2, np.uint32)
    bs, bi = C.c_double(), C.c_uint32()
    m = lib().orc_select_chain(_p(order), _p(score), C.c_int64(order.size), C.byref(bs), C.byref(bi), _p(tied), C.c_int64(tied.size))
    return bs.value, bi.value, tied[:m].copy()


def place(reads, offsets, idx, trim_start=0, trim_end=0, dedup=False, min_read_support=-1, seed_mask_fraction=0.0,
          force_leaf=False, skip_node=0xFFFFFFFF, want_scores=False, quals=None, min_seed_quality=0):
    """orc_place: the compute part of placeLite on the CPU. idx has hash/parent/child/offsets/parent_index + k,s,t,l,open."""
    reads = np.ascontiguousarray(reads, np.uint8); offsets = np.ascontiguousarray(offsets, np.uint64)
    N = idx.parent_index.size
    best = np.zeros(5, np.float64); bidx = np.zeros(5, np.uint32); tcount = np.zeros(5, np.int64)
    cap = max(N + 2, 4)
    tied = np.zeros((5, cap), np.uint32)
    scores = np.zeros((max(N, 1), 5), np.float64) if want_scores else None
    stats = np.zeros(7, np.float64)
    if quals is not None:
        quals = np.ascontiguousarray(quals, np.uint8)
        assert quals.size == reads.size
    lib().orc_place_q(_p(reads), _p(quals) if quals is not None else None, _p(offsets), C.c_uint64(offsets.size - 1), _p(idx.hash),
                      _p(idx.parent), _p(idx.child), _p(idx.offsets), _p(idx.parent_index), C.c_uint64(N), idx.k, idx.s, idx.t, idx.l,
                      int(idx.open), trim_start, trim_end, int(bool(dedup)), min_read_support, C.c_double(seed_mask_fraction),
                      int(bool(force_leaf)), C.c_uint32(skip_node), int(min_seed_quality), _p(best), _p(bidx),
                      _p(tcount), _p(tied), C.c_int64(cap), _p(scores) if scores is not None else None, _p(stats))
    return dict(best_score=best, best_index=bidx, tied=[tied[m, :tcount[m]].copy() for m in range(5)], scores=scores,
                unique_seeds=int(stats[0]), min_support=int(stats[1]), kept=int(stats[2]), magnitude=stats[3], log_sum=stats[4],
                wc_denominator=stats[5], total_frequency=int(stats[6]))


def hpc_compress(seq):
    """seeding::hpcCompress (seeding.cpp:286-306): a base is dropped when it equals its predecessor ignoring case (std::toupper)."""
    out = bytearray()
    for i, c in enumerate(seq):
        if i == 0 or bytes([c]).upper() != bytes([seq[i - 1]]).upper():
            out.append(c)
    return bytes(out)
