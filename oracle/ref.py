"""ctypes binding of oracle/_ref/libpanmap_ref.so: the reference's own translation units (TEST INFRASTRUCTURE)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
_SO = os.path.join(REF_DIR, "libpanmap_ref.so")
REFERENCE_ROOT = "/root/reference"
_lib = None


def available():
    return os.path.exists(_SO)


def build():
    """Build from /root/reference when present (this container); the GPU box only uses the prebuilt files."""
    if not os.path.isdir(REFERENCE_ROOT):
        return available()
    subprocess.run(["make", "-f", os.path.join(_HERE, "ref_build", "Makefile"), "-j8"], check=True, cwd=os.path.dirname(_HERE),
                   stdout=subprocess.DEVNULL)
    return available()


class PlaceOut(C.Structure):
    _fields_ = [("best_score", C.c_double * 5), ("best_index", C.c_uint32 * 5), ("tied_count", C.c_int64 * 5), ("total_reads", C.c_int64),
                ("read_unique_seed_count", C.c_uint64), ("total_read_seed_frequency", C.c_int64), ("read_magnitude", C.c_double),
                ("unique_seeds", C.c_int64), ("seconds", C.c_double)]


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_SO)
        L.ref_last_error.restype = C.c_char_p
        L.ref_rolling_syncmers.restype = C.c_int64
        L.ref_node_genome.restype = C.c_int64
        L.ref_index_open.restype = C.c_void_p
        L.ref_index_open.argtypes = [C.c_char_p]
        L.ref_index_close.argtypes = [C.c_void_p]
        L.ref_index_num_nodes.restype = C.c_int64
        L.ref_index_num_nodes.argtypes = [C.c_void_p]
        L.ref_place.restype = C.c_void_p
        L.ref_place.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.POINTER(PlaceOut)]
        L.ref_place_free.argtypes = [C.c_void_p]
        L.ref_place_tied.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_place_seed_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_node_metrics.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_select_chain.restype = C.c_int64
        L.ref_build_index.argtypes = [C.c_char_p, C.c_char_p] + [C.c_int] * 8
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def hash_seq(seq):
    b = seq.encode() if isinstance(seq, str) else bytes(seq)
    f, r = C.c_uint64(), C.c_uint64()
    lib().ref_hash_seq(b, len(b), C.byref(f), C.byref(r))
    return f.value, r.value


def rolling_syncmers(seq, k, s, open=False, t=0, return_all=True):
    b = seq.encode() if isinstance(seq, str) else bytes(seq)
    cap = max(len(b) - k + 1, 1)
    h = np.zeros(cap, np.uint64); rv = np.zeros(cap, np.uint8); sy = np.zeros(cap, np.uint8); ps = np.zeros(cap, np.int64)
    n = lib().ref_rolling_syncmers(b, C.c_int64(len(b)), k, s, int(bool(open)), t, int(bool(return_all)), _p(h), _p(rv), _p(sy), _p(ps), C.c_int64(cap))
    return h[:n], rv[:n], sy[:n], ps[:n]


def build_index(panman, idx_path, k=19, s=8, t=0, l=3, open=False, flank_mask=250, hpc=False, threads=1):
    rc = lib().ref_build_index(os.fsencode(panman), os.fsencode(idx_path), k, s, t, l, int(bool(open)), flank_mask, int(bool(hpc)), threads)
    if rc != 0:
        raise RuntimeError(lib().ref_last_error().decode())


def node_genome(panman, node_id):
    cap = 1 << 26
    buf = C.create_string_buffer(cap)
    n = lib().ref_node_genome(os.fsencode(panman), node_id.encode(), buf, C.c_int64(cap))
    if n < 0:
        raise RuntimeError(lib().ref_last_error().decode())
    return buf.raw[:n].decode()


class RefIndex:
    def __init__(self, idx_path):
        self.h = lib().ref_index_open(os.fsencode(idx_path))
        if not self.h:
            raise RuntimeError(lib().ref_last_error().decode())
        self.n_nodes = lib().ref_index_num_nodes(self.h)

    def close(self):
        if self.h:
            lib().ref_index_close(self.h)
            self.h = None

    def place(self, r1, r2="", out_tsv="/dev/null", threads=1, min_read_support=-1, seed_mask_fraction=0.0, trim_start=0, trim_end=0,
              dedup=False, force_leaf=False, min_seed_quality=0, stage_timers=False):
        """the reference placement::placeLite with CLI-default options.  stage_timers: also return the reference's own stage times
        (its debug log lines, captured by the driver): stage_ms = dict(read_processing, seeding, dedup, traversal, total)"""
        o = PlaceOut()
        lib().ref_set_stage_timers(int(bool(stage_timers)))
        lib().ref_set_min_seed_quality(int(min_seed_quality))
        k = lib().ref_place(self.h, os.fsencode(r1), os.fsencode(r2), os.fsencode(out_tsv), threads, min_read_support,
                            seed_mask_fraction, trim_start, trim_end, int(dedup), int(force_leaf), 0, C.byref(o))
        lib().ref_set_min_seed_quality(0)
        if not k:
            raise RuntimeError(lib().ref_last_error().decode())
        tied = []
        for m in range(5):
            t = np.zeros(max(o.tied_count[m], 1), np.uint32)
            lib().ref_place_tied(k, m, _p(t))
            tied.append(t[:o.tied_count[m]].copy())
        h = np.zeros(max(o.unique_seeds, 1), np.uint64); c = np.zeros(max(o.unique_seeds, 1), np.int64)
        lib().ref_place_seed_table(k, _p(h), _p(c))
        h, c = h[:o.unique_seeds], c[:o.unique_seeds]
        order = np.argsort(h, kind="stable")
        lib().ref_place_free(k)
        out = dict(best_score=np.array(o.best_score), best_index=np.array(o.best_index), tied=tied, total_reads=o.total_reads,
                   kept=o.read_unique_seed_count, total_frequency=o.total_read_seed_frequency, magnitude=o.read_magnitude,
                   unique_seeds=o.unique_seeds, seconds=o.seconds, table_hash=h[order], table_count=c[order])
        if stage_timers:
            ms = (C.c_double * 5)()
            lib().ref_last_stage_ms(ms)
            lib().ref_set_stage_timers(0)
            out["stage_ms"] = dict(zip(("read_processing", "seeding", "dedup", "traversal", "total"), [float(x) for x in ms]))
        return out

    def node_metrics(self, table_hash, table_count, min_read_support=-1):
        N = self.n_nodes
        metrics = np.zeros((N, 7), np.float64); scores = np.zeros((N, 5), np.float64); scal = np.zeros(6, np.float64)
        th = np.ascontiguousarray(table_hash, np.uint64); tc = np.ascontiguousarray(table_count, np.int64)
        rc = lib().ref_node_metrics(self.h, _p(th), _p(tc), C.c_int64(th.size), min_read_support, _p(metrics), _p(scores), _p(scal))
        if rc != 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return metrics, scores, dict(min_support=int(scal[0]), kept=int(scal[1]), magnitude=scal[2], log_sum=scal[3], wc_denominator=scal[4],
                                     total_frequency=int(scal[5]))


def select_chain(order, score):
    order = np.ascontiguousarray(order, np.uint32); score = np.ascontiguousarray(score, np.float64)
    tied = np.zeros(2 * order.size + 2, np.uint32)
    bs, bi = C.c_double(), C.c_uint32()
    m = lib().ref_select_chain(_p(order), _p(score), C.c_int64(order.size), C.byref(bs), C.byref(bi), _p(tied), C.c_int64(tied.size))
    return bs.value, bi.value, tied[:m].copy()


def write_index(path, idx):
    """flat arrays (hash/parent/child/offsets/parent_index + k,s,t,l,open) -> uncompressed .idx via the reference's capnp schema"""
    h = np.ascontiguousarray(idx.hash, np.uint64); p = np.ascontiguousarray(idx.parent, np.int16); c = np.ascontiguousarray(idx.child, np.int16)
    o = np.ascontiguousarray(idx.offsets, np.uint64); pi = np.ascontiguousarray(idx.parent_index, np.uint32)
    f = lib().ref_write_index
    f.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64] + [C.c_int] * 5
    rc = f(os.fsencode(path), _p(h), _p(p), _p(c), _p(o), _p(pi), pi.size, h.size, idx.k, idx.s, idx.t, idx.l,
           int(idx.open) | (int(getattr(idx, "hpc", 0)) << 1))
    if rc != 0:
        raise RuntimeError(lib().ref_last_error().decode())
