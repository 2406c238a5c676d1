"""ctypes binding of tools/synth/pm_synth.cpp: synthetic PanMAN-shaped index + reads (SURVEY.md §8d shapes)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpm_synth.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "pm_synth.cpp")
    dep = os.path.join(_HERE, "..", "..", "panmap_b200", "csrc", "pm_logic.cuh")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wl,--version-script=" + os.path.join(_HERE, "exports.map"),
                        "-o", _SO, src], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.synth_generate.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_uint64, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_double]
        L.synth_num_deltas.restype = C.c_uint64
        L.synth_reads_bytes.restype = C.c_uint64
        L.synth_truth_node.restype = C.c_uint32
        _lib = L
    return _lib


class Synth:
    pass


def generate(n_nodes, genome_len, lam, n_reads, read_len=150, sub_rate=0.005, n_rate=0.001, k=19, s=8, t=0, l=3, open=0, seed=0, truth_frac=0.618):
    L = lib()
    L.synth_generate(n_nodes, genome_len, lam, n_reads, read_len, sub_rate, n_rate, k, s, t, l, open, seed, truth_frac)
    D = L.synth_num_deltas()
    o = Synth()
    o.hash = np.zeros(D, np.uint64); o.parent = np.zeros(D, np.int16); o.child = np.zeros(D, np.int16)
    o.offsets = np.zeros(n_nodes + 1, np.uint64); o.parent_index = np.zeros(n_nodes, np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    L.synth_get_index(p(o.hash), p(o.parent), p(o.child), p(o.offsets), p(o.parent_index))
    o.reads = np.zeros(max(L.synth_reads_bytes(), 1), np.uint8)[:L.synth_reads_bytes()]
    o.read_offsets = np.zeros(n_reads + 1, np.uint64)
    if n_reads:
        buf = np.zeros(L.synth_reads_bytes(), np.uint8)
        L.synth_get_reads(p(buf), p(o.read_offsets))
        o.reads = buf
    o.truth = L.synth_truth_node()
    g = np.zeros(genome_len, np.uint8); L.synth_get_truth_genome(p(g)); o.truth_genome = g.tobytes()
    o.k, o.s, o.t, o.l, o.open = k, s, t, l, open
    o.n_nodes, o.n_deltas = n_nodes, int(D)
    L.synth_free()
    return o
