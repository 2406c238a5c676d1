"""Shared test helpers: data locations, FASTQ reading (kseq semantics), small synthetic indexes."""
import gzip
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DATA = os.path.join(ROOT, "oracle", "_ref", "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")
SARS_IDX = os.path.join(REF_DATA, "sars_20000.k19s8t0l3.idx")
RSV_IDX = os.path.join(REF_DATA, "rsv_4K.k19s8t0l3.idx")
MAMMOTH_IDX = os.path.join(REF_DATA, "extended_mammoth.k15s8t0l1.idx")
# index-builder parity: the bundled .panman files and the indexes the reference builds from them with --flank-mask 0
SARS_PANMAN = os.path.join(REF_DATA, "sars_20000_twilight_dipper.panman")
RSV_PANMAN = os.path.join(REF_DATA, "rsv_4K.panman")
MAMMOTH_PANMAN = os.path.join(REF_DATA, "extended_mammoth.panman")
SARS_IDX_F0 = os.path.join(REF_DATA, "sars_20000.k19s8t0l3.flank0.idx")
RSV_IDX_F0 = os.path.join(REF_DATA, "rsv_4K.k19s8t0l3.flank0.idx")
MAMMOTH_IDX_F0 = os.path.join(REF_DATA, "extended_mammoth.k15s8t0l1.flank0.idx")
ISOLATE_R1 = os.path.join(REF_DATA, "isolate_R1.fastq.gz")
ISOLATE_R2 = os.path.join(REF_DATA, "isolate_R2.fastq.gz")
ISOLATE_TSV = os.path.join(REF_DATA, "isolate.placement.tsv")


def read_fastx(path):
    """sequences of a FASTA/FASTQ(.gz) file in file order (what kseq_read yields in extractReadSequences,
    placement.cpp:164-176)."""
    op = gzip.open if path.endswith(".gz") else open
    seqs = []
    with op(path, "rb") as f:
        data = f.read()
    lines = data.split(b"\n")
    i = 0
    n = len(lines)
    while i < n:
        ln = lines[i]
        if ln.startswith(b"@"):
            seq = []
            i += 1
            while i < n and not lines[i].startswith(b"+"):
                seq.append(lines[i].strip())
                i += 1
            s = b"".join(seq)
            i += 1
            got = 0
            while i < n and got < len(s):
                got += len(lines[i].strip())
                i += 1
            seqs.append(s)
        elif ln.startswith(b">"):
            seq = []
            i += 1
            while i < n and not lines[i].startswith(b">"):
                seq.append(lines[i].strip())
                i += 1
            seqs.append(b"".join(seq))
        else:
            i += 1
    return seqs


def interleave(r1, r2):
    """seeding::perfect_shuffle (seeding.hpp:33-43): R1_0, R2_0, R1_1, R2_1, ..."""
    assert len(r1) == len(r2)
    out = [None] * (2 * len(r1))
    out[0::2] = r1
    out[1::2] = r2
    return out


def isolate_reads():
    return interleave(read_fastx(ISOLATE_R1), read_fastx(ISOLATE_R2))


class FlatIdx:
    """arrays in reference-native widths + seeding parameters (duck-types panmap_b200.HostIndex for the oracle)"""

    def __init__(self, hash, parent, child, offsets, parent_index, k, s, t, l, open=0):
        self.hash = np.ascontiguousarray(hash, np.uint64)
        self.parent = np.ascontiguousarray(parent, np.int16)
        self.child = np.ascontiguousarray(child, np.int16)
        self.offsets = np.ascontiguousarray(offsets, np.uint64)
        self.parent_index = np.ascontiguousarray(parent_index, np.uint32)
        self.k, self.s, self.t, self.l, self.open = k, s, t, l, open


def random_tree(n, rng):
    """DFS pre-order parents: node i hangs under i-1 with p=0.5, else under a random proper ancestor of i-1."""
    parent = np.zeros(n, np.uint32)
    path = [0]
    for i in range(1, n):
        if rng.random() < 0.5 or len(path) == 1:
            p = i - 1
        else:
            cut = int(rng.integers(0, len(path) - 1))
            p = path[cut]
            del path[cut + 1:]
        while path and path[-1] != p:
            path.pop()
        parent[i] = p
        path.append(i)
    return parent


def synthetic_index(n_nodes, rng, universe=4000, root_seeds=600, max_changes=12, k=19, s=8, t=0, l=3, big_node=None, zero_frac=0.25):
    """A structurally valid seed-delta index built from per-node seed multisets (counts 0..4): deltas are exact
    (hash, parentCount, childCount) differences sorted by hash with parentCount != childCount."""
    parent = random_tree(n_nodes, rng)
    hashes = rng.integers(1, 2**63, size=universe, dtype=np.int64).astype(np.uint64)
    hashes = np.unique(hashes)
    universe = hashes.size
    state = {}
    dh, dp, dc, off = [], [], [], [0]
    stack = []
    genomes = [None] * n_nodes
    for v in range(n_nodes):
        base = {} if v == 0 else dict(genomes[parent[v]])
        changes = {}
        if v == 0:
            for j in rng.choice(universe, size=root_seeds, replace=False):
                changes[int(j)] = int(rng.choice([1, 1, 1, 2, 3]))
        elif big_node is not None and v == big_node[0]:
            for j in rng.choice(universe, size=big_node[1], replace=False):
                changes[int(j)] = int(rng.choice([0, 1, 1, 2]))
        elif rng.random() >= zero_frac:
            for j in rng.choice(universe, size=int(rng.integers(1, max_changes + 1)), replace=False):
                changes[int(j)] = int(rng.choice([0, 0, 1, 1, 1, 2, 4]))
        rows = []
        for j, c in changes.items():
            p = base.get(j, 0)
            if p != c:
                rows.append((int(hashes[j]), p, c))
                if c:
                    base[j] = c
                else:
                    base.pop(j, None)
        rows.sort()
        for h, p, c in rows:
            dh.append(h); dp.append(p); dc.append(c)
        off.append(len(dh))
        genomes[v] = base
    idx = FlatIdx(np.array(dh, np.uint64), np.array(dp, np.int16), np.array(dc, np.int16), np.array(off, np.uint64), parent, k, s, t, l)
    return idx, hashes, genomes


def random_reads(rng, n, lo=30, hi=160, p_n=0.01, p_lower=0.02):
    out = []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1))
        a = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=L)
        m = rng.random(L)
        a = np.where(m < p_n, ord("N"), a)
        a = np.where((m >= p_n) & (m < p_n + p_lower), a | 0x20, a).astype(np.uint8)
        out.append(a.tobytes())
    return out


def relerr(a, b, floor=1e-9):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)
