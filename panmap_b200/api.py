"""ctypes binding of include/panmap_b200.h (mirrors placement::placeLite / seeding::rollingSyncmers)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
METRICS = ("log_raw", "log_cosine", "containment", "weighted_containment", "log_containment")
PM_NONE = 0xFFFFFFFF


class PanmapError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"panmap_b200 error {code}: {msg}")
        self.code = code


def lib_path():
    return os.path.join(_HERE, "libpanmap_b200.so")


def build(force=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    mk = os.path.join(_HERE, "csrc", "Makefile")
    args = ["make", "-f", mk] + (["-B"] if force else [])
    subprocess.run(args, check=True, cwd=os.path.dirname(_HERE))


class SeedParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("s", C.c_int32), ("t", C.c_int32), ("l", C.c_int32), ("open", C.c_int32), ("hpc", C.c_int32)]


class IndexDesc(C.Structure):
    _fields_ = [("n_nodes", C.c_uint64), ("n_deltas", C.c_uint64), ("delta_hash", C.c_void_p), ("delta_parent", C.c_void_p),
                ("delta_child", C.c_void_p), ("node_offsets", C.c_void_p), ("parent_index", C.c_void_p), ("seed", SeedParams)]


class PlaceParams(C.Structure):
    """== placement::TraversalParams (placement.hpp:28-54), hot-path subset; defaults are the CLI defaults."""
    _fields_ = [("trim_start", C.c_int32), ("trim_end", C.c_int32), ("min_read_support", C.c_int32), ("dedup_reads", C.c_int32),
                ("force_leaf", C.c_int32), ("skip_node_index", C.c_uint32), ("seed_mask_fraction", C.c_double),
                ("want_node_scores", C.c_int32), ("min_seed_quality", C.c_int32)]

    def __init__(self, **kw):
        super().__init__()
        self.min_read_support = -1
        self.skip_node_index = PM_NONE
        for k, v in kw.items():
            setattr(self, k, v)


class IndexExtras(C.Structure):
    """== pm_index_extras: what a LiteIndex holds besides the seed deltas (ids, identicalToParent, blockRanges, substitutionMatrix)"""
    _fields_ = [("node_ids", C.POINTER(C.c_char_p)), ("identical_to_parent", C.c_void_p), ("block_ranges", C.c_void_p), ("n_blocks", C.c_uint64),
                ("substitution_matrix", C.c_void_p)]


class PlaceResult(C.Structure):
    _fields_ = [("best_score", C.c_double * 5), ("best_index", C.c_uint32 * 5), ("tied_count", C.c_uint64 * 5),
                ("total_reads", C.c_uint64), ("unique_seeds", C.c_uint64), ("read_unique_seed_count", C.c_uint64),
                ("total_read_seed_frequency", C.c_int64), ("min_read_support", C.c_int64), ("read_magnitude", C.c_double),
                ("log_containment_denominator", C.c_double), ("weighted_containment_denominator", C.c_double),
                ("stage_ms", C.c_float * 8)]


_lib = None


def lib():
    """Load libpanmap_b200.so; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise PanmapError(-2, f"{p} is missing: run __graft_entry__.build() (there is no CPU fallback)")
    L = C.CDLL(p)
    L.pm_last_error.restype = C.c_char_p
    L.pm_launch_count.restype = C.c_uint64
    L.pm_host_index_node_id.restype = C.c_char_p
    L.pm_host_index_node_id.argtypes = [C.c_void_p, C.c_uint64]
    L.pm_host_index_read.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.pm_host_index_free.argtypes = [C.c_void_p]
    L.pm_host_index_desc.argtypes = [C.c_void_p, C.POINTER(IndexDesc)]
    L.pm_host_index_extras.argtypes = [C.c_void_p, C.POINTER(IndexExtras)]
    L.pm_host_index_write.argtypes = [C.c_char_p, C.POINTER(IndexDesc), C.POINTER(IndexExtras), C.c_int, C.POINTER(C.c_uint64)]
    L.pm_index_open_cached.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    L.pm_index_image_write.argtypes = [C.POINTER(IndexDesc), C.POINTER(C.c_char_p), C.c_uint32, C.c_uint32, C.c_char_p, C.POINTER(C.c_uint64)]
    L.pm_index_create_from_image.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    L.pm_index_node_id.restype = C.c_char_p
    L.pm_index_node_id.argtypes = [C.c_void_p, C.c_uint64]
    L.pm_index_create.argtypes = [C.POINTER(IndexDesc), C.c_int, C.POINTER(C.c_void_p)]
    L.pm_index_create_shard.argtypes = [C.POINTER(IndexDesc), C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]
    L.pm_index_destroy.argtypes = [C.c_void_p]
    for f in ("pm_index_num_nodes", "pm_index_num_deltas", "pm_index_num_distinct_seeds"):
        getattr(L, f).restype = C.c_uint64
        getattr(L, f).argtypes = [C.c_void_p]
    L.pm_index_shard_range.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.pm_index_genome_metrics.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.pm_index_bfs_ranks.argtypes = [C.c_void_p, C.c_void_p]
    L.pm_workspace_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.pm_workspace_destroy.argtypes = [C.c_void_p]
    L.pm_place.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(PlaceParams), C.POINTER(PlaceResult)]
    L.pm_place_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(PlaceParams), C.POINTER(PlaceResult)]
    L.pm_packed_chunks.restype = C.c_uint64
    L.pm_packed_chunks.argtypes = [C.c_void_p, C.c_uint64]
    L.pm_pack_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
    L.pm_place_quality.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(PlaceParams), C.POINTER(PlaceResult)]
    L.pm_reads_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.pm_reads_upload_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.pm_place_resident.argtypes = [C.c_void_p, C.POINTER(PlaceParams), C.POINTER(PlaceResult)]
    L.pm_get_tied.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
    L.pm_get_node_scores.argtypes = [C.c_void_p, C.c_void_p]
    L.pm_get_node_metrics.argtypes = [C.c_void_p, C.c_void_p]
    L.pm_last_kernel_ms.argtypes = [C.c_void_p, C.c_void_p]
    L.pm_workspace_set_stage_timers.argtypes = [C.c_void_p, C.c_int]
    L.pm_get_seed_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.pm_hash_seq.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    L.pm_rolling_syncmers.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pm_read_seeds.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(SeedParams), C.c_int, C.c_int,
                                C.c_void_p, C.c_void_p]
    L.pm_stage_seed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(PlaceParams)]
    L.pm_stage_seed_resident.argtypes = [C.c_void_p, C.POINTER(PlaceParams)]
    L.pm_stage_table_size.restype = C.c_int64
    L.pm_stage_table_size.argtypes = [C.c_void_p]
    L.pm_stage_table_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.pm_stage_table_import.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.pm_stage_table_export_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.pm_stage_table_import_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.pm_stage_table_import_dev_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_uint64]
    L.pm_stage_records_export_all.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.pm_stage_score.argtypes = [C.c_void_p, C.POINTER(PlaceParams)]
    L.pm_stage_records_size.restype = C.c_int64
    L.pm_stage_records_size.argtypes = [C.c_void_p, C.c_int]
    L.pm_stage_records_export.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.pm_stage_select.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(PlaceResult)]
    L.pm_comm_unique_id.argtypes = [C.c_void_p]
    L.pm_comm_create_nccl.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.pm_comm_create_local.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]
    L.pm_comm_destroy.argtypes = [C.c_void_p]
    L.pm_place_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(PlaceParams), C.POINTER(PlaceResult)]
    L.pm_place_sharded_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(PlaceParams), C.POINTER(PlaceResult)]
    L.pm_place_sharded_resident.argtypes = [C.c_void_p, C.POINTER(PlaceParams), C.POINTER(PlaceResult)]
    L.pm_place_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(PlaceParams), C.POINTER(PlaceResult)]
    L.pm_place_multi_resident.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(PlaceParams), C.POINTER(PlaceResult)]
    L.pm_comm_last_traffic.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.pm_read_fastx.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_char_p, C.c_uint64]
    L.pm_free.argtypes = [C.c_void_p]
    L.pm_place_files.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(PlaceParams),
                                 C.POINTER(PlaceResult), C.c_char_p, C.c_uint64]
    L.pm_host_alloc.restype = C.c_void_p
    L.pm_host_alloc.argtypes = [C.c_uint64]
    L.pm_host_free.argtypes = [C.c_void_p]
    _lib = L
    return L


def _ck(rc):
    if rc < 0:
        raise PanmapError(rc, lib().pm_last_error().decode(errors="replace"))
    return rc


def launch_count():
    """kernel launches of the library in this process so far"""
    return int(lib().pm_launch_count())


def device_count():
    return lib().pm_device_count()


def pack_reads(reads):
    """list of bytes/str -> (uint8 array of concatenated bases, uint64 offsets[n+1])"""
    bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, dtype=np.uint8)
    return buf, off


def host_pack_reads(reads, offsets, threads=0):
    """ASCII reads -> the 4-bit layout of pm_place_packed (pm_pack_reads on the host); returns a 16-byte aligned uint8 array"""
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = offsets.size - 1
    chunks = int(lib().pm_packed_chunks(offsets.ctypes.data_as(C.c_void_p), n))
    raw = np.zeros(chunks * 16 + 16, dtype=np.uint8)
    shift = (-raw.ctypes.data) % 16
    out = raw[shift:shift + chunks * 16]
    _ck(lib().pm_pack_reads(reads.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p), threads))
    return out


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class HostIndex:
    """A parsed ``.idx`` (index_single_mode.cpp:1561-1636) or a set of flat arrays in reference-native widths."""

    def __init__(self, hash, parent, child, offsets, parent_index, k, s, t, l, open=0, hpc=0, node_ids=None):
        self.hash = np.ascontiguousarray(hash, dtype=np.uint64)
        self.parent = np.ascontiguousarray(parent, dtype=np.int16)
        self.child = np.ascontiguousarray(child, dtype=np.int16)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.parent_index = np.ascontiguousarray(parent_index, dtype=np.uint32)
        self.k, self.s, self.t, self.l, self.open, self.hpc = int(k), int(s), int(t), int(l), int(open), int(hpc)
        self.node_ids = node_ids
        self.n_nodes = int(self.parent_index.size)
        self.n_deltas = int(self.hash.size)
        self.identical_to_parent = None     # u8[n_nodes]
        self.block_ranges = None            # u32[n_blocks, 2]
        self.substitution_matrix = None     # f64[16]

    def write(self, path, zstd_level=-1):
        """IndexBuilder::writeIndex (index_single_mode.cpp:1593-1636): PMI1 header + LiteIndex message; zstd_level < 0 = uncompressed.
        Returns the file size."""
        d = self.desc()
        x = IndexExtras()
        keep = []
        if self.node_ids is not None:
            ids = (C.c_char_p * self.n_nodes)(*[i.encode() for i in self.node_ids])
            keep.append(ids)
            x.node_ids = C.cast(ids, C.POINTER(C.c_char_p))
        if self.identical_to_parent is not None:
            a = np.ascontiguousarray(self.identical_to_parent, np.uint8); keep.append(a); x.identical_to_parent = _ptr(a)
        if self.block_ranges is not None and len(self.block_ranges):
            a = np.ascontiguousarray(self.block_ranges, np.uint32).reshape(-1); keep.append(a); x.block_ranges = _ptr(a); x.n_blocks = a.size // 2
        if self.substitution_matrix is not None:
            a = np.ascontiguousarray(self.substitution_matrix, np.float64).reshape(-1); keep.append(a); x.substitution_matrix = _ptr(a)
        n = C.c_uint64()
        _ck(lib().pm_host_index_write(os.fsencode(path), C.byref(d), C.byref(x), int(zstd_level), C.byref(n)))
        return n.value

    @classmethod
    def read(cls, path):
        L = lib()
        h = C.c_void_p()
        _ck(L.pm_host_index_read(os.fsencode(path), C.byref(h)))
        return cls._from_handle(h)

    @classmethod
    def build_from_panman(cls, panman_path, k=19, s=8, t=0, l=3, open=0, hpc=0, flank_mask=0, device=0):
        """the product's own index builder (pm_index_build; reference IndexBuilder::buildIndex, index_single_mode.cpp:1227-1392): every
        node genome of the .panman seeded on the GPU and diffed against its parent's; == panmap --flank-mask 0 delta for delta"""
        L = lib()
        L.pm_index_build.argtypes = [C.c_char_p, C.POINTER(SeedParams), C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        sp = SeedParams(int(k), int(s), int(t), int(l), int(open), int(hpc))
        h = C.c_void_p()
        _ck(L.pm_index_build(os.fsencode(panman_path), C.byref(sp), int(flank_mask), int(device), C.byref(h)))
        return cls._from_handle(h)

    @classmethod
    def _from_handle(cls, h):
        L = lib()
        try:
            d = IndexDesc()
            _ck(L.pm_host_index_desc(h, C.byref(d)))
            N, D = d.n_nodes, d.n_deltas

            def arr(p, n, dt):
                if n == 0:
                    return np.zeros(0, dtype=dt)
                return np.ctypeslib.as_array(C.cast(p, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,)).copy()
            ids = [L.pm_host_index_node_id(h, i).decode() for i in range(N)]
            out = cls(arr(d.delta_hash, D, np.uint64), arr(d.delta_parent, D, np.int16), arr(d.delta_child, D, np.int16),
                      arr(d.node_offsets, N + 1, np.uint64), arr(d.parent_index, N, np.uint32),
                      d.seed.k, d.seed.s, d.seed.t, d.seed.l, d.seed.open, d.seed.hpc, ids)
            x = IndexExtras()
            _ck(L.pm_host_index_extras(h, C.byref(x)))
            if x.identical_to_parent:
                out.identical_to_parent = arr(x.identical_to_parent, N, np.uint8)
            if x.block_ranges:
                out.block_ranges = arr(x.block_ranges, 2 * x.n_blocks, np.uint32).reshape(-1, 2)
            if x.substitution_matrix:
                out.substitution_matrix = arr(x.substitution_matrix, 16, np.float64)
            return out
        finally:
            L.pm_host_index_free(h)

    def desc(self):
        d = IndexDesc()
        d.n_nodes, d.n_deltas = self.n_nodes, self.n_deltas
        d.delta_hash, d.delta_parent, d.delta_child = _ptr(self.hash), _ptr(self.parent), _ptr(self.child)
        d.node_offsets, d.parent_index = _ptr(self.offsets), _ptr(self.parent_index)
        d.seed = SeedParams(self.k, self.s, self.t, self.l, self.open, self.hpc)
        return d


class Index:
    """Flattened index resident in HBM (replaces LiteTree + the SoA hookup of placement.cpp:1021-1092)."""

    def __init__(self, host, device=0, shard=0, n_shards=1):
        self.host = host
        self._h = C.c_void_p()
        self.cache_hit = None
        if host is None:
            return
        d = host.desc()
        _ck(lib().pm_index_create_shard(C.byref(d), device, shard, n_shards, C.byref(self._h)))
        self.n_nodes = host.n_nodes

    @classmethod
    def open_cached(cls, idx_path, image_path=None, device=0, shard=0, n_shards=1):
        """open a .idx through its cached flattened image (written on a miss); .cache_hit says which way it went"""
        self = cls(None)
        hit = C.c_int()
        _ck(lib().pm_index_open_cached(os.fsencode(idx_path), os.fsencode(image_path) if image_path else None, device, shard, n_shards,
                                       C.byref(self._h), C.byref(hit)))
        self.cache_hit = bool(hit.value)
        self.n_nodes = int(lib().pm_index_num_nodes(self._h))
        return self

    @classmethod
    def from_image(cls, image_path, device=0):
        self = cls(None)
        _ck(lib().pm_index_create_from_image(os.fsencode(image_path), device, C.byref(self._h)))
        self.n_nodes = int(lib().pm_index_num_nodes(self._h))
        return self

    def node_id(self, i):
        return lib().pm_index_node_id(self._h, int(i)).decode()

    def node_ids(self):
        """LiteNode ids: the host index's when there is one, else the ones that travelled with the file / image (None when there are none)"""
        if self.host is not None:
            return self.host.node_ids
        if not hasattr(self, "_ids"):
            self._ids = [self.node_id(i) for i in range(self.n_nodes)] if self.n_nodes and self.node_id(0) else None
        return self._ids

    def close(self):
        if self._h:
            lib().pm_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_distinct_seeds(self):
        return lib().pm_index_num_distinct_seeds(self._h)

    def shard_range(self):
        b, e = C.c_uint64(), C.c_uint64()
        _ck(lib().pm_index_shard_range(self._h, C.byref(b), C.byref(e)))
        return b.value, e.value

    def genome_metrics(self):
        m = np.zeros(self.n_nodes, dtype=np.float64)
        u = np.zeros(self.n_nodes, dtype=np.int64)
        _ck(lib().pm_index_genome_metrics(self._h, _ptr(m), _ptr(u)))
        return m, u

    def bfs_ranks(self):
        r = np.zeros(self.n_nodes, dtype=np.uint32)
        _ck(lib().pm_index_bfs_ranks(self._h, _ptr(r)))
        return r


class Placement:
    """== placement::PlacementResult (placement.hpp:157-235): scores, best node, tied nodes per metric + read stats."""

    def __init__(self, res, tied, node_ids=None):
        self.raw = res
        self.best_score = {m: res.best_score[i] for i, m in enumerate(METRICS)}
        self.best_index = {m: res.best_index[i] for i, m in enumerate(METRICS)}
        self.tied = {m: tied[i] for i, m in enumerate(METRICS)}
        self.node_ids = node_ids
        self.stage_ms = list(res.stage_ms)

    def tsv(self):
        """the <prefix>.placement.tsv the reference writes (placement.cpp:1952-1985)"""
        def name(i):
            return self.node_ids[i] if self.node_ids is not None and i < len(self.node_ids) else ""
        lines = ["metric\tscore\tnodes"]
        for m in METRICS:
            t = self.tied[m]
            nodes = ",".join(name(int(i)) for i in t) if len(t) else (name(self.best_index[m]) if self.best_index[m] != PM_NONE else "")
            lines.append(f"{m}\t{self.best_score[m]:.6f}\t{nodes}")
        return "\n".join(lines) + "\n"


class Workspace:
    """Per-sample state (one CUDA stream); ``place`` == the compute part of placement::placeLite."""

    def __init__(self, index):
        self.index = index
        self._h = C.c_void_p()
        _ck(lib().pm_workspace_create(index._h, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().pm_workspace_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _finish(self, res):
        tied = []
        for m in range(5):
            n = int(res.tied_count[m])
            t = np.zeros(n, dtype=np.uint32)
            if n:
                _ck(lib().pm_get_tied(self._h, m, _ptr(t), n))
            tied.append(t)
        return Placement(res, tied, self.index.node_ids())

    def place(self, reads, offsets, params=None):
        params = params or PlaceParams()
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        res = PlaceResult()
        _ck(lib().pm_place(self._h, _ptr(reads), offsets.ctypes.data_as(C.c_void_p), offsets.size - 1, C.byref(params), C.byref(res)))
        return self._finish(res)

    def place_quality(self, reads, quals, offsets, params):
        """--min-seed-quality: quals = one Phred+33 byte per base at the reads' offsets (pm_place_quality)"""
        reads = np.ascontiguousarray(reads, dtype=np.uint8); quals = np.ascontiguousarray(quals, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        if quals.size != reads.size:
            raise ValueError("quals must have one byte per base")
        res = PlaceResult()
        _ck(lib().pm_place_quality(self._h, _ptr(reads), _ptr(quals), offsets.ctypes.data_as(C.c_void_p), offsets.size - 1, C.byref(params),
                                   C.byref(res)))
        return self._finish(res)

    def place_packed(self, packed, offsets, params=None):
        """reads as 4-bit codes (host_pack_reads / a packing parser): uint8 array of 16-byte chunks, 16-byte aligned"""
        params = params or PlaceParams()
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        if packed.ctypes.data % 16:
            raise ValueError("packed reads must be 16-byte aligned")
        res = PlaceResult()
        _ck(lib().pm_place_packed(self._h, packed.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p), offsets.size - 1, C.byref(params), C.byref(res)))
        return self._finish(res)

    def place_packed_raw(self, packed_ptr, offsets_ptr, n_reads, params):
        res = PlaceResult()
        _ck(lib().pm_place_packed(self._h, packed_ptr, offsets_ptr, n_reads, C.byref(params), C.byref(res)))
        return res

    def place_raw(self, reads_ptr, offsets_ptr, n_reads, params):
        """host pointers (e.g. pinned buffers); returns the PlaceResult struct only"""
        res = PlaceResult()
        _ck(lib().pm_place(self._h, reads_ptr, offsets_ptr, n_reads, C.byref(params), C.byref(res)))
        return res

    def upload(self, reads, offsets):
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        _ck(lib().pm_reads_upload(self._h, _ptr(reads), offsets.ctypes.data_as(C.c_void_p), offsets.size - 1))

    def upload_device(self, d_reads_ptr, d_offsets_ptr, h_offsets):
        """sample already in HBM (device pointers) -> this workspace; h_offsets: the same offsets as a host array"""
        h_offsets = np.ascontiguousarray(h_offsets, dtype=np.uint64)
        _ck(lib().pm_reads_upload_device(self._h, d_reads_ptr, d_offsets_ptr, h_offsets.ctypes.data_as(C.c_void_p), h_offsets.size - 1))

    def place_resident(self, params=None, full=True):
        params = params or PlaceParams()
        res = PlaceResult()
        _ck(lib().pm_place_resident(self._h, C.byref(params), C.byref(res)))
        return self._finish(res) if full else res

    def stage_timers(self, on=True):
        """CUDA events between the stages of a placement (stage_ms[0..6], last_kernel_ms): off by default, ~25 us per placement when on"""
        _ck(lib().pm_workspace_set_stage_timers(self._h, 1 if on else 0))

    def last_kernel_ms(self):
        """CUDA-event times (ms) of pack_reads, syncmers_*, count_seeds of the last place_resident call"""
        out = (C.c_float * 3)()
        _ck(lib().pm_last_kernel_ms(self._h, out))
        return [float(x) for x in out]

    def node_scores(self):
        out = np.zeros((self.index.n_nodes, 5), dtype=np.float64)
        _ck(lib().pm_get_node_scores(self._h, _ptr(out)))
        return out

    def node_metrics(self):
        out = np.zeros((self.index.n_nodes, 5), dtype=np.float64)
        _ck(lib().pm_get_node_metrics(self._h, _ptr(out)))
        return out

    def seed_table(self, cap=None):
        cap = int(cap or (1 << 22))
        while True:
            h = np.zeros(cap, dtype=np.uint64)
            c = np.zeros(cap, dtype=np.int64)
            n = _ck(lib().pm_get_seed_table(self._h, _ptr(h), _ptr(c), cap))
            if n <= cap:
                o = np.argsort(h[:n], kind="stable")
                return h[:n][o], c[:n][o]
            cap = n

    # ---- staged multi-GPU protocol (see include/panmap_b200.h) ----
    def stage_seed(self, reads, offsets, params):
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        _ck(lib().pm_stage_seed(self._h, _ptr(reads), offsets.ctypes.data_as(C.c_void_p), offsets.size - 1, C.byref(params)))

    def stage_seed_resident(self, params):
        _ck(lib().pm_stage_seed_resident(self._h, C.byref(params)))

    def stage_table_export(self):
        n = _ck(lib().pm_stage_table_size(self._h))
        h = np.zeros(max(n, 1), dtype=np.uint64)
        c = np.zeros(max(n, 1), dtype=np.int64)
        if n:
            _ck(lib().pm_stage_table_export(self._h, _ptr(h), _ptr(c), n))
        return h[:n], c[:n]

    def stage_table_import(self, h, c):
        h = np.ascontiguousarray(h, dtype=np.uint64)
        c = np.ascontiguousarray(c, dtype=np.int64)
        _ck(lib().pm_stage_table_import(self._h, _ptr(h), _ptr(c), h.size))

    def stage_table_export_dev(self, d_hash_ptr, d_count_ptr, cap):
        n = C.c_uint64()
        _ck(lib().pm_stage_table_export_dev(self._h, d_hash_ptr, d_count_ptr, cap, C.byref(n)))
        return n.value

    def stage_table_import_dev(self, d_hash_ptr, d_count_ptr, n):
        _ck(lib().pm_stage_table_import_dev(self._h, d_hash_ptr, d_count_ptr, n))

    def stage_table_import_dev_async(self, d_hash_ptr, d_count_ptr, n, clear_first=True, expected_total=None):
        _ck(lib().pm_stage_table_import_dev_async(self._h, d_hash_ptr, d_count_ptr, n, int(clear_first), n if expected_total is None else expected_total))

    def stage_records_all(self, cap=512):
        """all five record lists with one device synchronisation (buffers are kept on the object)"""
        while True:
            b = getattr(self, "_recbuf", None)
            if b is None or b[0] != cap:
                b = (cap, np.zeros(5, np.uint32), np.zeros((5, cap), np.uint32), np.zeros((5, cap), np.uint32), np.zeros((5, cap), np.float64))
                self._recbuf = b
            _, cnt, r, v, s = b
            _ck(lib().pm_stage_records_export_all(self._h, _ptr(cnt), _ptr(r), _ptr(v), _ptr(s), cap))
            if int(cnt.max()) < cap:
                return [(r[m, :cnt[m]].copy(), v[m, :cnt[m]].copy(), s[m, :cnt[m]].copy()) for m in range(5)]
            cap = max(_ck(lib().pm_stage_records_size(self._h, m)) for m in range(5)) + 1

    def stage_score(self, params):
        _ck(lib().pm_stage_score(self._h, C.byref(params)))

    def stage_records(self):
        out = []
        for m in range(5):
            n = _ck(lib().pm_stage_records_size(self._h, m))
            r = np.zeros(max(n, 1), dtype=np.uint32)
            v = np.zeros(max(n, 1), dtype=np.uint32)
            s = np.zeros(max(n, 1), dtype=np.float64)
            if n:
                _ck(lib().pm_stage_records_export(self._h, m, _ptr(r), _ptr(v), _ptr(s), n))
            out.append((r[:n], v[:n], s[:n]))
        return out

    def stage_select(self, records, total_reads):
        """records: per metric (rank, node, score) arrays holding ALL ranks' records"""
        counts = np.array([len(r[0]) for r in records], dtype=np.uint32)
        keep = [tuple(np.ascontiguousarray(a) for a in r) for r in records]
        P = C.c_void_p * 5
        pr = P(*[a[0].ctypes.data if a[0].size else None for a in keep])
        pn = P(*[a[1].ctypes.data if a[1].size else None for a in keep])
        ps = P(*[a[2].ctypes.data if a[2].size else None for a in keep])
        res = PlaceResult()
        _ck(lib().pm_stage_select(self._h, _ptr(counts), pr, pn, ps, total_reads, C.byref(res)))
        return self._finish(res)


COMM_ID_BYTES = 128


def comm_unique_id():
    """the 128-byte id rank 0 creates and hands to the other ranks (ncclGetUniqueId behind pm_comm_unique_id)"""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _ck(lib().pm_comm_unique_id(buf))
    return buf.raw


class Comm:
    """One rank of a group of GPUs that place ONE sample together (pm_comm): the workspace sits on shard `rank` of the node range."""

    def __init__(self, handle, workspace, rank, n_ranks):
        self._h, self.ws, self.rank, self.n_ranks = handle, workspace, rank, n_ranks

    @classmethod
    def nccl(cls, workspace, unique_id, rank, n_ranks):
        """one process per GPU; NCCL on the workspace stream"""
        h = C.c_void_p()
        _ck(lib().pm_comm_create_nccl(workspace._h, unique_id, rank, n_ranks, C.byref(h)))
        return cls(h, workspace, rank, n_ranks)

    @classmethod
    def local(cls, workspaces):
        """one process, one host thread, n workspaces (any devices): returns the n communicators in rank order"""
        n = len(workspaces)
        hs = (C.c_void_p * n)(*[w._h for w in workspaces])
        out = (C.c_void_p * n)()
        _ck(lib().pm_comm_create_local(hs, n, out))
        return [cls(C.c_void_p(out[r]), workspaces[r], r, n) for r in range(n)]

    def transport(self):
        lib().pm_comm_transport.restype = C.c_char_p
        return lib().pm_comm_transport(self._h).decode()

    def close(self):
        if self._h:
            lib().pm_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def place_sharded(self, reads, offsets, params=None, full=True):
        """NCCL transport: this rank's slice of the reads (offsets start at 0); every rank gets the same Placement"""
        params = params or PlaceParams()
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        res = PlaceResult()
        _ck(lib().pm_place_sharded(self._h, _ptr(reads), offsets.ctypes.data_as(C.c_void_p), offsets.size - 1, C.byref(params), C.byref(res)))
        return self.ws._finish(res) if full else res

    def place_sharded_raw(self, reads_ptr, offsets_ptr, n_reads, params):
        res = PlaceResult()
        _ck(lib().pm_place_sharded(self._h, reads_ptr, offsets_ptr, n_reads, C.byref(params), C.byref(res)))
        return res

    def place_sharded_packed_raw(self, packed_ptr, offsets_ptr, n_reads, params):
        res = PlaceResult()
        _ck(lib().pm_place_sharded_packed(self._h, packed_ptr, offsets_ptr, n_reads, C.byref(params), C.byref(res)))
        return res

    def place_sharded_resident(self, params=None, full=True):
        params = params or PlaceParams()
        res = PlaceResult()
        _ck(lib().pm_place_sharded_resident(self._h, C.byref(params), C.byref(res)))
        return self.ws._finish(res) if full else res

    def last_traffic(self):
        a, b = C.c_uint64(), C.c_uint64()
        _ck(lib().pm_comm_last_traffic(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value


def place_multi(comms, reads, offsets, params=None):
    """local transport: the whole sample in, sliced over the ranks; returns rank 0's Placement (all ranks hold the same)"""
    params = params or PlaceParams()
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    hs = (C.c_void_p * len(comms))(*[c._h for c in comms])
    res = PlaceResult()
    _ck(lib().pm_place_multi(hs, len(comms), _ptr(reads), offsets.ctypes.data_as(C.c_void_p), offsets.size - 1, C.byref(params), C.byref(res)))
    return comms[0].ws._finish(res)


def place_multi_resident(comms, params=None, full=True):
    params = params or PlaceParams()
    hs = (C.c_void_p * len(comms))(*[c._h for c in comms])
    res = PlaceResult()
    _ck(lib().pm_place_multi_resident(hs, len(comms), C.byref(params), C.byref(res)))
    return comms[0].ws._finish(res) if full else res


def hash_seq(seqs, device=0):
    """GPU seeding::hashSeq for a batch of k-mers -> (forward[], reverse[]) uint64 arrays; raises on non-ACGT like the reference"""
    buf, off = pack_reads(seqs)
    f = np.zeros(len(seqs), np.uint64); r = np.zeros(len(seqs), np.uint64)
    _ck(lib().pm_hash_seq(device, _ptr(buf), off.ctypes.data_as(C.c_void_p), len(seqs), f.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p)))
    return f, r


def rolling_syncmers(seqs, k, s, open=False, t=0, device=0):
    """GPU seeding::rollingSyncmers(seq,k,s,open,t,returnAll=false) for a batch: list of (hash[], is_reverse[], pos[])."""
    buf, off = pack_reads(seqs)
    n = len(seqs)
    lens = np.diff(off.astype(np.int64))
    win = np.maximum(lens - k + 1, 0)
    woff = np.concatenate([[0], np.cumsum(win)]).astype(np.int64)
    tot = int(woff[-1])
    h = np.zeros(max(tot, 1), dtype=np.uint64)
    r = np.zeros(max(tot, 1), dtype=np.uint8)
    p = np.zeros(max(tot, 1), dtype=np.int64)
    c = np.zeros(max(n, 1), dtype=np.uint64)
    _ck(lib().pm_rolling_syncmers(device, _ptr(buf), off.ctypes.data_as(C.c_void_p), n, k, s, int(bool(open)), t,
                                  _ptr(h), _ptr(r), _ptr(p), _ptr(c)))
    return [(h[woff[i]:woff[i] + int(c[i])], r[woff[i]:woff[i] + int(c[i])], p[woff[i]:woff[i] + int(c[i])]) for i in range(n)]


def read_seeds(seqs, k, s, t, l, open=False, trim_start=0, trim_end=0, device=0):
    """per-read seeds exactly as placeLite counts them (placement.cpp:1598-1686): list of hash arrays"""
    buf, off = pack_reads(seqs)
    n = len(seqs)
    lens = np.diff(off.astype(np.int64))
    win = np.maximum(lens - k + 1, 0)
    woff = np.concatenate([[0], np.cumsum(win)]).astype(np.int64)
    tot = int(woff[-1])
    h = np.zeros(max(tot, 1), dtype=np.uint64)
    c = np.zeros(max(n, 1), dtype=np.uint64)
    sp = SeedParams(k, s, t, l, int(bool(open)), 0)
    _ck(lib().pm_read_seeds(device, _ptr(buf), off.ctypes.data_as(C.c_void_p), n, C.byref(sp), trim_start, trim_end, _ptr(h), _ptr(c)))
    return [h[woff[i]:woff[i] + int(c[i])] for i in range(n)]


def panman_genomes(panman_path, coords=False):
    """every node's ungapped genome of a .panman in DFS pre-order (pm_panman_genomes; the reference's getStringFromReference,
    panmap_utils.cpp:7-193): (uint8 bases, uint64 offsets[n + 1], uint32 parent_index[n], [ids])"""
    L = lib()
    L.pm_panman_genomes.argtypes = [C.c_char_p] + [C.POINTER(C.c_void_p)] * 5 + [C.POINTER(C.c_uint64)]
    b, o, pi, ids, co, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint64()
    _ck(L.pm_panman_genomes(os.fsencode(panman_path), C.byref(b), C.byref(o), C.byref(pi), C.byref(ids), C.byref(co) if coords else None, C.byref(n)))
    try:
        off = np.ctypeslib.as_array(C.cast(o, C.POINTER(C.c_uint64)), shape=(n.value + 1,)).copy()
        tot = int(off[-1])
        bases = np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_uint8)), shape=(max(tot, 1),))[:tot].copy()
        par = np.ctypeslib.as_array(C.cast(pi, C.POINTER(C.c_uint32)), shape=(max(n.value, 1),))[:n.value].copy()
        names = C.string_at(ids).decode().split("\n")[:n.value]
        cc = np.ctypeslib.as_array(C.cast(co, C.POINTER(C.c_uint32)), shape=(max(tot, 1),))[:tot].copy() if coords else None
    finally:
        for q in (b, o, pi, ids, co):
            if q:
                L.pm_free(q)
    return (bases, off, par, names, cc) if coords else (bases, off, par, names)


def read_fastx(reads1, reads2=""):
    """extractReadSequences (placement.cpp:164-197) of the C++ host shim: (uint8 bases, uint64 offsets)"""
    b, o, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
    err = C.create_string_buffer(512)
    rc = lib().pm_read_fastx(os.fsencode(reads1), os.fsencode(reads2), C.byref(b), C.byref(o), C.byref(n), err, 512)
    if rc < 0:
        raise PanmapError(rc, err.value.decode(errors="replace"))
    off = np.ctypeslib.as_array(C.cast(o, C.POINTER(C.c_uint64)), shape=(n.value + 1,)).copy()
    tot = int(off[-1])
    buf = np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_uint8)), shape=(max(tot, 1),))[:tot].copy()
    lib().pm_free(b); lib().pm_free(o)
    return buf, off


def place_files(workspace, reads1, reads2="", out_tsv="", params=None):
    """placement::placeLite through the C++ host shim (files in, TSV out); returns the PlaceResult struct"""
    params = params or PlaceParams()
    ids = workspace.index.node_ids() or []
    arr = (C.c_char_p * max(len(ids), 1))(*[i.encode() for i in ids])
    res = PlaceResult()
    err = C.create_string_buffer(512)
    rc = lib().pm_place_files(workspace.index._h, workspace._h, arr, len(ids), os.fsencode(reads1), os.fsencode(reads2), os.fsencode(out_tsv),
                              C.byref(params), C.byref(res), err, 512)
    if rc < 0:
        raise PanmapError(rc, err.value.decode(errors="replace"))
    return res
