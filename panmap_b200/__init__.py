"""panmap_b200 -- B200-native placement hot path of panmap (closed-syncmer seeding + per-node seed-delta scoring).

The product is the CUDA shared library ``libpanmap_b200.so`` (C ABI in ``include/panmap_b200.h``); this package is a
thin ctypes binding used by the tests, ``bench.py`` and ``__graft_entry__``.  There is no CPU fallback: every compute
call raises :class:`PanmapError` when the library or a CUDA device is missing.
"""
from .api import (  # noqa: F401
    PanmapError, METRICS, lib, lib_path, build, device_count, launch_count, HostIndex, Index, Workspace, PlaceParams, PlaceResult,
    hash_seq, rolling_syncmers, read_seeds, pack_reads, read_fastx, panman_genomes, place_files, Comm, comm_unique_id, host_pack_reads, place_multi, place_multi_resident,
)
