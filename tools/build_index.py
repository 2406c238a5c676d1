#!/usr/bin/env python3
"""`.panman` -> `.idx` with the product's own builder (pm_index_build: genomes materialised, seeded, sorted and diffed on the GPU;
== panmap --flank-mask 0).  usage: tools/build_index.py <in.panman> <out.idx> [k s t l] [--zstd LEVEL]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panmap_b200 as pm


def main(argv):
    args = [a for a in argv if not a.startswith("--")]
    zstd = int(argv[argv.index("--zstd") + 1]) if "--zstd" in argv else -1
    if "--zstd" in argv:
        args.remove(str(zstd))
    if len(args) < 2:
        print(__doc__); return 2
    k, s, t, l = ([int(x) for x in args[2:6]] + [19, 8, 0, 3][len(args[2:6]):])[:4]
    t0 = time.perf_counter()
    idx = pm.HostIndex.build_from_panman(args[0], k=k, s=s, t=t, l=l)
    t1 = time.perf_counter()
    n = idx.write(args[1], zstd_level=zstd)
    print(f"{idx.n_nodes} nodes, {idx.n_deltas} seed deltas (k={k} s={s} t={t} l={l}); built in {t1 - t0:.2f} s, {n} bytes written to {args[1]}")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
