#!/usr/bin/env python3
"""Per-CUDA-source-line instruction and stall-sample shares of one kernel from an .ncu-rep (--import-source on, -lineinfo)."""
import csv
import io
import subprocess
import sys


def main(rep, kernel, skip="0"):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kernel,
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    lines = []
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            ci = hdr.index("Instructions Executed"); si = hdr.index("# Samples")
            continue
        if hdr and len(r) > ci and r[0] not in ("", "File Path", "Function Name"):
            try:
                lines.append((int(r[0]), float(r[ci] or 0), float(r[si] or 0), r[1]))
            except ValueError:
                pass
    ti = sum(l[1] for l in lines) or 1
    ts = sum(l[2] for l in lines) or 1
    print(f"total warp instructions {ti:.0f}, samples {ts:.0f}")
    for ln, n, s, src in lines:
        if n > 0.005 * ti or s > 0.01 * ts:
            print(f"{ln:5d} inst {100*n/ti:5.1f}%  stall {100*s/ts:5.1f}%  {src.strip()[:140]}")


if __name__ == "__main__":
    main(*sys.argv[1:])
