#!/usr/bin/env python3
"""Key metrics per kernel from an .ncu-rep (ncu --set full): duration, DRAM bytes, issue/occupancy, top stalls."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "smsp__average_warp_latency_per_inst_issued.ratio"]


def traffic_json(paths, dest):
    """profiles/ncu_traffic.json: dram__bytes_read + dram__bytes_write per launch (mean over the captured launches) of every kernel in the
    given --set full captures; bench.py reads it for roofline.traffic"""
    import json
    import os
    acc = {}
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            name = r[ki].split("(")[0].replace("void ", "").split("<")[0].strip()
            b = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
            acc.setdefault(name, []).append(b)
    d = {"capture": ", ".join(os.path.basename(p) for p in paths), "unit": "bytes per launch (dram read + write, ncu --set full)",
         "kernels": {k: sum(v) / len(v) for k, v in acc.items()}}
    json.dump(d, open(dest, "w"), indent=1)
    print(json.dumps(d))


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("==", r[ki].split("(")[0])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:62s} {r[i]:>16s} {units[i]}")
        st = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[i]]
        st.sort(reverse=True)
        print("   top stalls (warps per issue):", ", ".join(f"{h.split('stalled_')[1].split('_per')[0]}={v:.2f}" for v, h in st[:6]))


if __name__ == "__main__":
    if sys.argv[1] == "--json":      # tools/ncu_summary.py --json profiles/ncu_traffic.json a.ncu-rep [b.ncu-rep ...]
        traffic_json(sys.argv[3:], sys.argv[2])
    else:
        main(sys.argv[1])
