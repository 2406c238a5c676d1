#!/usr/bin/env python3
"""cuobjdump -sass of the named kernels of libpanmap_b200.so -> profiles/sass/<tag>_<kernel>.sass (encodings stripped) plus an
opcode histogram per kernel (profiles/sass/<tag>_mix.txt).  usage: tools/sass_dump.py <tag> <regex> [<regex> ...]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(tag, pats):
    so = os.path.join(ROOT, "panmap_b200", "libpanmap_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    dem = {}
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    out_dir = os.path.join(ROOT, "profiles", "sass")
    os.makedirs(out_dir, exist_ok=True)
    mix_lines = []
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        d = dem.get(name) or subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        short = d.split("(")[0].replace("void ", "").replace("pm::", "")
        if not any(re.search(p, short) for p in pats):
            continue
        ins = []
        for line in f.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", line)
            if m:
                ins.append((m.group(1), m.group(2).strip()))
        fn = re.sub(r"[^A-Za-z0-9_]+", "_", short).strip("_")
        with open(os.path.join(out_dir, f"{tag}_{fn}.sass"), "w") as o:
            o.write(f"// {d}\n")
            for a, i in ins:
                o.write(f"/*{a}*/ {i}\n")
        ops = collections.Counter()
        for _, i in ins:
            t = i.split()
            op = t[1] if t[0].startswith("@") else t[0]
            ops[op.split(".")[0]] += 1
        mix_lines.append(f"== {short}: {len(ins)} instructions (static)")
        mix_lines.append("   " + ", ".join(f"{k} {v}" for k, v in ops.most_common(28)))
    with open(os.path.join(out_dir, f"{tag}_mix.txt"), "w") as o:
        o.write("\n".join(mix_lines) + "\n")
    print("\n".join(mix_lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
