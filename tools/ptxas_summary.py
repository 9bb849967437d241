#!/usr/bin/env python3
"""Static resource table of every kernel in libforge3d_b200.so: registers, stack, spills, static shared memory (ptxas -v) and the
SASS instruction count (cuobjdump).  No GPU needed.  usage: python tools/ptxas_summary.py > profiles/rNN_static_kernels.txt"""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    from forge3d_b200 import build as b

    cmd = [b._nvcc(), "-Xptxas", "-v", *b.NVCC_FLAGS, "-o", "/tmp/_f3d_ptxas.so", *[str(b.CSRC / n) for n in b.SOURCES]]
    txt = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, check=True).stdout
    rows, cur = [], None
    for line in txt.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = {"fn": m.group(1)}
            rows.append(cur)
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and cur is not None and "stack" not in cur:
            cur.update(stack=int(m.group(1)), st=int(m.group(2)), ld=int(m.group(3)))
            continue
        m = re.search(r"Used (\d+) registers", line)
        if m and cur is not None:
            cur["regs"] = int(m.group(1))
            sm = re.search(r"(\d+) bytes smem", line)
            cur["smem"] = int(sm.group(1)) if sm else 0
    sass = subprocess.run(["cuobjdump", "-sass", "/tmp/_f3d_ptxas.so"], stdout=subprocess.PIPE, text=True, check=True).stdout
    counts, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts[name] = 0
        elif name and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
            counts[name] += 1
    print(f"{'kernel':46s} {'regs':>5s} {'stack B':>8s} {'spills st/ld':>13s} {'static smem B':>14s} {'SASS instr':>11s}")
    for r in rows:
        r["name"] = subprocess.run(["c++filt", r["fn"]], capture_output=True, text=True).stdout.strip().split("(")[0].replace("f3d::", "")
    for r in sorted(rows, key=lambda r: r["name"]):
        print(f"{r['name']:46s} {r.get('regs', 0):5d} {r.get('stack', 0):8d} {str(r.get('st', 0)) + '/' + str(r.get('ld', 0)):>13s} "
              f"{r.get('smem', 0):14d} {counts.get(r['fn'], 0):11d}")


if __name__ == "__main__":
    main()
