#!/usr/bin/env python3
"""Dynamic per-function breakdown of one kernel from an `ncu --set full --import-source on` report, without a GPU.

usage: tools/ncu_regions.py REPORT.ncu-rep LIB.so KERNEL_SUBSTRING [--lines]
The report's SASS page gives, per instruction, warp-instructions executed, thread-instructions executed and stall samples; the
library (the SAME build that was profiled) gives the source line of every instruction (nvdisasm --print-line-info, innermost
inlined line).  Instructions are matched by their offset inside the kernel, then grouped by the C++ function that contains the
line (scanned from the sources).  Output: share of warp-instructions, average active lanes, share of stall samples per function.
"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "forge3d_b200" / "csrc"


def function_map(path):
    """line -> name of the function whose definition most recently started at column 0 / template / __device__ prefix."""
    out, cur = {}, "?"
    pat = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__|__host__|inline|auto)\b.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(")
    lines = path.read_text().splitlines()
    pending_template = False
    for i, ln in enumerate(lines, 1):
        m = pat.match(ln)
        if m and not ln.rstrip().endswith(";"):
            cur = m.group(1)
        m2 = re.match(r"\s*auto\s+([a-z_]+)\s*=\s*\[&\]", ln)
        if m2:
            cur = cur.split("::")[0] + "::" + m2.group(1)
        out[i] = cur
    return out


def sass_lines(lib, kernel):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", str(Path(lib).resolve())], cwd=d, check=True, stdout=subprocess.DEVNULL)
        text = "".join(subprocess.run(["nvdisasm", "--print-line-info", str(c)], stdout=subprocess.PIPE, text=True, check=True).stdout
                       for c in sorted(Path(d).glob("*.cubin")))
    res, in_k, cur = {}, False, ("?", 0)
    for line in text.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+)", line)
        if m:
            in_k = kernel in m.group(1)
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", line)
        if m:
            res[int(m.group(1), 16)] = (cur, m.group(2))
    return res


def main():
    rep, lib, kernel = sys.argv[1:4]
    by_line = "--lines" in sys.argv
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    # the page is a sequence of per-kernel blocks: "Kernel Name",<name> / header / instructions
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and r and r[0].startswith("0x"):
            cur["rows"].append(r)
    mangled_hint = kernel
    blk = [b for b in blocks if all(tok in b["name"].replace(" ", "") for tok in re.split(r"[ ,]+", sys.argv[4] if len(sys.argv) > 4 and not sys.argv[4].startswith("--") else ""))]
    blk = [b for b in blk if b["rows"]]
    want = [b for b in blk if (sys.argv[5] if len(sys.argv) > 5 and not sys.argv[5].startswith("--") else "") in b["name"]]
    b = (want or blk)[0]
    hdr = b["hdr"]
    ia, ii, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    base = int(b["rows"][0][ia], 16)
    lines = sass_lines(lib, mangled_hint)
    fmaps = {p.name: function_map(p) for p in CSRC.glob("f3d_*.cuh")}
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in b["rows"]:
        off = int(r[ia], 16) - base
        (f, ln), _ = lines.get(off, (("?", 0), ""))
        key = f"{f}:{ln}" if by_line else (fmaps.get(f, {}).get(ln, f) if f in fmaps else ("IEEE div/sqrt/rcp + intrinsics" if "intrinsics" in f or "math" in f or "device_functions" in f else f))
        vals = [int(r[ii].replace(",", "") or 0), int(r[it].replace(",", "") or 0), int(r[isamp].replace(",", "") or 0)]
        for k in range(3):
            agg[key][k] += vals[k]
            tot[k] += vals[k]
    print(f"{b['name']}: {tot[0] / 1e6:.1f} M warp-instr, {tot[1] / max(tot[0], 1):.1f} lanes average, {tot[2]} stall samples")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[: (60 if by_line else 30)]:
        if v[0] == 0:
            continue
        print(f"  {key:44s} {v[0] / 1e6:8.1f} M ({100 * v[0] / tot[0]:5.1f} %)  {v[1] / v[0]:5.1f} lanes  {100 * v[2] / max(tot[2], 1):5.1f} % of samples")


if __name__ == "__main__":
    main()
