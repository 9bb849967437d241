#!/usr/bin/env python3
"""Static SASS attribution: instructions of one kernel grouped by the source line (-lineinfo) they were generated from.

usage: tools/sass_lines.py LIB.so KERNEL_SUBSTRING [FILE_SUBSTRING]
Needs no GPU (cuobjdump + nvdisasm).  Inlined code is attributed to the innermost source line, which is what one wants for
"how many instructions does expand_top cost per call site".  Complements profiles/ncu_summary.py (dynamic counts from ncu).
"""
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path


def main():
    lib, kernel = sys.argv[1], sys.argv[2]
    only = sys.argv[3] if len(sys.argv) > 3 else ""
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", str(Path(lib).resolve())], cwd=d, check=True, stdout=subprocess.DEVNULL)
        text = "".join(subprocess.run(["nvdisasm", "--print-line-info", str(c)], stdout=subprocess.PIPE, text=True, check=True).stdout
                       for c in sorted(Path(d).glob("*.cubin")))   # one cubin per translation unit
    counts = collections.Counter()
    in_kernel, cur = False, ("?", 0)
    for line in text.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+)", line)
        if m:
            in_kernel = kernel in m.group(1)
            continue
        if not in_kernel:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
            counts[cur] += 1
    total = sum(counts.values())
    print(f"# {kernel}: {total} instructions")
    per_file = collections.Counter()
    for (f, ln), n in counts.items():
        per_file[f] += n
    for f, n in per_file.most_common():
        print(f"# {f}: {n}")
    for (f, ln), n in sorted(counts.items()):
        if only in f:
            print(f"{f}:{ln}\t{n}")


if __name__ == "__main__":
    main()
