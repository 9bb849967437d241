#!/usr/bin/env python3
"""Per-source-line view of an `ncu --set full --import-source on` report (no GPU, no matching .so needed: the report carries the
sources).  usage: tools/ncu_lines.py REPORT.ncu-rep [TOP_N]
For every kernel in the report: warp-instructions, average active lanes, and the TOP_N source lines by warp-instructions."""
import collections
import csv
import subprocess
import sys


def num(s):
    try:
        return int(float(s.replace(",", "")))
    except ValueError:
        return 0


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass,cuda"], stdout=subprocess.PIPE, text=True).stdout
    cur = fpath = hdr = None
    data = collections.defaultdict(dict)
    for r in csv.reader(txt.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1]
        elif r[0] == "Function Name":
            cur = r[1]
        elif r[0] == "Line No":
            hdr = r
        elif cur and r[0].isdigit():
            d = dict(zip(hdr, r))
            key = (fpath.split("/")[-1], int(r[0]))
            e = data[cur].setdefault(key, [r[1], 0, 0, 0])
            e[1] += num(d["Instructions Executed"]); e[2] += num(d["Thread Instructions Executed"]); e[3] += num(d["# Samples"])
    for k, v in data.items():
        tot = sum(x[1] for x in v.values())
        ts = sum(x[3] for x in v.values())
        print(f"===== {k}: warp-instr {tot}, lanes {sum(x[2] for x in v.values()) / max(tot, 1):.1f}, samples {ts}")
        for (f, l), (s, i, t, sm) in sorted(v.items(), key=lambda kv: -kv[1][1])[:top]:
            print(f"{100 * i / max(tot, 1):5.1f}% {t / max(i, 1):4.1f} st{100 * sm / max(ts, 1):5.1f}%  {f[4:16]}:{l}  {s.strip()[:120]}")


if __name__ == "__main__":
    main()
