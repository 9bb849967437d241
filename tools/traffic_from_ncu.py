#!/usr/bin/env python3
"""profiles/traffic.json from an `ncu --set full` report of tools/ab_bench.py: DRAM bytes (read + write) per kernel launch, folded
into bytes per FRAME (k_ptrace / k_ascent / k_trace / k_accum serve a batch of `--batch` frames per launch, k_shade one frame),
stamped with the hash of the terrain-path sources the library was built from (forge3d_b200.build.hot_hash; bench.py refuses a
capture taken from other sources).  usage: python tools/traffic_from_ncu.py REPORT.ncu-rep [--batch 4] [--name profiles/...]"""
import argparse
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--defines", default="")
    args = ap.parse_args()
    from bench import algorithmic_bytes_per_frame
    from forge3d_b200 import build as b

    txt = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    ki, ri, wi, ti = (h.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for r in rows[2:]:
        e = per.setdefault(r[ki], [0, 0.0, 0.0])
        e[0] += 1
        e[1] += float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
        e[2] += float(r[ti])
    out, frame = {}, 0.0
    for k, (n, byt, ms) in sorted(per.items()):
        per_launch = byt / n
        per_frame = per_launch if "k_shade" in k else per_launch / args.batch
        out[k] = {"launches_captured": n, "dram_bytes_per_launch": per_launch, "dram_bytes_per_frame": per_frame, "ms_per_launch": ms / n}
        frame += per_frame
    doc = {"frame_dram_bytes": frame, "per_kernel": out, "hot": b.hot_hash(), "defines": args.defines,
           "source": f"profiles/{Path(args.report).stem}_summary.txt (ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, "
                     f"batch of {args.batch} frames per launch except k_shade)",
           "algorithmic_bytes_per_frame": algorithmic_bytes_per_frame(1920, 1080, 2048)}
    (ROOT / "profiles" / "traffic.json").write_text(json.dumps(doc, indent=1) + "\n")
    print(json.dumps({"frame_dram_bytes": frame, "hot": doc["hot"]}))


if __name__ == "__main__":
    main()
