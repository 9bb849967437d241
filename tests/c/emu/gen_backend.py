#!/usr/bin/env python3
"""Generates tests/c/emu/_build/f3d_backend_emu.cpp from forge3d_b200/csrc/f3d_backend.cu: the file is taken verbatim
except that every `kernel<<<grid, block, smem, stream>>>(args);` becomes a call of the SIMT interpreter
(`emu::launch(grid, block, smem, [&]{ kernel(args); })`).  Test infrastructure: lets the CPU suite run the product's
host driver AND kernels (compiled by g++ against tests/c/emu/cuda_runtime.h) against the oracle."""
import re
import sys
from pathlib import Path


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def transform(src: str) -> str:
    out, pos = [], 0
    for m in re.finditer(r"([A-Za-z_][\w:]*(?:<[^<>;()]*>)?)\s*<<<", src):
        if m.start() < pos:
            continue
        cfg_end = src.index(">>>", m.end())
        cfg = _split_top(src[m.end():cfg_end])
        i = src.index("(", cfg_end)
        depth, j = 0, i
        while True:
            depth += src[j] == "("
            depth -= src[j] == ")"
            if depth == 0:
                break
            j += 1
        args = src[i + 1:j]
        grid, block = cfg[0], cfg[1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        out.append(src[pos:m.start()])
        out.append(f"::emu::launch(dim3({grid}), dim3({block}), (size_t)({smem}), [&]() {{ {m.group(1)}({args}); }})")
        pos = j + 1
    out.append(src[pos:])
    return "".join(out)


if __name__ == "__main__":
    src_path, dst_path = Path(sys.argv[1]), Path(sys.argv[2])
    text = transform(src_path.read_text())
    assert "<<<" not in text
    # dynamic shared memory of the CTA being interpreted (the kernels declare it `extern __shared__`): defined once
    footer = ("\n// dynamic shared memory of the CTA being interpreted (the kernels declare it `extern __shared__`)\n"
              "namespace f3d { alignas(16) unsigned char smem_raw[256 * 1024]; }\n") if src_path.name == "f3d_backend.cu" else ""
    dst_path.write_text(f"// GENERATED from {src_path.name} by tests/c/emu/gen_backend.py - do not edit\n" + text + footer)
