#!/usr/bin/env python3
"""Golden vectors for the straight-alpha compositor, produced by the REFERENCE's own code: the function
`_alpha_composite_rgba` is cut out of /root/reference/python/forge3d/map_scene.py (ast, no package import: forge3d's __init__
needs its native module) and executed on seeded inputs.  Writes tests/golden/alpha_composite_vectors.npz (bottom, top, out).
Run here (the GPU box has no /root/reference); the fixture and this script are committed."""
import ast
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference/python/forge3d/map_scene.py")


def reference_function():
    tree = ast.parse(SRC.read_text())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "_alpha_composite_rgba")
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"Any": object}
    exec(compile(mod, str(SRC), "exec"), ns)
    return ns["_alpha_composite_rgba"], fn.lineno, fn.end_lineno


def main():
    f, l0, l1 = reference_function()
    rng = np.random.default_rng(20261017)
    bottom = rng.integers(0, 256, (96, 128, 4), dtype=np.uint8)
    top = rng.integers(0, 256, (96, 128, 4), dtype=np.uint8)
    # every alpha value, both alpha orderings, the extremes of the colour channels
    top[0, :, 3] = np.arange(128) * 2
    top[1, :, 3] = 255 - np.arange(128) * 2
    bottom[2, :, :3] = 255; top[2, :, :3] = 0
    bottom[3, :, :3] = 0; top[3, :, :3] = 255
    bottom[4:8, :, 3] = 0
    top[8:12, :, 3] = 0
    top[12:16, :, 3] = 255
    out = np.asarray(f(bottom, top))
    np.savez_compressed(ROOT / "tests" / "golden" / "alpha_composite_vectors.npz", bottom=bottom, top=top, out=out)
    print(f"reference function {SRC}:{l0}-{l1}; {out.shape} vectors, sha1 of out:", __import__("hashlib").sha1(out.tobytes()).hexdigest()[:12])


if __name__ == "__main__":
    main()
