"""Regenerates the golden fixtures under tests/golden/ from the reference checkout.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_fixtures.py
Fixtures (all small, committed):
  mini_dem_128.npy          forge3d.datasets.mini_dem()[::2, ::2] normalised to [0,1]
                            exactly as tests/test_hybrid_terrain_pt.py:52-60 (_dem()).
  mini_dem_reference.png    byte copy of the reference's golden render
                            tests/golden/hybrid_terrain/mini_dem_reference.png (a data
                            fixture, not source): the only pinned pixels of this path.
"""
import shutil
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
HERE = Path(__file__).resolve().parent


def main():
    sys.path.insert(0, str(REF / "python"))
    from forge3d.datasets import mini_dem  # pure-Python part of the reference package

    dem = mini_dem()[::2, ::2].astype(np.float32)
    dem -= dem.min()
    dem /= max(float(dem.max()), 1e-6)
    np.save(HERE / "mini_dem_128.npy", dem)
    shutil.copyfile(REF / "tests/golden/hybrid_terrain/mini_dem_reference.png", HERE / "mini_dem_reference.png")
    print("wrote", HERE / "mini_dem_128.npy", dem.shape, dem.dtype, float(dem.min()), float(dem.max()))


if __name__ == "__main__":
    main()
