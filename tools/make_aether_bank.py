#!/usr/bin/env python3
"""Packs the reference's shipped AETHER LUT bank into forge3d_b200/data/aether_bank.npz.

Source: /root/reference/src/core/atmosphere/precomputed/turbidity-{1,2,4,8,10}.bin (598 032 bytes each; layout in
src/core/atmosphere/precomputed.rs:5-24: transmittance 32x8, single scattering 17x17x128, accumulated scattering
17x17x128, aerial 8x8x8 - all RGBA16F little-endian - then 4 f32 order deltas).  Each file is checked against the
SHA-256 the reference locks in precomputed.rs:35-41.  The single-scattering table is not consumed by the path-traced
snapshot's AETHER post (aether_post.rs uploads transmittance, accumulated scattering and aerial only) and is dropped.
These are data assets (like tests/golden/mini_dem_128.npy), not source; run in the build container only:
    python tools/make_aether_bank.py [/root/reference]
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
SHA256 = {
    1: "9ead28087343283942d0bf834aecfb7b3a7b0ea513b830731c2cf9bc77a15f0b",
    2: "c6a77bd25241d6123078cace17d9a2181b520c44ac0e871274d755e092e565bc",
    4: "350a1d13863ac0f4a38a3be585e663a8e8c701c14cb5484760cb0d5ccbe772cd",
    8: "56594423699db4a644650e21f231824f19cb52c0abd718c88b7e4f21f00759cf",
    10: "633b77f0a6d8c31a4640e666ba45c7068fa31b7f623b1f711e5117583c1f51f5",
}
T_BYTES = 32 * 8 * 4 * 2
S_BYTES = 17 * 17 * 128 * 4 * 2
A_BYTES = 8 * 8 * 8 * 4 * 2


def main(ref: Path) -> None:
    out = {}
    for t, want in SHA256.items():
        raw = (ref / "src/core/atmosphere/precomputed" / f"turbidity-{t}.bin").read_bytes()
        assert len(raw) == T_BYTES + 2 * S_BYTES + A_BYTES + 16, len(raw)
        assert hashlib.sha256(raw).hexdigest() == want, f"turbidity-{t}.bin does not match the reference's locked hash"
        off = 0
        out[f"t{t}_transmittance"] = np.frombuffer(raw, "<u2", T_BYTES // 2, off).reshape(8, 32, 4)
        off += T_BYTES + S_BYTES                       # skip single scattering
        out[f"t{t}_scattering"] = np.frombuffer(raw, "<u2", S_BYTES // 2, off).reshape(128, 17, 17, 4)
        off += S_BYTES
        out[f"t{t}_aerial"] = np.frombuffer(raw, "<u2", A_BYTES // 2, off).reshape(8, 8, 8, 4)
        off += A_BYTES
        out[f"t{t}_order_deltas"] = np.frombuffer(raw, "<f4", 4, off)
    dst = ROOT / "forge3d_b200" / "data" / "aether_bank.npz"
    dst.parent.mkdir(exist_ok=True)
    np.savez_compressed(dst, **out)
    print(dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    main(Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
