"""Oracle vs the reference's pyramid unit tests
(src/path_tracing/hybrid_compute/terrain_heightfield.rs:528-612) and refraction model
(src/geo/refraction.rs).  CPU only."""
import numpy as np
import pytest

from oracle import oracle


def ramp(w, h):
    i = np.arange(w * h)
    return ((i % w).astype(np.float32) * np.float32(0.5) + (i // w).astype(np.float32) * np.float32(0.25)).reshape(h, w)


def test_minmax_invariant_per_node():
    levels, _, _ = oracle.build_minmax(ramp(256, 256))
    for lv in levels:
        real = np.isfinite(lv[..., 0]) | np.isfinite(lv[..., 1])
        assert (lv[..., 0][real] <= lv[..., 1][real]).all()


def test_mip_count_and_dims():
    levels, cw, ch = oracle.build_minmax(ramp(256, 256))
    assert (cw, ch) == (255, 255)
    assert levels[0].shape[:2] == (256, 256)
    assert levels[-1].shape[:2] == (1, 1)
    assert len(levels) == 9
    levels, cw, ch = oracle.build_minmax(ramp(100, 37))
    assert (cw, ch) == (99, 36)
    assert levels[0].shape[:2] == (64, 128)  # (h, w) = padded 128 x 64
    assert levels[-1].shape[:2] == (1, 1)
    assert len(levels) == 8


def test_parent_covers_children():
    levels, _, _ = oracle.build_minmax(ramp(64, 64))
    for l in range(1, len(levels)):
        ph, pw = levels[l].shape[:2]
        chh, cww = levels[l - 1].shape[:2]
        for y in range(ph):
            for x in range(pw):
                p = levels[l][y, x]
                for dy in range(2):
                    for dx in range(2):
                        c = levels[l - 1][min(2 * y + dy, chh - 1), min(2 * x + dx, cww - 1)]
                        assert p[0] <= c[0] and p[1] >= c[1]


def test_root_covers_full_range():
    h = ramp(33, 17)
    levels, _, _ = oracle.build_minmax(h)
    assert levels[-1][0, 0, 0] == h.min() and levels[-1][0, 0, 1] == h.max()


def test_flat_dem_is_valid():
    levels, _, _ = oracle.build_minmax(np.full((16, 16), 5.0, np.float32))
    for lv in levels:
        real = np.isfinite(lv[..., 0])
        assert (lv[real] == 5.0).all()
    assert tuple(levels[-1][0, 0]) == (5.0, 5.0)


def test_degenerate_dems_error():
    with pytest.raises(oracle.OracleError, match="at least 2x2"):
        oracle.build_minmax(np.ones((1, 1), np.float32))
    with pytest.raises(oracle.OracleError, match="non-finite"):
        oracle.build_minmax(np.full((2, 2), np.nan, np.float32))


def test_padding_is_sentinel():
    levels, _, _ = oracle.build_minmax(ramp(100, 37))
    l0 = levels[0]
    assert np.isposinf(l0[36:, :, 0]).all() and np.isneginf(l0[36:, :, 1]).all()
    assert np.isposinf(l0[:, 99:, 0]).all() and np.isneginf(l0[:, 99:, 1]).all()


def test_effective_radius_models():
    # src/geo/refraction.rs:146-186 + SURVEY section 9.9 (default inv_two_r_prime ~ 6.8e-8 1/m)
    inv, en = oracle.earth_curvature("ellipsoid", 0.0, 6371008.8, "bennett", 0.13, 1013.25, 15.0, 225.0)
    assert en and 6.7e-8 < inv < 6.9e-8
    inv_flat, en_flat = oracle.earth_curvature("flat", 0.0, 6371008.8, "none")
    assert not en_flat and inv_flat == 0.0
    with pytest.raises(oracle.OracleError, match="flat earth only supports"):
        oracle.earth_curvature("flat", 0.0, 6371008.8, "bennett")
    with pytest.raises(oracle.OracleError, match="less than 1"):
        oracle.earth_curvature("sphere", 0.0, 6371008.8, "effective_radius", 1.0)
    # WGS84 at 45 deg: east-west radius exceeds north-south radius
    a0, _ = oracle.earth_curvature("ellipsoid", 45.0, 0, "none", azimuth_deg=0.0)
    a90, _ = oracle.earth_curvature("ellipsoid", 45.0, 0, "none", azimuth_deg=90.0)
    assert a90 < a0  # larger radius -> smaller 1/(2R)
    s, _ = oracle.earth_curvature("sphere", 0.0, 6371008.8, "effective_radius", 0.13)
    assert s == pytest.approx(0.5 / (6371008.8 / 0.87), rel=1e-6)


def test_pinned_elementary_functions():
    import ctypes as C
    L = oracle.lib()
    s, c = C.c_float(), C.c_float()
    xs = np.linspace(0.0, 2 * np.pi, 20001).astype(np.float32)
    err = 0.0
    for x in xs[::7]:
        L.f3do_sincos(float(x), C.byref(s), C.byref(c))
        err = max(err, abs(s.value - np.sin(np.float64(x))), abs(c.value - np.cos(np.float64(x))))
    assert err < 3e-7
    for y, x in [(0.0, 1.0), (1.0, 0.0), (-1.0, 0.0), (0.5, -0.5), (-0.3, -0.9), (1e-3, 1.0), (2.0, 0.1)]:
        assert abs(L.f3do_atan2(y, x) - np.arctan2(y, x)) < 5e-7
    for v in np.linspace(-1, 1, 401):
        assert abs(L.f3do_acos(float(v)) - np.arccos(v)) < 1e-6
    # f16 round-trip agrees with numpy's IEEE binary16 conversion (RN-even)
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.uniform(-2, 2, 2000), rng.uniform(-70000, 70000, 200),
                           [0.0, 1.0, 65504.0, 65519.9, 65520.0, 6e-8, 2.98e-8, 2.99e-8, 1e-10]]).astype(np.float32)
    for v in vals:
        assert L.f3do_f32_to_f16(float(v)) == int(np.float16(v).view(np.uint16)), v
    for hbits in range(0, 0x7C00, 37):
        assert L.f3do_f16_to_f32(hbits) == float(np.uint16(hbits).view(np.float16))


def test_oracle_build_has_no_fma_contraction():
    """The numerics contract forbids FMA contraction; the oracle is compiled with -ffp-contract=off.  Check the
    machine code: no fused multiply-add instruction may appear in the shared object."""
    import shutil
    import subprocess

    if shutil.which("objdump") is None:
        import pytest

        pytest.skip("objdump not available")
    path = oracle.build()
    asm = subprocess.run(["objdump", "-d", str(path)], capture_output=True, text=True, check=True).stdout
    fused = [l for l in asm.splitlines() if any(m in l for m in ("vfmadd", "vfmsub", "vfnmadd", "vfnmsub"))]
    assert not fused, fused[:5]
    flags = (path.parent / "Makefile").read_text()
    assert "-ffp-contract=off" in flags and "-ffast-math" not in flags.replace("-fno-fast-math", "")


def test_cuda_build_flags_pin_the_numerics_contract():
    from forge3d_b200 import build as fbuild

    flags = " ".join(fbuild.NVCC_FLAGS)
    for needed in ("-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "arch=compute_100a,code=sm_100a", "-lineinfo"):
        assert needed in flags
    assert "--use_fast_math" not in flags and "-use_fast_math" not in flags
