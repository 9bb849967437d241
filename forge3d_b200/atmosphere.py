"""AETHER atmosphere hand-off for the path-traced snapshot: settings, the shipped LUT bank and the LUT handle.

Host-side stand-in (Python, where the reference is Rust + PyO3) for what sits between
`hybrid_render_terrain_reference(..., atmosphere=...)` and the post pass:
  * `AtmosphereSettings`            python/forge3d/atmosphere.py:25-63
  * `resolve_atmosphere`            extract_atmosphere_lut_handle, src/py_functions/path_tracing/terrain_reference.rs:46-210
  * `AtmosphereConfig.validate`     src/core/atmosphere/bake.rs:132-230
  * `load_shipped`                  load_precomputed_atmosphere_luts + precomputed_bracket, bake.rs:674-770, and the
                                    f16 anchor interpolation of src/core/atmosphere/precomputed.rs:54-133
  * `AtmosphereLutHandle`           src/core/atmosphere/runtime.rs:44-90 (payload + config; what the C ABI's
                                    f3d_atmosphere points into)
The bank itself is a data asset: forge3d_b200/data/aether_bank.npz, packed from the reference's five
turbidity anchors by tools/make_aether_bank.py (SHA-256 checked against precomputed.rs:35-41).  Baking custom LUTs
(`atmosphere_bake_luts`, bake.rs, 2.3 k lines) is out of scope: a custom table enters as
`AtmosphereLutHandle.from_arrays(...)`.
"""
from __future__ import annotations

import math
from collections.abc import Mapping
from dataclasses import dataclass, field, replace
from pathlib import Path
from typing import Any

import numpy as np

TURBIDITY_BANK = (1.0, 2.0, 4.0, 8.0, 10.0)           # precomputed.rs:5
_BANK_PATH = Path(__file__).resolve().parent / "data" / "aether_bank.npz"
_ALLOWED_KEYS = ("enabled", "lut_handle", "turbidity", "ozone_du", "mie_g", "ground_albedo", "scattering_orders")


def _f32(v) -> float:
    return float(np.float32(v))


@dataclass(frozen=True)
class LutDimensions:
    """LutDimensions, bake.rs:31-58 (defaults == the shipped bank's dimensions, :60-73)."""
    transmittance_mu: int = 32
    transmittance_height: int = 8
    scattering_mu_view: int = 17
    scattering_mu_sun: int = 17
    scattering_height: int = 8
    scattering_nu: int = 16
    aerial_distance: int = 8
    aerial_mu_view: int = 8
    aerial_height: int = 8

    def axes(self):
        return (self.transmittance_mu, self.transmittance_height, self.scattering_mu_view, self.scattering_mu_sun,
                self.scattering_height, self.scattering_nu, self.aerial_distance, self.aerial_mu_view, self.aerial_height)

    def validate(self) -> None:
        if any(a < 2 for a in self.axes()):
            raise AtmosphereError("invalid atmosphere configuration: every atmosphere LUT axis must contain at least two samples")
        if any(a > 256 for a in self.axes()):
            raise AtmosphereError("invalid atmosphere configuration: atmosphere LUT axes are capped at 256 samples")


class AtmosphereError(ValueError):
    """AtmosphereError (bake.rs:17-29); the message is the reference's Display text."""


@dataclass(frozen=True)
class AtmosphereConfig:
    """AtmosphereConfig, bake.rs:132-162; scalars are f32 in the reference and are rounded to f32 here."""
    turbidity: float = 2.0
    ozone_du: float = 300.0
    mie_g: float = 0.8
    bottom_radius_m: float = 6_360_000.0
    top_radius_m: float = 6_460_000.0
    rayleigh_scale_height_m: float = 8_000.0
    mie_scale_height_m: float = 1_200.0
    max_aerial_distance_m: float = 160_000.0
    ground_albedo: float = 0.3
    scattering_orders: int = 4
    dimensions: LutDimensions = field(default_factory=LutDimensions)

    def __post_init__(self):
        for name in ("turbidity", "ozone_du", "mie_g", "bottom_radius_m", "top_radius_m", "rayleigh_scale_height_m",
                     "mie_scale_height_m", "max_aerial_distance_m", "ground_albedo"):
            object.__setattr__(self, name, _f32(getattr(self, name)))

    def validate(self) -> None:
        """AtmosphereConfig::validate, bake.rs:165-230 (same order, same text)."""
        bad = lambda msg: AtmosphereError(f"invalid atmosphere configuration: {msg}")
        scalars = (self.turbidity, self.ozone_du, self.mie_g, self.bottom_radius_m, self.top_radius_m,
                   self.rayleigh_scale_height_m, self.mie_scale_height_m, self.max_aerial_distance_m, self.ground_albedo)
        if not all(math.isfinite(v) for v in scalars):
            raise bad("all scalar parameters must be finite")
        if not 1.0 <= self.turbidity <= 10.0:
            raise bad("turbidity must be in [1, 10]")
        if not 0.0 <= self.ozone_du <= 600.0:
            raise bad("ozone must be in [0, 600] DU")
        if not 0.0 <= self.mie_g <= _f32(0.99):
            raise bad("mie_g must be in [0, 0.99]")
        if self.bottom_radius_m <= 0.0 or self.top_radius_m <= self.bottom_radius_m:
            raise bad("top radius must exceed a positive bottom radius")
        if self.rayleigh_scale_height_m <= 0.0 or self.mie_scale_height_m <= 0.0 or self.max_aerial_distance_m <= 0.0:
            raise bad("scale heights and aerial distance must be positive")
        if not 0.0 <= self.ground_albedo <= 1.0:
            raise bad("ground albedo must be in [0, 1]")
        if not 2 <= int(self.scattering_orders) <= 8:
            raise bad("scattering_orders must be in [2, 8]")
        self.dimensions.validate()


@dataclass
class AtmosphereSettings:
    """Physical inputs of the shipped-LUT path (python/forge3d/atmosphere.py:25-63)."""
    turbidity: float = 2.0
    ozone_du: float = 300.0
    mie_g: float = 0.8
    ground_albedo: float = 0.3
    scattering_orders: int = 4

    def __post_init__(self) -> None:
        for name in ("turbidity", "ozone_du", "mie_g", "ground_albedo"):
            if not math.isfinite(float(getattr(self, name))):
                raise ValueError(f"{name} must be finite")
        if not 1.0 <= float(self.turbidity) <= 10.0:
            raise ValueError("turbidity must be in [1.0, 10.0]")
        if not 0.0 <= float(self.ozone_du) <= 600.0:
            raise ValueError("ozone_du must be in [0.0, 600.0]")
        if not 0.0 <= float(self.mie_g) <= 0.99:
            raise ValueError("mie_g must be in [0.0, 0.99]")
        if not 0.0 <= float(self.ground_albedo) <= 1.0:
            raise ValueError("ground_albedo must be in [0.0, 1.0]")
        if isinstance(self.scattering_orders, bool) or not isinstance(self.scattering_orders, int):
            raise TypeError("scattering_orders must be an integer")
        if not 2 <= self.scattering_orders <= 8:
            raise ValueError("scattering_orders must be in [2, 8]")


class AtmosphereLutHandle:
    """Immutable LUT payload + the physical configuration it was made for (runtime.rs:44-90).

    `transmittance` (height, mu, 4), `scattering` (height*nu, mu_sun, mu_view, 4) and `aerial`
    (height, mu_view, distance, 4) hold RGBA16F bit patterns as uint16: the exact bytes the reference uploads as
    textures (aether_post.rs:365-440), x fastest."""

    def __init__(self, config: AtmosphereConfig, transmittance, scattering, aerial, *, precomputed_bracket=None):
        config.validate()
        d = config.dimensions
        want = {"transmittance": (d.transmittance_height, d.transmittance_mu, 4),
                "scattering": (d.scattering_height * d.scattering_nu, d.scattering_mu_sun, d.scattering_mu_view, 4),
                "aerial": (d.aerial_height, d.aerial_mu_view, d.aerial_distance, 4)}
        arrays = {}
        for name, arr, vmax in (("transmittance", transmittance, 1.0), ("scattering", scattering, 65504.0), ("aerial", aerial, 1.0)):
            a = np.asarray(arr)
            if a.dtype not in (np.float16, np.uint16):
                raise TypeError(f"{name} LUT must hold RGBA16F texels (float16 or their uint16 bit patterns), got {a.dtype}")
            # an OWNED copy: the handle is immutable, so neither the caller's array may be frozen nor may a later write
            # to it reach the device behind the checks below
            a = np.array(a.view(np.uint16), dtype=np.uint16, copy=True, order="C")
            if a.shape != want[name]:   # validate_runtime_luts, runtime.rs:222-235
                raise AtmosphereError(f"invalid atmosphere configuration: runtime LUT payload dimensions {a.shape} do not "
                                      f"match metadata {want[name]}")
            vals = a.view(np.float16).astype(np.float32)
            if not np.isfinite(vals).all() or vals.min() < 0.0 or vals.max() > vmax:   # validate_lut_payload, :122-150
                raise AtmosphereError(f"invalid atmosphere configuration: runtime {name} payload components must be finite "
                                      f"and in [0, {vmax}]")
            a = np.ascontiguousarray(a)
            a.setflags(write=False)
            arrays[name] = a
        aer = arrays["aerial"].view(np.float16)
        if (aer[..., :3] != 0).any():   # runtime.rs:264-276
            raise AtmosphereError("invalid atmosphere configuration: runtime aerial-perspective payload must store zero RGB "
                                  "and unit-bounded transmittance alpha")
        self.config = config
        self.transmittance, self.scattering, self.aerial = arrays["transmittance"], arrays["scattering"], arrays["aerial"]
        self.precomputed_bracket = precomputed_bracket

    @classmethod
    def from_arrays(cls, config: AtmosphereConfig, transmittance, scattering, aerial) -> "AtmosphereLutHandle":
        """Adopt an externally baked table (AtmosphereLutHandle::from_luts)."""
        return cls(config, transmittance, scattering, aerial)

    @property
    def byte_size(self) -> int:
        return int(self.transmittance.nbytes + self.scattering.nbytes + self.aerial.nbytes)


_bank = None


def _load_bank():
    global _bank
    if _bank is None:
        if not _BANK_PATH.exists():
            raise RuntimeError(f"{_BANK_PATH} is missing: the shipped AETHER LUT bank is part of the package data "
                               "(tools/make_aether_bank.py regenerates it from the reference's anchors)")
        with np.load(_BANK_PATH) as z:
            _bank = {k: z[k] for k in z.files}
    return _bank


def _precomputed_bracket(t: float):
    """precomputed_bracket, bake.rs:674-688 (f32 arithmetic)."""
    if not TURBIDITY_BANK[0] <= t <= TURBIDITY_BANK[4]:
        raise AtmosphereError(f"precomputed atmosphere bank does not support turbidity {t}; shipped range is [1, 10]")
    for i in range(4):
        a, b = np.float32(TURBIDITY_BANK[i]), np.float32(TURBIDITY_BANK[i + 1])
        if np.float32(t) <= b:
            return i, i + 1, np.float32((np.float32(t) - a) / (b - a))
    return 4, 4, np.float32(0.0)


def _interpolate_f16(name: str, lower: int, upper: int, factor: np.float32) -> np.ndarray:
    """interpolate_f16, precomputed.rs:64-88: exact anchor when the factor selects one, else
    f16(a + (b - a) * factor) evaluated in f32 and rounded to nearest-even f16 (half::f16::from_f32)."""
    bank = _load_bank()
    lo = bank[f"t{int(TURBIDITY_BANK[lower])}_{name}"]
    if lower == upper or factor <= 0.0:
        return lo.copy()
    hi = bank[f"t{int(TURBIDITY_BANK[upper])}_{name}"]
    if factor >= 1.0:
        return hi.copy()
    a = lo.view(np.float16).astype(np.float32)
    b = hi.view(np.float16).astype(np.float32)
    return (a + (b - a) * np.float32(factor)).astype(np.float32).astype(np.float16).view(np.uint16)


def load_shipped(config: AtmosphereConfig | None = None) -> AtmosphereLutHandle:
    """AtmosphereLutHandle::load_shipped(config) (runtime.rs:70-72 -> load_precomputed_atmosphere_luts, bake.rs:690-770):
    only turbidity may differ from the defaults; anything else needs a baked handle and is refused, never substituted."""
    config = AtmosphereConfig() if config is None else config
    config.validate()
    defaults = AtmosphereConfig()
    unsupported = lambda msg: AtmosphereError(f"precomputed atmosphere bank does not support {msg}")
    if config.dimensions != LutDimensions():
        raise unsupported(f"dimensions={config.dimensions}; shipped dimensions are {LutDimensions()}")
    for name in ("ozone_du", "mie_g", "bottom_radius_m", "top_radius_m", "rayleigh_scale_height_m", "mie_scale_height_m",
                 "max_aerial_distance_m", "ground_albedo"):
        a, b = getattr(config, name), getattr(defaults, name)
        if np.float32(a).tobytes() != np.float32(b).tobytes():
            raise unsupported(f"{name}={a}; shipped value is {b}")
    if config.scattering_orders != 4:
        raise unsupported(f"scattering_orders={config.scattering_orders}; shipped value is 4")
    lower, upper, factor = _precomputed_bracket(config.turbidity)
    return AtmosphereLutHandle(config, _interpolate_f16("transmittance", lower, upper, factor),
                               _interpolate_f16("scattering", lower, upper, factor),
                               _interpolate_f16("aerial", lower, upper, factor),
                               precomputed_bracket=(TURBIDITY_BANK[lower], TURBIDITY_BANK[upper]))


def resolve_atmosphere(obj: Any) -> AtmosphereLutHandle | None:
    """extract_atmosphere_lut_handle (terrain_reference.rs:46-210): None, a handle, a mapping or an object with
    AETHER settings -> handle (or None when `enabled` is False).  Same exception types and message text."""
    if obj is None:
        return None
    if isinstance(obj, AtmosphereLutHandle):
        return obj
    is_mapping = isinstance(obj, Mapping)
    if is_mapping:
        for key in obj.keys():
            if not isinstance(key, str):
                raise TypeError("atmosphere mapping keys must be strings")
            if key not in _ALLOWED_KEYS:
                raise ValueError(f"unknown atmosphere setting {key!r}; expected one of {', '.join(_ALLOWED_KEYS)}")
    missing = object()

    def item(name):
        if is_mapping:
            return obj[name] if name in obj else missing
        return getattr(obj, name, missing)

    values = {name: item(name) for name in _ALLOWED_KEYS}
    if not is_mapping and all(v is missing for v in values.values()):
        raise TypeError("atmosphere must be an AtmosphereLutHandle, a mapping, or an object with recognized AETHER settings")
    if values["enabled"] is not missing:
        if not isinstance(values["enabled"], (bool, np.bool_)):
            raise TypeError("atmosphere.enabled must be a bool")
        if not values["enabled"]:
            return None
    handle = values["lut_handle"]
    if handle is not missing and handle is not None:
        if not isinstance(handle, AtmosphereLutHandle):
            raise TypeError("atmosphere.lut_handle must be an AtmosphereLutHandle returned by atmosphere_bake_luts()")
        cfg = handle.config
        for name in ("turbidity", "ozone_du", "mie_g", "ground_albedo"):
            if values[name] is not missing:
                supplied, expected = np.float32(values[name]), np.float32(getattr(cfg, name))
                if supplied.tobytes() != expected.tobytes():
                    raise ValueError(f"atmosphere.{name}={float(supplied)} does not match the exact LUT handle value "
                                     f"{float(expected)}; refusing to substitute or relabel transport")
        if values["scattering_orders"] is not missing and int(values["scattering_orders"]) != cfg.scattering_orders:
            raise ValueError(f"atmosphere.scattering_orders={int(values['scattering_orders'])} does not match the exact LUT "
                             f"handle value {cfg.scattering_orders}; refusing to substitute or relabel transport")
        return handle
    overrides = {}
    for name in ("turbidity", "ozone_du", "mie_g", "ground_albedo"):
        if values[name] is not missing:
            overrides[name] = float(values[name])
    if values["scattering_orders"] is not missing:
        overrides["scattering_orders"] = int(values["scattering_orders"])
    config = replace(AtmosphereConfig(), **overrides)
    try:
        config.validate()
    except AtmosphereError as error:
        raise ValueError(f"invalid AETHER settings: {error}") from None
    try:
        return load_shipped(config)
    except AtmosphereError as error:
        raise RuntimeError(
            f"PROMETHEUS AETHER could not resolve the shipped LUT bank: {error}. Custom physical inputs require "
            "lut_handle=atmosphere_bake_luts(...) from an atmosphere-bake build; no nearby or default LUT was substituted."
        ) from None
