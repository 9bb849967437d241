"""forge3d_b200: B200-native (sm_100a CUDA) backend for forge3d's path-traced DEM snapshot path.

Drop-in for `forge3d.hybrid_render_terrain_reference` (python/forge3d/__init__.py:288-293 of the
reference) and nothing else; see DESIGN.md for the scope contract.
"""
from .path_tracing import hybrid_render_terrain_reference

__all__ = ["hybrid_render_terrain_reference"]
__version__ = "0.1.0"
