"""
fastremap_b200 -- B200-native (sm_100a CUDA) implementation of fastremap's relabel hot path.

Drop-in for the hot-path names of the reference's module surface
(seung-lab/fastremap, fastremap/__init__.py:1-24): `import fastremap_b200 as fastremap`.
Every name of the reference's module surface is served by the CUDA library; by design there is no
CPU fallback anywhere in this package.
"""
from .api import (  # noqa: F401
  component_map,
  foreground,
  indices,
  inverse_component_map,
  mask,
  mask_except,
  minmax,
  pixel_pairs,
  point_cloud,
  refit,
  remap,
  remap_from_array,
  remap_from_array_kv,
  renumber,
  unique,
)
from .layout import ascontiguousarray, asfortranarray, tobytes, transpose  # noqa: F401
from .dtypes import fit_dtype, narrow_dtype, widen_dtype  # noqa: F401
from ._lib import FastremapB200Error  # noqa: F401


__all__ = [
  "ascontiguousarray", "asfortranarray", "component_map", "fit_dtype", "foreground", "indices",
  "inverse_component_map", "mask", "mask_except", "minmax", "narrow_dtype", "pixel_pairs", "point_cloud",
  "refit", "remap", "remap_from_array", "remap_from_array_kv", "renumber", "tobytes", "transpose", "unique",
  "widen_dtype",
]
__version__ = "0.1.0"
