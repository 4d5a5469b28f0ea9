"""unnatural_planets_b200 -- B200-native generation core for unnatural-planets.

The product is the C-ABI CUDA library `libupg.so` (include/upg.h, sources in csrc/).  This package is
the thin Python binding used by the tests and bench.py; it holds no compute of its own and there is
no CPU fallback: importing works anywhere, every compute call needs the library and a CUDA device.
"""
from .capi import FIELDS, UpgError, library_path, load_library  # noqa: F401
from .planet import ELEVATIONS, SHAPES, COLORINGS, DEFAULT_SEED, Planet, dilate, isosurface, raster, mesh_clean, mesh_chunking, png_encode, slab_range, stitch_slabs, texture_tail  # noqa: F401
