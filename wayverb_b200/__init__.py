"""wayverb_b200 -- B200-native (sm_100a) implementation of wayverb's two GPU hot
loops behind the reference's own entry points.

Python here is only the thinnest host mirror used by tests and bench.py; the
product is libwvb200.so (C ABI, include/wvb200.h) plus the C++ shim in
include/wayverb_b200/ that keeps `waveguide::run` / `raytracer::run`.
"""
from . import _lib  # noqa: F401
from .waveguide import (Mesh, Waveguide, build_mesh, cuboid_mesh, hard_source, soft_source, node_receiver,  # noqa: F401
                        run, slab_range)

from .raytracer import ImageSource, RayTracer, reflection_depth  # noqa: F401,E402
from . import scene  # noqa: F401,E402

__all__ = ["build_mesh", "ImageSource", "RayTracer", "reflection_depth", "scene", "Mesh", "Waveguide", "cuboid_mesh", "hard_source", "soft_source", "node_receiver", "run",
           "slab_range"]
