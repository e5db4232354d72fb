"""bvhtracer_b200 -- B200-native closest-hit engine behind bvhtracer's Integrator / Scene::intersect seam.

The product is libbvht_cuda.so (hand-written CUDA for sm_100a behind the C ABI of include/bvht.h).  This
package only binds it (ctypes) and mirrors the reference's host-side interface; it contains no CPU
implementation of the traced path and raises if the CUDA library is missing.
"""
from . import _ffi
from ._ffi import (BVH_NODE, CAMERA, FLAG_FAST, FLAG_LEAF_ACCEL, FLAG_STAMP_INSTANCE, FLAG_STRICT, FLT_MAX, HIT, INSTANCE,
                   MISS_ID, RAY, TLAS_NODE, BvhtError)
from .engine import Engine

__all__ = ["Engine", "BvhtError", "BVH_NODE", "TLAS_NODE", "INSTANCE", "CAMERA", "RAY", "HIT", "FLAG_STRICT", "FLAG_FAST",
           "FLAG_LEAF_ACCEL", "FLAG_STAMP_INSTANCE", "FLT_MAX", "MISS_ID"]
