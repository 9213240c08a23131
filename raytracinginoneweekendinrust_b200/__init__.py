"""shimmer-b200: a B200-native path-tracing backend behind the reference's ``Renderer::render``.

The package is a thin host layer over ``lib/libshimmer_b200.so`` (hand-written sm_100a
kernels behind the C ABI of ``include/shimmer_b200.h``).  There is no CPU fallback.
"""
from .capi import Camera, RenderParams, ShimError, Stats  # noqa: F401
from .api import Renderer, Scene, make_params  # noqa: F401

__all__ = ["Camera", "RenderParams", "ShimError", "Stats", "Renderer", "Scene", "make_params"]
