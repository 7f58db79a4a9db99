"""gsrast_b200 — B200 (sm_100a) Gaussian-splat forward rasterizer, a drop-in for GSRast's
splat draw path (CudaRasterizer::Rasterizer::forward contract).  The compute lives in
gsrast_b200/libgsrast_b200.so (sources: gsrast_b200/csrc, C ABI: include/gsrast_b200.h);
this package is the host-side mirror of the reference's operator interface.  No CPU fallback.
"""
from . import camera, scene  # noqa: F401  (numpy only)

__all__ = ["camera", "scene", "rasterizer", "views", "_lib"]


def __getattr__(name):  # torch-dependent modules are loaded on first use
    if name in ("rasterizer", "views", "_lib"):
        import importlib

        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
