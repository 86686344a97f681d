"""Polyline boundaries for Walk on Stars: the abstract interface and the GPU brute-force / BVH implementation."""
import importlib

_EXPORTS = {"PolyLines": "Polylines", "PolyLinesSimple": "PolylinesSimple", "PolyLinesBVH": "PolylinesSimple"}
__all__ = sorted(_EXPORTS)


def __getattr__(name):
    # resolved on first use so that importing the package does not load the CUDA binding
    if name in _EXPORTS:
        return getattr(importlib.import_module(f"{__name__}.{_EXPORTS[name]}"), name)
    raise AttributeError(name)
