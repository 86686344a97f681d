"""Brute-force polyline queries on the GPU — drop-in for the reference's ``geometry/PolylinesSimple.py``.

Every method is one batched launch of the CUDA primitive the walk kernel itself uses
(``csrc/wost_device.cuh``), reached through the C ABI (``include/wost.h``: ``wost_geom_*``).  Results
match the reference's TorchScript functions bit for bit, quirks included: the ray query
returns the *segment* parameter ``s`` (Q1), normals are the left normals of the segment direction (Q3),
the closing vertex of a loop is never a silhouette vertex (Q4).  There is no CPU path.
"""
from __future__ import annotations


import numpy as np
import torch

try:  # package import (dcrmontecarlo_b200.geometry...) or reference-style import (geometry...)
    from .Polylines import PolyLines
    from .. import _native as nat
except ImportError:  # pragma: no cover - reference-style sys.path layout
    from geometry.Polylines import PolyLines
    import _native as nat


def _queries(point):
    """-> (host float32 (B,2), was_single, device of the input tensor)."""
    dev = point.device if isinstance(point, torch.Tensor) else torch.device("cpu")
    q = nat.host_f32(point)
    single = q.ndim == 1
    return q.reshape(-1, 2), single, dev


class PolyLinesSimple(PolyLines):
    def __init__(self, points: torch.Tensor):
        super().__init__(points)
        self._scene = None
        self._scene_key = None
        self._scene_src = None

    # the scene holds the polyline in both the Dirichlet and the Neumann slot so every query is available
    def _scene_for(self):
        """O(1) per query: the scene is rebuilt only when ``points`` is another tensor or was modified in place (torch
        bumps a tensor's version counter on every in-place write) -- not by comparing the vertex data on every call."""
        p = self.points
        dev = nat.current_device()
        if isinstance(p, torch.Tensor):
            key = (p._version, dev)
            if self._scene is None or self._scene_src is not p or self._scene_key != key:
                pts = nat.host_f32(p).reshape(-1, 2)
                self._scene, self._scene_src, self._scene_key = nat.Scene(pts, pts), p, key
            return self._scene
        pts = nat.host_f32(p).reshape(-1, 2)                           # not a tensor (assigned by hand): compare the data
        key = (pts.tobytes(), dev)
        if self._scene is None or self._scene_key != key:
            self._scene, self._scene_src, self._scene_key = nat.Scene(pts, pts), None, key
        return self._scene

    @staticmethod
    def funcToPolyline(func, x_min: float, x_max: float, resolution: float) -> "PolyLinesSimple":
        """Heightmap ``y = func(x)`` sampled every ``resolution``.  Like the reference (:227-240)
        the samples start at 0, not at ``x_min``."""
        x = torch.arange(0, x_max, resolution)
        return PolyLinesSimple(torch.stack((x, func(x)), dim=-1))

    def distance(self, point: torch.Tensor) -> torch.Tensor:
        """Distance to the polyline (reference :214-224 -> distance_to_polyline_jit :26-49)."""
        sc = self._scene_for()
        q, single, dev = _queries(point)
        out = np.empty(len(q), np.float32)
        nat.check(nat.lib().wost_geom_distance(sc.handle, 0, nat.ptr(q), len(q), nat.ptr(out), None, nat.current_stream(sc.device)))
        t = torch.from_numpy(out).to(dev)
        return t[0] if single else t

    def isSilhouette(self, point: torch.Tensor) -> torch.Tensor:
        """Boolean mask over the interior vertices (reference :242-253 -> is_silhouette_jit :52-81)."""
        sc = self._scene_for()
        q, single, dev = _queries(point)
        nv = max(len(self) - 2, 0)
        mask = np.zeros((len(q), nv), np.uint8)
        if nv:
            nat.check(nat.lib().wost_geom_silhouette(sc.handle, 1, nat.ptr(q), len(q), None, nat.ptr(mask), nat.current_stream(sc.device)))
        t = torch.from_numpy(mask.astype(bool)).to(dev)
        return t[0] if single else t

    def silhouetteDistance(self, point: torch.Tensor) -> torch.Tensor:
        """Distance to the closest silhouette vertex, ``inf`` if there is none (reference :255-265 -> :84-102)."""
        sc = self._scene_for()
        q, single, dev = _queries(point)
        out = np.empty(len(q), np.float32)
        nat.check(nat.lib().wost_geom_silhouette(sc.handle, 1, nat.ptr(q), len(q), nat.ptr(out), None, nat.current_stream(sc.device)))
        t = torch.from_numpy(out).to(dev)
        return t[0] if single else t

    def crossProduct2D(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        """``a_x b_y - a_y b_x`` with (2,)/(N,2) broadcasting (reference :267-279 -> :14-23).  Host-side helper."""
        if a.dim() == 1 and b.dim() == 2:
            a = a.unsqueeze(0).expand_as(b)
        elif b.dim() == 1 and a.dim() == 2:
            b = b.unsqueeze(0).expand_as(a)
        return a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]

    def rayIntersection(self, point: torch.Tensor, direction: torch.Tensor) -> torch.Tensor:
        """Per-segment hit parameter, ``inf`` where the ray misses (reference :281-292 -> :105-132)."""
        sc = self._scene_for()
        q, single, dev = _queries(point)
        d = np.ascontiguousarray(np.broadcast_to(nat.host_f32(direction).reshape(-1, 2), q.shape))
        out = np.empty((len(q), len(self) - 1), np.float32)
        nat.check(nat.lib().wost_geom_ray(sc.handle, 1, nat.ptr(q), nat.ptr(d), len(q), nat.ptr(out), nat.current_stream(sc.device)))
        t = torch.from_numpy(out).to(dev)
        return t[0] if single else t

    def intersectPolylines(self, point: torch.Tensor, direction: torch.Tensor, r):
        """First hit within ``r`` -> ``(hit point, normal, True)``, else ``(point + r*dir, 0, False)``
        (reference :294-307 -> intersect_polylines_jit :135-197)."""
        sc = self._scene_for()
        q, single, dev = _queries(point)
        B = len(q)
        d = np.ascontiguousarray(np.broadcast_to(nat.host_f32(direction).reshape(-1, 2), q.shape))
        rr = np.ascontiguousarray(np.broadcast_to(nat.host_f32(r).reshape(-1), (B,)))
        pt, nr = np.empty((B, 2), np.float32), np.empty((B, 2), np.float32)
        found, seg = np.empty(B, np.uint8), np.empty(B, np.int32)
        nat.check(nat.lib().wost_geom_intersect(sc.handle, 1, nat.ptr(q), nat.ptr(d), nat.ptr(rr), B, nat.ptr(pt), nat.ptr(nr),
                                                nat.ptr(found), nat.ptr(seg), nat.current_stream(sc.device)))
        self.last_hit_segment = torch.from_numpy(seg)
        tp, tn = torch.from_numpy(pt).to(dev), torch.from_numpy(nr).to(dev)
        if single:
            return tp[0], tn[0], bool(found[0])
        return tp, tn, torch.from_numpy(found.astype(bool)).to(dev)


class PolyLinesBVH(PolyLinesSimple):
    """The spatially accelerated polyline the reference's abstract base was written for (``geometry/Polylines.py:8-63``;
    SURVEY §8 f-3), under the name a user of the reference would look for.  Every polyline gets its hierarchy automatically
    once it is large enough (implicit tree over the index order with silhouette normal cones, built in ``wost_scene_create``:
    >= 48 Dirichlet / >= 192 Neumann segments, 32-wide cooperative trees from 512 segments) and answers every query with the
    bits of the brute-force loops, so this class adds nothing but ``has_hierarchy``."""

    @property
    def has_hierarchy(self) -> bool:
        return len(self.points) - 1 >= 48
