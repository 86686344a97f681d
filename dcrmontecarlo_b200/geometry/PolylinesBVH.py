"""``PolyLinesBVH`` — the spatially accelerated polyline the reference's abstract base was written for
(``geometry/Polylines.py:8-63``).

In this implementation every polyline gets its hierarchy automatically once it is large enough (implicit tree over the
index order with silhouette normal cones, built in ``wost_scene_create``: >= 48 Dirichlet / >= 192 Neumann segments,
32-wide cooperative trees from 512 segments), and the queries are bit-identical to the brute-force loops.  The class is
therefore ``PolyLinesSimple`` under the name a user of the reference would look for; ``has_hierarchy`` tells whether
this polyline is large enough to get one.
"""
from __future__ import annotations

try:
    from .PolylinesSimple import PolyLinesSimple
except ImportError:  # pragma: no cover - reference-style sys.path layout
    from geometry.PolylinesSimple import PolyLinesSimple


class PolyLinesBVH(PolyLinesSimple):
    @property
    def has_hierarchy(self) -> bool:
        return len(self.points) - 1 >= 48
