"""Abstract polyline boundary — same interface as the reference's ``geometry/Polylines.py:8-63``."""
import torch


class PolyLines:
    """A polyline boundary given by its vertices ``points`` of shape ``(N, 2)``.

    Subclasses answer the five geometric queries Walk on Stars needs.  Queries take one point
    ``(2,)`` like the reference, or a batch ``(B, 2)``.
    """

    def __init__(self, points: torch.Tensor):
        self.points = points

    def __len__(self):
        return self.points.shape[0]

    def __getitem__(self, idx):
        return self.points[idx]

    def _todo(self, name):
        raise NotImplementedError(f"{type(self).__name__} does not implement {name}()")

    def distance(self, point: torch.Tensor) -> torch.Tensor:
        self._todo("distance")

    def isSilhouette(self, point: torch.Tensor) -> torch.Tensor:
        self._todo("isSilhouette")

    def silhouetteDistance(self, point: torch.Tensor) -> torch.Tensor:
        self._todo("silhouetteDistance")

    def rayIntersection(self, point: torch.Tensor, direction: torch.Tensor) -> torch.Tensor:
        self._todo("rayIntersection")

    def intersectPolylines(self, point: torch.Tensor, direction: torch.Tensor, r: float):
        self._todo("intersectPolylines")
