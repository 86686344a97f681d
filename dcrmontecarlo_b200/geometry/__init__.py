from .Polylines import PolyLines
from .PolylinesSimple import PolyLinesSimple

__all__ = ["PolyLines", "PolyLinesSimple"]
