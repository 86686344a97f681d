from .WoStSolver import WostSolver_2D

__all__ = ["WostSolver_2D"]
