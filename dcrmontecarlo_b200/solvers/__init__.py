"""Walk-on-Stars solvers (the CUDA walk kernel behind the reference's solver class)."""
import importlib

__all__ = ["WostSolver_2D"]


def __getattr__(name):
    if name == "WostSolver_2D":
        return importlib.import_module(f"{__name__}.WoStSolver").WostSolver_2D
    raise AttributeError(name)
