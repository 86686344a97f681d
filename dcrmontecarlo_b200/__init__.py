"""B200-native Walk-on-Stars engine behind the DCRMonteCarlo Python API.

Sub-modules mirror the reference layout (``solvers.WoStSolver``, ``geometry.PolylinesSimple``,
``geometry.Polylines``, ``utils``); put this directory on ``sys.path`` to import them under the
reference's own module names.
"""
__version__ = "0.1.0"
