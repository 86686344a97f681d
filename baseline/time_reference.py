#!/usr/bin/env python
"""Times the UNMODIFIED Python reference (Tsuchijo/DCRMonteCarlo) on the host cores.  Measurement infrastructure only.

The reference is pure Python; `__graft_entry__.build()` copies its sources (solvers/, geometry/, utils.py, __init__.py,
tests/) from /root/reference into the git-ignored `baseline/_ref/`, which travels to the GPU box with the snapshot.  This
script imports it from there with a 3-file matplotlib stub (the reference imports matplotlib at module scope, utils.py:7-8,
and it is not installed), builds the DC-resistivity scene of `tests/testGeophysicalScenario.py:84-139` with the script's
OWN callables (`dcr_current_source`, `conductivity_field`), and runs `WostSolver_2D.solve` (solvers/WoStSolver.py:319-353)
in `workers` processes, one per core, each on its own evaluation point(s).  eps = 0.9 because the shipped eps = 1.0 takes
zero steps.  Steps are counted by wrapping `dirichletBoundary.distance`, which the walk loop calls exactly
once per step (solvers/WoStSolver.py:208).

    python baseline/time_reference.py [--workers N] [--walks W] [--points-per-worker P] [--repeats K] [--electrodes 175]

Prints one JSON line: steps/s over all workers (sum of steps / longest worker time), per-core rate, sizes, core count.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time
import warnings
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = HERE / "_ref"


def available() -> str | None:
    """None if the reference copy is there, else why not."""
    need = [REF / "solvers" / "WoStSolver.py", REF / "geometry" / "PolylinesSimple.py", REF / "utils.py", REF / "tests" / "testGeophysicalScenario.py"]
    missing = [str(p.relative_to(HERE)) for p in need if not p.exists()]
    return ("baseline/_ref is incomplete (run __graft_entry__.build() where /root/reference exists): missing " + ", ".join(missing)) if missing else None


def _import_reference():
    stub = Path(tempfile.mkdtemp(prefix="mplstub_")) / "matplotlib"
    stub.mkdir()
    (stub / "__init__.py").write_text("")
    (stub / "pyplot.py").write_text("Figure = object\n")
    (stub / "patches.py").write_text("Circle = object\n")
    sys.path.insert(0, str(stub.parent))
    sys.path.insert(0, str(REF))
    warnings.filterwarnings("ignore")
    os.environ["TQDM_DISABLE"] = "1"
    spec = importlib.util.spec_from_file_location("ref_testGeophysicalScenario", REF / "tests" / "testGeophysicalScenario.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                       # defines the callables; its __main__ block does not run
    return mod


def _build_solver(mod):
    """tests/testGeophysicalScenario.py:84-139, with the script's own callables."""
    import torch

    half = 100.0
    dirichlet = torch.tensor([[-half, -half], [half, -half], [half, half], [-half, half], [-half, -half]])
    neumann = torch.tensor([[-half, half], [half, half]])
    solver = mod.WostSolver_2D(dirichletBoundary=mod.PolyLinesSimple(dirichlet), dirichletBoundaryFunction=lambda point: 0.0,
                               neumannBoundary=mod.PolyLinesSimple(neumann), source=mod.dcr_current_source,
                               alpha=mod.conductivity_field, sigma=None)
    counter = {"steps": 0}
    inner = solver.dirichletBoundary.distance

    def counted(point):
        counter["steps"] += 1
        return inner(point)

    solver.dirichletBoundary.distance = counted
    return solver, counter


def _worker(idx, n_workers, n_electrodes, pts_per_worker, conn):
    import io
    import contextlib

    import torch

    torch.set_num_threads(1)
    torch.manual_seed(42 + idx)
    import numpy as np

    np.random.seed(42 + idx)
    os.environ["TQDM_DISABLE"] = "1"
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        mod = _import_reference()
        t0 = time.perf_counter()
        solver, counter = _build_solver(mod)
        ctor_s = time.perf_counter() - t0
    xs = torch.linspace(-40.0, 40.0, n_electrodes)
    first = (idx * n_electrodes) // n_workers
    sel = [(first + k) % n_electrodes for k in range(pts_per_worker)]
    pts = torch.stack([xs[sel], torch.zeros(len(sel))], dim=1)
    conn.send(("ready", ctor_s))
    while True:
        msg = conn.recv()
        if msg is None:
            break
        walks = msg
        counter["steps"] = 0
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):   # the solver's tqdm bar
            solver.solve(pts, nWalks=walks, maxSteps=500, eps=0.9)
        conn.send((counter["steps"], time.perf_counter() - t0))


class ReferencePool:
    """`workers` processes, each holding its own reference solver."""

    def __init__(self, workers: int, n_electrodes: int = 175, pts_per_worker: int = 1):
        ctx = mp.get_context("fork")
        self.workers, self.procs, self.conns = workers, [], []
        for i in range(workers):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(i, workers, n_electrodes, pts_per_worker, b), daemon=True)
            p.start()
            self.procs.append(p); self.conns.append(a)
        self.ctor_s = max(c.recv()[1] for c in self.conns)

    def step(self, walks: int):
        """One solve per worker, all at once.  Returns (total steps, wall seconds, per-worker seconds)."""
        t0 = time.perf_counter()
        for c in self.conns:
            c.send(walks)
        res = [c.recv() for c in self.conns]
        return sum(r[0] for r in res), time.perf_counter() - t0, [r[1] for r in res]

    def close(self):
        for c in self.conns:
            c.send(None)
        for p in self.procs:
            p.join(timeout=10)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=0)
    ap.add_argument("--walks", type=int, default=8)
    ap.add_argument("--points-per-worker", type=int, default=1)
    ap.add_argument("--repeats", type=int, default=1)
    ap.add_argument("--electrodes", type=int, default=175)
    a = ap.parse_args()
    why = available()
    if why:
        print(json.dumps({"unavailable": why})); return
    cores = len(os.sched_getaffinity(0))
    workers = a.workers or min(cores, 32)
    pool = ReferencePool(workers, a.electrodes, a.points_per_worker)
    steps = secs = 0.0
    per = []
    for _ in range(a.repeats):
        s, t, w = pool.step(a.walks)
        steps += s; secs += t; per += w
    pool.close()
    print(json.dumps({"value": steps / secs, "unit": "walk-steps/s", "cores": workers, "host_cores": cores, "kind": "reference",
                      "per_core_value": steps / sum(per), "ctor_seconds": pool.ctor_s,
                      "sample": f"{workers} processes x {a.points_per_worker} electrode(s) x {a.walks} walks x {a.repeats} solve() call(s) of the unmodified "
                                f"Python reference (solvers/WoStSolver.py:319-353; each call refills its 10 000-sample radius cache, solvers/utils.py:181-195), "
                                f"DCR scene of tests/testGeophysicalScenario.py with eps=0.9: {int(steps)} steps in {secs:.1f} s"}))


if __name__ == "__main__":
    main()
