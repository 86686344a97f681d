#!/usr/bin/env python
"""RMSE vs wall time per GPU count (BASELINE metric, second half): run under torchrun with N = 1, 2, 4, 8 ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/rmse_vs_time_multi.py

For the scenarios with an analytic solution, sweeps the walk count, solves with `distributed.solve_sharded` (points or
walk ranges sharded over the ranks, statistics gathered over NCCL) and logs (n_gpus, walks, wall seconds, RMSE).  The
estimates are bit-identical for every N (global Philox counters, fixed-order reduction), so the RMSE column repeats and
only the time column moves.
"""
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dcrmontecarlo_b200 import scenarios as sc  # noqa: E402
from dcrmontecarlo_b200.distributed import solve_sharded  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = []
for key, walks in (("cfg1a", [2400, 38400, 614400, 9830400]), ("cfg3", [2400, 38400, 614400, 2457600]), ("cfg1b", [2400, 38400, 614400])):
    s = sc.ALL[key]()
    solver = s.make_solver()
    exact = s.analytic(s.points).double().cuda()
    solve_sharded(solver, s.points, 1024, s.max_steps, s.eps, seed=1)        # warm-up (handles, NCCL)
    for W in walks:
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        r = solve_sharded(solver, s.points, W, s.max_steps, s.eps, seed=1000 + W)
        torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
        rmse = float(torch.sqrt(((r["mean"] - exact) ** 2).mean()))
        chk = float(r["mean"].sum())
        if rank == 0:
            rec = dict(cfg=key, n_gpus=world, walks=W, wall_s=dt, rmse=rmse, steps=r["steps"], steps_per_s=r["steps"] / dt,
                       by_points=r["by_points"], checksum=chk)
            out.append(rec); print(json.dumps(rec), flush=True)
dist.barrier(); dist.destroy_process_group()
