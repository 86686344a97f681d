"""Multi-GPU path on real GPUs (needs >= 2 devices: `gpurun --gpus 2 -- python -m pytest tests -m gpu`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, key, n_pts, W, ret):
    import torch.distributed as dist

    from dcrmontecarlo_b200 import scenarios as sc
    from dcrmontecarlo_b200.distributed import solve_sharded

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    s = sc.ALL[key]()
    solver = s.make_solver()
    r = solve_sharded(solver, s.points[:n_pts], W, s.max_steps, s.eps, seed=123)
    if rank == 0:
        ret["mean"], ret["m2"], ret["steps"], ret["by_points"] = r["mean"].cpu().numpy(), r["m2"].cpu().numpy(), int(r["steps"]), r["by_points"]
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("key,n_pts,W", [("cfg2", 40, 600), ("cfg5", 1, 4096)])
def test_sharded_results_equal_single_gpu_bits(key, n_pts, W):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from dcrmontecarlo_b200 import scenarios as sc

    with socket.socket() as so:
        so.bind(("127.0.0.1", 0)); port = so.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, key, n_pts, W, ret), nprocs=2, join=True)
    s = sc.ALL[key]()
    single = s.make_solver().solve_raw(s.points[:n_pts], W, s.max_steps, s.eps, seed=123)
    assert ret["by_points"] == (n_pts >= 2)
    assert np.array_equal(ret["mean"], single["mean"]) and np.array_equal(ret["m2"], single["m2"])
    assert ret["steps"] == int(single["steps"][0])


def _survey_worker(rank, world, port, shared, ret):
    import torch.distributed as dist

    from dcrmontecarlo_b200 import scenarios as sc
    from dcrmontecarlo_b200.geometry.PolylinesSimple import PolyLinesSimple
    from dcrmontecarlo_b200.survey import DCRSurvey, DipoleSource

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    s = sc.cfg5(21)
    srcs = [DipoleSource((-30.0 + 4 * k, 0.0), (30.0 - 4 * k, 0.0)) for k in range(5)]
    survey = DCRSurvey(PolyLinesSimple(s.dirichlet), PolyLinesSimple(s.neumann), s.alpha, s.points, srcs, sink_sign=+1.0)
    out = survey.run(nWalks=600, maxSteps=s.max_steps, eps=s.eps, seed=321, shared_walks=shared)
    if rank == 0:
        ret["pot"], ret["steps"] = out["potentials"].copy(), out["steps"]
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("shared", [True, False])
def test_survey_sharded_over_two_gpus_equals_single_gpu_bits(shared):
    """DCRSurvey.run on 2 ranks (electrodes sharded with shared walks, sources round-robin without) == 1 GPU, bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    with socket.socket() as so:
        so.bind(("127.0.0.1", 0)); port = so.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_survey_worker, args=(2, port, shared, ret), nprocs=2, join=True)
    one = {}
    _survey_worker(0, 1, 0, shared, one)
    assert np.array_equal(ret["pot"], one["pot"]) and ret["steps"] == one["steps"]
