#!/usr/bin/env python
"""bench.py — WoSt walk-steps/s and RMSE-vs-time on B200(s), next to the reference's CPU solver timed in the same run.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-cpu-baseline] [--quick]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Headline workload (config.workload, BASELINE.json configs[4]): the DC-resistivity survey scene of the reference's
tests/testGeophysicalScenario.py:84-151 (Dirichlet square +-100, insulating Neumann surface, Gaussian source pair, two
smooth-circle conductivity anomalies => delta tracking, source term and reflecting boundary: walk_kernel<NEU,SRC,DELTA>)
with eps = 0.9 (the shipped eps = 1.0 takes zero steps), on a line of 175 electrodes x WALKS walks each.
One "step" = one pass of the hot path over that batch through the product's sharded driver
(`dcrmontecarlo_b200.distributed.solve_sharded`): every rank walks its shard and ONE all_gather_into_tensor (NCCL) leaves
the per-electrode statistics on every rank — the collective is inside the timed region.  Weak scaling: walks per
electrode grow with N, per-GPU work is constant.

Prints ONE JSON line (rank 0).  `value` = whole-job walk-steps/s with inputs resident in HBM; `e2e` = the same with HOST
buffers (H2D of the points, D2H of the statistics inside the timed region).  At N = 1 the line also carries, per
reference scenario (cfg 1a / 1b / 2 / 3 / 4 / 5): throughput, e2e, the CPU port, profiler-derived issue / lane figures
and the time-to-solution at the reference's own shipped sizes.  At every N: strong scaling of two fixed global jobs with a
checksum of the gathered estimates (must not depend on N) and an RMSE-vs-time sweep.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ELECTRODES = 175
WALKS = int(os.environ.get("WOST_BENCH_WALKS", 32768))          # walks per electrode per GPU and step
METRIC, UNIT = "wost_walk_steps_per_sec", "walk-steps/s"
WORKLOAD = ("cfg5 DCR survey scene (reference tests/testGeophysicalScenario.py:84-151, eps=0.9, maxSteps=500): Dirichlet square +-100, "
            "Neumann surface, Gaussian source pair, smooth-circle conductivity anomalies, delta tracking; 175 electrodes "
            "(BASELINE.json configs[4])")
KERNEL_METRICS = ROOT / "profiles" / "r2_kernel_metrics.json"    # per-kernel constants read from the committed ncu captures


# ---- algorithmic flop counts -------------------------------------------------------------------------------
def field_flops(f, jet=False):
    """fp32 operations of one evaluation of an analytic field, counted like SURVEY §8(d) (FMA = 2, everything else,
    transcendentals included, = 1); bilinear table = 14."""
    if f is None:
        return 0.0
    desc = f.describe()
    if int(desc["kind"]) == 1:
        return 14.0 * (3 if jet else 1)
    n = 0.0
    for t in desc["terms"]:
        if int(t["kind"]) == 1:                                   # smooth circle: 2 sub, fma, sqrt, sub, mul, exp, add, div, mul
            n += 12.0
        else:
            n += 1.0 + int(t["px"]) + int(t["py"]) + (8.0 if float(t["q"]) != 0.0 else 0.0) + 6.0 * ((int(t["t1"]) != 0) + (int(t["t2"]) != 0))
        n += 1.0                                                  # the sum
    return n * (4.0 if jet else 1.0)


def f_step(s, sp_mode=0):
    SD = len(s.dirichlet) - 1
    f = 27.0 * SD + 12.0
    if s.neumann is not None:
        VN = len(s.neumann)
        f += 23.0 * max(VN - 2, 0) + 24.0 * (VN - 1) + 24.0
    if s.f is not None:
        f += 22.0 + field_flops(s.f)
    if s.delta:
        Fa = field_flops(s.alpha)
        Fsp = (field_flops(s.sigma) + 1.0) if sp_mode == 1 else (field_flops(s.sigma) + field_flops(s.alpha, jet=True) + 15.0)
        f += 25.0 + 3.0 * Fa + 0.5 * Fsp                          # p_int = 0.5 (SURVEY's worked value)
    return f


def launches_per_solve(s):
    """kernels of one wost_solve: [alpha at the evaluation points] + walk + block statistics + merge"""
    return 3 + (1 if s.delta else 0)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.reasons |= {n for bit, n in names.items() if mask & bit}
            except Exception:
                pass
            self._stop_evt.wait(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=10)                                            # an NVML query can take a while under load
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def head_scenario(n_walks=WALKS):
    from dcrmontecarlo_b200 import scenarios as sc

    return sc.cfg5(ELECTRODES, n_walks)


# ---- CPU arms ---------------------------------------------------------------------------------------------------------------
def cpu_port(s, target_seconds=10.0, n_threads=0, sigma_bar=None):
    """The CPU oracle (C port of the reference walk loop, Philox mode, OpenMP over evaluation points) on a bounded sample."""
    from oracle import wost_oracle as orc

    cores = n_threads or len(os.sched_getaffinity(0))
    sb = sigma_bar if sigma_bar is not None else (s.sigma_bar or 0.0)
    prob = orc.Problem.from_scenario(s, sigma_bar=sb)
    pts = s.points
    reps = (max(cores * 2, 16) + len(pts) - 1) // len(pts)
    pts = pts.repeat(reps, 1)[: max(cores * 2, 16)].contiguous()
    prob.solve(pts, 2, s.max_steps, s.eps, rng_mode=orc.RNG_PHILOX, seed=1, n_threads=cores)     # tables, thread pool
    walks, dt, r = 8, 0.0, None
    for _ in range(6):                                                   # grow the sample until it takes about target_seconds
        t0 = time.perf_counter()
        r = prob.solve(pts, walks, s.max_steps, s.eps, rng_mode=orc.RNG_PHILOX, seed=2, n_threads=cores)
        dt = time.perf_counter() - t0
        if dt >= 0.5 * target_seconds or walks >= (1 << 20):
            break
        walks = int(min(1 << 20, max(walks * 2, walks * 0.9 * target_seconds / max(dt, 1e-3))))
    return {"value": r["steps"] / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(pts)} points x {walks} walks of the same scene ({r['steps']} steps in {dt:.2f} s), oracle/wost_oracle.c, OpenMP over points"}


def python_reference_subprocess(walks=8):
    """The unmodified Python reference on this box's cores, in a fresh interpreter (this process holds a CUDA context)."""
    try:
        out = subprocess.run([sys.executable, str(ROOT / "baseline" / "time_reference.py"), "--walks", str(walks), "--electrodes", str(ELECTRODES)],
                             capture_output=True, text=True, timeout=600)
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
        return json.loads(line)
    except Exception as e:                                               # noqa: BLE001
        return {"unavailable": f"baseline/time_reference.py failed: {type(e).__name__}: {e}"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores: the UNMODIFIED Python
    reference from baseline/_ref (all cores, one process each), and beside it the C port; the port alone if the copy of
    the reference did not travel."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "baseline"))
    import time_reference as tr

    s = head_scenario()
    port = cpu_port(s, target_seconds=4.0, sigma_bar=10.0)
    why = tr.available()
    common = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    if why is None:
        cores = len(os.sched_getaffinity(0))
        workers = min(cores, 32)
        pool = tr.ReferencePool(workers, ELECTRODES, 1)
        st, dt, _ = pool.step(2)                                         # calibrate: one step ~ 5 s (includes the per-call cache refill)
        per_walk = max(dt - 2.0, 0.2) / 2.0
        walks = int(max(2, min(64, round(3.0 / per_walk))))
        tot_steps, tot_t = 0, 0.0
        for it in range(args.warmup + args.steps):
            st, dt, _ = pool.step(walks)
            if it >= args.warmup:
                tot_steps += st; tot_t += dt
        pool.close()
        val = tot_steps / tot_t
        sample = (f"each step = {workers} processes x 1 electrode x {walks} walks through the unmodified WostSolver_2D.solve "
                  f"(reference solvers/WoStSolver.py:319-353) imported from baseline/_ref, eps=0.9; solver construction ({pool.ctor_s:.1f} s) not timed")
        line = dict(common, value=val, ms_per_step=1e3 * tot_t / args.steps,
                    config={"workload": WORKLOAD, "electrodes_per_step": workers, "walks_per_electrode": walks},
                    cpu_baseline={"value": val, "unit": UNIT, "cores": workers, "kind": "reference", "sample": sample, "port": port},
                    e2e={"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    else:
        line = dict(common, value=port["value"], ms_per_step=None, config={"workload": WORKLOAD},
                    cpu_baseline=dict(port, python_reference_unavailable=why),
                    e2e={"value": port["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line), flush=True)


# ---- GPU measurements ---------------------------------------------------------------------------------------------------------
def tile_points(points, n):
    reps = (n + len(points) - 1) // len(points)
    return points.repeat(reps, 1)[:n].contiguous()


def time_solves(torch, fn, reps, warm, dist=None):
    """CUDA-event time of `reps` back-to-back calls of fn(i) on the current stream after `warm` warm-ups (ms)."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    if dist is not None:                                                 # collective callers: all ranks start together
        dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    outs = [fn(warm + i) for i in range(reps)]
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), outs


def shipped_size_us(torch, solver, s, points, walks, quick):
    """Time to solution at the reference's own problem size: microseconds per solve, device-resident (CUDA events over
    back-to-back solves) and with host buffers through the C ABI (wall clock); fastest of three blocks, all reported."""
    sp_d = points.cuda()
    r0 = 40 if quick else 100
    dev, host = [], []
    for blk in range(3):
        ms0, _ = time_solves(torch, lambda i: solver.solve_raw(sp_d, walks, s.max_steps, s.eps, seed=300 + 1000 * blk + i, device_outputs=True), r0, 10 if blk == 0 else 2)
        dev.append(ms0 * 1e3 / r0)
        for i in range(3):
            solver.solve_raw(points, walks, s.max_steps, s.eps, seed=i)
        t0 = time.perf_counter()
        for i in range(r0):
            solver.solve_raw(points, walks, s.max_steps, s.eps, seed=400 + 1000 * blk + i)
        host.append((time.perf_counter() - t0) * 1e6 / r0)
    return {"points": len(points), "walks": walks, "device_resident_us": min(dev), "host_buffers_us": min(host),
            "device_resident_us_blocks": dev, "host_buffers_us_blocks": host}


def config_record(torch, nat, name, s, n_points, walks, peak_tf, metrics, quick):
    """One row of the per-scenario table: throughput-sized run, e2e, CPU port, and the reference's shipped size."""
    solver = s.make_solver()
    pts_h = tile_points(s.points, n_points) if n_points else s.points
    pts_d = pts_h.cuda()
    reps = 3 if quick else 5
    ms, outs = time_solves(torch, lambda i: solver.solve_raw(pts_d, walks, s.max_steps, s.eps, seed=100 + i, device_outputs=True)["steps"], reps, 3)
    steps = sum(int(o[0]) for o in outs)
    rate = steps / (ms * 1e-3)
    jit = nat.jit_last_note() == ""
    # e2e: host buffers through the C ABI (H2D of the points, D2H of the statistics inside)
    pin = pts_h.pin_memory()
    best = None
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); n = 0
        for i in range(reps):
            n += int(solver.solve_raw(pin, walks, s.max_steps, s.eps, seed=200 + i)["steps"][0])
        dt = time.perf_counter() - t0
        best = max(best or 0.0, n / dt)
    # the reference's own problem size: time to solution
    shipped = shipped_size_us(torch, solver, s, s.points, s.n_walks, quick)
    F = f_step(s, getattr(solver, "sp_mode", 0))
    cpu = cpu_port(s, target_seconds=1.0 if quick else 2.0, sigma_bar=float(solver.sigma_bar) if s.delta else 0.0)
    m = metrics.get(name, {})
    rec = {"steps_per_s": rate, "ms_per_pass": ms / reps, "points": len(pts_h), "walks": walks, "steps_per_walk": steps / (reps * len(pts_h) * walks),
           "e2e_steps_per_s": best, "specialised_kernel": jit, "kernel": f"walk_kernel<NEU={int(s.neumann is not None)},SRC={int(s.f is not None)},DELTA={int(s.delta)}>",
           "flops_per_walk_step": F, "algorithmic_fp32_frac": rate * F / 1e12 / peak_tf,
           "issue_active_frac": m.get("issue_active_frac"), "lane_efficiency": m.get("lane_efficiency"),
           "thread_instructions_per_step": m.get("thread_instructions_per_step"), "icache_hit_rate": m.get("icache_hit_rate"),
           "metrics_source": m.get("source"),
           "cpu_port": {"steps_per_s": cpu["value"], "cores": cpu["cores"], "sample": cpu["sample"]}, "vs_cpu_port": rate / cpu["value"],
           "shipped_size": shipped}
    return rec


def rmse_sweep(torch, dist_mod, key, walks_list):
    """RMSE over the evaluation points against the analytic solution vs wall time of the sharded solve (host clock, the
    estimates read back to the host), for growing walk counts."""
    from dcrmontecarlo_b200 import scenarios as sc

    s = sc.ALL[key]()
    solver = s.make_solver()
    exact = s.analytic(s.points).double()
    out = []
    dist_mod.solve_sharded(solver, s.points, 64, s.max_steps, s.eps, seed=1)["mean"].cpu()           # warm-up
    for W in walks_list:
        for i in range(2):                                                  # shape warm-up (scratch buffers, specialised kernel)
            dist_mod.solve_sharded(solver, s.points, W, s.max_steps, s.eps, seed=7 + i)["mean"].cpu()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = dist_mod.solve_sharded(solver, s.points, W, s.max_steps, s.eps, seed=1000 + W)
        mean = r["mean"].cpu()
        dt = time.perf_counter() - t0
        out.append({"walks": W, "wall_s": dt, "rmse": float(torch.sqrt(((mean - exact) ** 2).mean())), "checksum": dist_mod.checksum(r["mean"])})
    return out


def strong_scaling(torch, dist, dist_mod, quick):
    """Two FIXED global jobs through the product's sharded paths; times are max over ranks (device events), checksums of the
    gathered estimates must be the same for every N."""
    from dcrmontecarlo_b200 import scenarios as sc
    from dcrmontecarlo_b200.geometry.PolylinesSimple import PolyLinesSimple
    from dcrmontecarlo_b200.survey import DCRSurvey, DipoleSource

    out = {}
    # (i) walk / point sharding: cfg 3 (Poisson source), 404 points x 614 400 walks
    s = sc.cfg3()
    solver = s.make_solver()
    W = 153600 if quick else 614400
    reps = 3
    ms, outs = time_solves(torch, lambda i: dist_mod.solve_sharded(solver, s.points, W, s.max_steps, s.eps, seed=4242), reps, 2, dist)
    t = torch.tensor([ms / reps], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    exact = s.analytic(s.points).double()
    mean = outs[-1]["mean"].cpu()
    out["cfg3_404pts_x_%d_walks" % W] = {"ms": float(t[0]), "steps": int(outs[-1]["steps"]), "steps_per_s": int(outs[-1]["steps"]) / (float(t[0]) * 1e-3),
                                          "checksum": dist_mod.checksum(outs[-1]["mean"]), "rmse": float(torch.sqrt(((mean - exact) ** 2).mean())),
                                          "sharding": "points" if outs[-1]["by_points"] else "walks"}
    # (ii) DCR survey: 175 electrodes x 64 source dipoles, shared walks, electrodes sharded over the ranks
    c5 = sc.cfg5(ELECTRODES)
    srcs = [DipoleSource((-38.0 + 1.1 * k, 0.0), (38.0 - 1.1 * k, 0.0)) for k in range(64)]
    survey = DCRSurvey(PolyLinesSimple(c5.dirichlet), PolyLinesSimple(c5.neumann), c5.alpha, c5.points, srcs, sink_sign=+1.0)
    Ws = 4096 if quick else 65536                                         # >= 1.4 M walks per GPU at N = 8: the tail of the longest walks stays small
    res = None
    for i in range(2):
        survey.run(nWalks=Ws, maxSteps=c5.max_steps, eps=c5.eps, seed=99, shared_walks=True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        res = survey.run(nWalks=Ws, maxSteps=c5.max_steps, eps=c5.eps, seed=99, shared_walks=True)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["cfg5_survey_175e_x_64src_x_%d_walks" % Ws] = {"ms": float(t[0]), "steps": int(res["steps"]), "steps_per_s": res["steps"] / (float(t[0]) * 1e-3),
                                                       "source_evaluations_per_s": res["steps"] * 64 / (float(t[0]) * 1e-3),
                                                       "checksum": dist_mod.checksum(torch.from_numpy(res["potentials"])), "sharding": "electrodes (shared walks)"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="smaller side measurements (per-config table, strong-scaling jobs)")
    ap.add_argument("--headline-only", action="store_true", help="only the timed headline steps (profiling target)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from dcrmontecarlo_b200 import _native as nat
    from dcrmontecarlo_b200 import distributed as dm
    from dcrmontecarlo_b200 import scenarios as sc

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    nat.require_cuda()
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- headline: weak scaling, walks per electrode grow with the number of GPUs --------------------------------------
    Wg = WALKS * world
    s = head_scenario(Wg)
    solver = s.make_solver()
    pts_dev = s.points.cuda()
    pts_pin = s.points.pin_memory()

    def step_resident(it):
        return dm.solve_sharded(solver, pts_dev, Wg, s.max_steps, s.eps, seed=1000 + it)

    def step_e2e(it):
        r = dm.solve_sharded(solver, pts_pin, Wg, s.max_steps, s.eps, seed=1000 + it)
        return r["mean"].cpu(), int(r["steps"])                          # D2H of the estimates and the step count

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for it in range(args.warmup):
        step_resident(it)
    jit_used = nat.jit_last_note() == ""
    sampler = ClockSampler(local)                                        # nvmlInit BEFORE the barrier: it takes milliseconds, differently on
    sync_all()                                                           # every rank, and the first gather would wait for the last rank to start
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    outs = []
    ev0.record()
    for it in range(args.steps):
        outs.append(step_resident(args.warmup + it))
        marks[it].record()                                               # per-step times (diagnostic; the figure is ev0 -> ev1)
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    each = torch.tensor([(ev0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]) for i in range(args.steps)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(each, op=dist.ReduceOp.MAX)
    each_ms = [round(float(v), 4) for v in each.tolist()]
    tot_steps = int(sum(int(o["steps"]) for o in outs))                   # already the sum over all ranks (it travels in the gather)
    head_checksum = dm.checksum(outs[-1]["mean"])
    clocks = sampler.stop()                                              # the clocks belong to the device-timed region
    if dist is not None:                                                 # every rank samples its own GPU: report the slowest, all reasons
        allc = [None] * world
        dist.all_gather_object(allc, clocks)
        mhz = [c["sm_mhz"] for c in allc if c["sm_mhz"] is not None]
        clocks = dict(clocks, sm_mhz=min(mhz) if mhz else None, sm_mhz_per_rank=[c["sm_mhz"] for c in allc],
                      reasons=sorted(set().union(*[set(c["reasons"]) for c in allc])))
    shard = outs[-1]["shard"]

    # ---- end to end: host buffers -----------------------------------------------------------------------------------------
    for it in range(args.warmup):
        step_e2e(it)
    gc.collect(); gc.disable()                                           # no collector pauses inside the host-timed region
    e2e_passes = []
    for rep in range(3):
        sync_all()
        t0 = time.perf_counter(); n_steps = 0
        for it in range(args.steps):
            n_steps += step_e2e(args.warmup + it)[1]
        torch.cuda.synchronize()
        e2e_passes.append((1e3 * (time.perf_counter() - t0), n_steps))
    gc.enable()
    e2e_ms, e2e_steps = min(e2e_passes)
    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])

    if args.headline_only:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": tot_steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                              "ms_per_step": ms / args.steps, "ms_each_step_max_over_ranks": each_ms, "specialised_kernel": jit_used,
                              "e2e_ms_per_step": e2e_ms / args.steps, "clocks": clocks}), flush=True)
        if dist is not None:
            dist.barrier(); dist.destroy_process_group()
        return

    # ---- side measurements (every rank takes part in the sharded ones) ----------------------------------------------------
    strong = strong_scaling(torch, dist, dm, args.quick)
    sweeps = {"cfg1a": rmse_sweep(torch, dm, "cfg1a", [150, 2400, 38400, 614400] if not args.quick else [150, 2400, 38400]),
              "cfg3": rmse_sweep(torch, dm, "cfg3", [150, 2400, 38400, 153600] if not args.quick else [150, 2400])}

    if rank == 0:
        value = tot_steps / (ms * 1e-3)
        peak_tf, eff_mhz = nat.fp32_peak(local)
        metrics = json.loads(KERNEL_METRICS.read_text()) if KERNEL_METRICS.exists() else {}
        F = f_step(s, solver.sp_mode)
        per_gpu_rate = value / world
        achieved_tf = per_gpu_rate * F / 1e12
        hm = metrics.get("cfg5", {})
        n_launch = launches_per_solve(s)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "ms_each_step_max_over_ranks": each_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "electrodes": ELECTRODES, "walks_per_electrode": Wg, "walks_per_electrode_per_gpu": WALKS,
                       "walk_steps_per_step": tot_steps / args.steps,
                       "l2_note": "every step uses a fresh Philox key and rewrites its per-walk totals (%d MiB per GPU); the working set is registers / shared memory, not L2" % (ELECTRODES * WALKS * 4 >> 20),
                       "parallelism": f"distributed.solve_sharded over {world} GPU(s): rank 0 shard = electrodes range({shard.p0},{shard.p1},{shard.pstride}) x walks [{shard.w0},{shard.w1}); one all_gather_into_tensor of (mean, M2, steps) per step inside the timed region",
                       "compat": "reference", "specialised_kernel": jit_used, "checksum_last_step": head_checksum},
            "roofline": {"bound": "fp32-issue", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                         "frac_note": "ALGORITHMIC fp32 flops of the reference's formulation (SURVEY §8(d)) per second over the measured FMA-chain peak: a throughput-equivalent figure, not pipe utilisation; the kernel is bound by instruction issue, see issue_active_frac",
                         "flops_per_walk_step": F, "issue_active_frac": hm.get("issue_active_frac"), "lane_efficiency": hm.get("lane_efficiency"),
                         "thread_instructions_per_step": hm.get("thread_instructions_per_step"), "icache_hit_rate": hm.get("icache_hit_rate"),
                         "metrics_source": hm.get("source"),
                         "traffic": hm.get("dram_bytes_per_launch"),
                         "traffic_note": f"algorithmic bytes per launch = {ELECTRODES * 8 + ELECTRODES * WALKS * 4} (points in, per-walk totals out, consumed from L2 by the statistics kernel); HBM is idle on this path",
                         "peak_source": "FMA-chain microbenchmark in this run (wost_fp32_peak); MEASURED_PEAKS.json has no fp32 entry",
                         "kernel": "wost_walk_jit = walk_body<NEU=1,SRC=1,DELTA=1> specialised for this solver's fields (NVRTC)" if jit_used else "walk_kernel<NEU=1,SRC=1,DELTA=1>",
                         "fp32_peak_effective_sm_mhz": eff_mhz},
            "e2e": {"value": e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": shard.n_points * 8 * world,
                    "d2h_bytes_per_step": (ELECTRODES * 8 + 8) * world, "ms_per_step": e2e_ms / args.steps,
                    "passes_ms_per_step": [p[0] / args.steps for p in e2e_passes]},
            "gpu_launches": n_launch * args.steps * world,
            "gpu_launches_note": f"{n_launch} kernels per rank and step (alpha at the electrodes, walk, block statistics, merge); NCCL's gather kernel not counted",
            "clocks": clocks,
            "strong_scaling": strong,
            "rmse_vs_time": sweeps,
        }
        if world == 1:
            table = {}
            N = 16384 if args.quick else 65536
            plan = [("cfg1a", sc.cfg1a(), N, 256), ("cfg1b", sc.cfg1b(), N, 64), ("cfg2", sc.cfg2(), N, 256), ("cfg3", sc.cfg3(), N, 256),
                    ("cfg4", sc.cfg4(), N, 64), ("cfg5", sc.cfg5(ELECTRODES, 100), 0, 8192 if args.quick else WALKS)]
            for name, scn, n_pts, walks in plan:
                if name == "cfg5":
                    scn9 = sc.cfg5(9, 100)
                    rec = config_record(torch, nat, name, scn, n_pts, walks, peak_tf, metrics, args.quick)
                    # the reference ships this scene with 9 electrodes x 100 walks (tests/testGeophysicalScenario.py:109-149)
                    rec["shipped_size"] = shipped_size_us(torch, scn9.make_solver(), scn9, scn9.points, 100, args.quick)
                else:
                    rec = config_record(torch, nat, name, scn, n_pts, walks, peak_tf, metrics, args.quick)
                table[name] = rec
            line["configs"] = table
            line["small_solve_us"] = {"cfg5_9x100": table["cfg5"]["shipped_size"], "cfg1b_16x150": table["cfg1b"]["shipped_size"]}
            if not args.no_cpu_baseline:
                port = cpu_port(s, target_seconds=8.0, sigma_bar=float(solver.sigma_bar))
                line["cpu_baseline"] = dict(port, python_reference=python_reference_subprocess())
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
