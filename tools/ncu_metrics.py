#!/usr/bin/env python
"""Selected metrics of one kernel launch from `ncu --set full` reports -> profiles/<tag>_<name>_ncu_full.csv and the
per-kernel constants bench.py quotes (profiles/<tag>_kernel_metrics.json).

    python tools/ncu_metrics.py r2 name=report.ncu-rep:steps_of_the_profiled_launch ...

`steps` (walk steps of the profiled launch, printed by tools/run_one.py) turns executed instructions into
thread-instructions per walk step.
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_average_branch_targets_threads_uniform.pct", "sm__icc_request_hit_rate.pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_local_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
]
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


def main():
    tag = sys.argv[1]
    table = {}
    for spec in sys.argv[2:]:
        name, rest = spec.split("=", 1)
        rep, _, steps = rest.partition(":")
        v, u = raw(rep)
        with open(ROOT / "profiles" / f"{tag}_{name}_ncu_full.csv", "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow(["metric", "unit", "value"])
            w.writerow(["kernel", "", v.get("Kernel Name", "")])
            for k in KEEP:
                if k in v:
                    w.writerow([k, u.get(k, ""), v[k]])
        f = lambda k: float(v[k].replace(",", ""))  # noqa: E731
        inst, lanes = f("smsp__inst_executed.sum"), f("smsp__thread_inst_executed_per_inst_executed.ratio")
        m = {"kernel": v.get("Kernel Name"), "issue_active_frac": f("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
             "lane_efficiency": lanes / 32.0, "active_lanes": lanes, "icache_hit_rate": f("sm__icc_request_hit_rate.pct") / 100.0,
             "no_instruction_stall_per_issue": f("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
             "registers": int(f("launch__registers_per_thread")), "warp_instructions": inst,
             "dram_bytes_per_launch": f("dram__bytes_read.sum") * SCALE[u["dram__bytes_read.sum"]] + f("dram__bytes_write.sum") * SCALE[u["dram__bytes_write.sum"]],
             "profiled_launch_ms": f("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[u["gpu__time_duration.sum"]],
             "source": f"profiles/{tag}_{name}_ncu_full.csv (ncu --set full --clock-control none, one launch)"}
        if steps:
            m["profiled_launch_steps"] = int(steps)
            m["thread_instructions_per_step"] = inst * lanes / int(steps)
        table[name] = m
        print(name, json.dumps(m))
    out = ROOT / "profiles" / f"{tag}_kernel_metrics.json"
    old = json.loads(out.read_text()) if out.exists() else {}
    old.update(table)
    out.write_text(json.dumps(old, indent=1) + "\n")


if __name__ == "__main__":
    main()
