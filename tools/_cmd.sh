set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r2b_pytest_gpu.log
# small-solve layout sweep
for L in 32 0 1 2 4 8; do echo "== WOST_LANES_PER_WARP=$L"; WOST_LANES_PER_WARP=$L python tools/small_solve.py 100 2>&1 | grep '"jit": true\| on '; done > gpurun_out/r2b_small_sweep.txt 2>&1
for S in 1 4; do echo "== auto, WOST_SMALL_WARPS_PER_SCHEDULER=$S"; WOST_SMALL_WARPS_PER_SCHEDULER=$S python tools/small_solve.py 100 2>&1 | grep ' on '; done >> gpurun_out/r2b_small_sweep.txt 2>&1
# steps of the profiled passes (seed = pass index) and occupancy A/B
for s in cfg5 cfg5_175e cfg4 cfg2 cfg1b cfg1a cfg3; do python tools/run_one.py $s 4; done > gpurun_out/r2b_run_one.txt 2>&1
for mb in 4 5 6; do for s in cfg5 cfg4 cfg2 cfg1b; do echo -n "minblocks=$mb "; WOST_JIT_MIN_BLOCKS=$mb python tools/run_one.py $s 4 | tail -1; done; done > gpurun_out/r2b_minblocks.txt 2>&1
for s in cfg5 cfg1a cfg3; do
  timeout 600 ncu --set full --clock-control none -k regex:walk --launch-skip 2 -c 1 -f -o gpurun_out/r2_full_$s python tools/run_one.py $s 4 > gpurun_out/r2_ncu_full_$s.log 2>&1
done
ls gpurun_out
