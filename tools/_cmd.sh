# scratch command file for gpurun calls: `gpurun --timeout 1800 -- 'bash tools/_cmd.sh'` (outputs under gpurun_out/)
set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --headline-only --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
for s in cfg5 cfg4 cfg2 cfg1b cfg1a cfg3; do
  timeout 600 ncu --set full --clock-control none -k regex:walk --launch-skip 2 -c 1 -f -o gpurun_out/full_$s python tools/run_one.py $s 4 > gpurun_out/ncu_full_$s.log 2>&1
done
# then here: python tools/ncu_metrics.py r2 cfg5=gpurun_out/full_cfg5.ncu-rep:<steps of pass 2> ...
